// librloa_b200: batched manipulator simulator — kernels and C ABI (include/rloa_b200.h).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "philox.cuh"
#include "sim_device.cuh"

namespace rloa {

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
enum { kModeStep = 0, kModeResetOnly = 1 };

// launch 1 of an env-step: thread = arm
// Block = 1 warp for large batches (measured best at 131,072 arms) or 4 warps for small ones: at 4096 arms every warp is
// latency-bound, four of them on one SM (one per scheduler) run as fast as alone, and the launch then occupies 32 SMs instead
// of 128 - the 16-CTA learn kernel that runs beside it needs 16 EMPTY SMs (a CTA takes all registers of one).
constexpr int kDynWideTpb = 128;
template <int NLMAX>
__global__ void __launch_bounds__(kDynWideTpb) sim_dynamics_kernel(const __grid_constant__ ModelDev M, SimArrays S) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env < S.n_envs) arm_dynamics<NLMAX>(M, S, env);
}

// collision phase of a stepSimulation: the contact rows of the CURRENT pose against the obstacle sphere and the
// target cube -> crow / near.  Runs with the action-independent half of the step (rloa_sim_prepare: beside the NAF update),
// so the solve kernel finds its rows ready and its own distance pass looks for hits only; thread = arm.
template <int NLMAX>
__global__ void __launch_bounds__(kDynWideTpb) sim_contacts_kernel(const __grid_constant__ ModelDev M, SimArrays S, float thr) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= S.n_envs) return;
    const V3 tg = ld3(S.target + 3 * (size_t)env), ob = ld3(S.obstacle + 3 * (size_t)env);
    const ObsOut o = fk_and_distances<NLMAX, false>(M, S.q + env, S.n_envs, ob, tg, 0.f, false, nullptr, 0, thr, true,
                                                    S.crow + (size_t)env * kMaxContacts * kContactRec);
    S.near[env] = o.ncontacts;
}

// launch 2: block = 32 arms x ndof columns, warp = one column of M^-1 for 32 arms.  The factor records of the
// block's 32 arms are staged once in shared memory ([link][component][arm], conflict-free) by all warps, so the
// ndof column warps stop re-reading them from L2 and the serial tree walks see shared-memory latency.
__global__ void __launch_bounds__(32 * kMaxDof) sim_minv_kernel(const __grid_constant__ ModelDev M, SimArrays S) {
    extern __shared__ float smF[];
    const int env0 = blockIdx.x * 32;
    const int N = S.n_envs, nrec = M.nl * kFRec;
    const int nthreads = 32 * blockDim.y, flat = threadIdx.y * 32 + threadIdx.x;
    for (int idx = flat; idx < nrec * 32; idx += nthreads) {
        const int rec = idx >> 5, arm = idx & 31;
        smF[idx] = env0 + arm < N ? S.F[(size_t)rec * N + env0 + arm] : 0.f;
    }
    __syncthreads();
    const int env = env0 + threadIdx.x;
    if (env < N) arm_minv_column(M, S, env, threadIdx.y, smF + threadIdx.x, 32, kFRec * 32);
}

// M^-1 and the velocity change dv of one arm, in registers, with two layouts:
//   PACKED (ndof <= 12): full rows as float2 pairs, so a row update dv += delta * M^-1[d][:] is D/2 packed
//       FFMA2 (fma.rn.f32x2, new on sm_100) instead of D scalar FFMAs — the Gauss-Seidel sweep is FMA-issue bound;
//   scalar (ndof <= 16): lower triangle only (136 registers), scalar FFMAs.
__device__ __forceinline__ constexpr int tri(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

template <int D, bool PACKED>
struct PgsState;

template <int D>
struct PgsState<D, false> {
    float Mi[D * (D + 1) / 2];
    float dv[D];
    __device__ __forceinline__ void load(const float* __restrict__ minv, int ndof, size_t sN, int env) {
#pragma unroll
        for (int a = 0; a < D; a++) {
            dv[a] = 0.f;
#pragma unroll
            for (int b = 0; b <= a; b++) Mi[tri(a, b)] = a < ndof ? minv[((size_t)a * ndof + b) * sN + env] : 0.f;
        }
    }
    __device__ __forceinline__ float diag(int d) const { return Mi[tri(d, d)]; }
    __device__ __forceinline__ float get(int d) const { return dv[d]; }
    __device__ __forceinline__ float mat(int a, int b) const { return Mi[tri(a, b)]; }
    __device__ __forceinline__ void add(int d, float x) { dv[d] += x; }
    // contact rows: J and U = M^-1 J^T are stored as pairs (dof 2k, 2k + 1), zero beyond D
    __device__ __forceinline__ float cdot(const float2* __restrict__ j) const {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int k = 0; k < D; k++) {
            const float jk = (k & 1) ? j[k >> 1].y : j[k >> 1].x;
            if (k & 1) a1 = fmaf(jk, dv[k], a1); else a0 = fmaf(jk, dv[k], a0);
        }
        return a0 + a1;
    }
    __device__ __forceinline__ void caxpy(float delta, const float2* __restrict__ u) {
#pragma unroll
        for (int k = 0; k < D; k++) dv[k] = fmaf(delta, (k & 1) ? u[k >> 1].y : u[k >> 1].x, dv[k]);
    }
    // dv += delta * M^-1[d][:]; the component the next row reads (`first`) is issued first
    __device__ __forceinline__ void axpy(int d, float delta, int first) {
        if (first >= 0 && first < D) dv[first] = fmaf(delta, Mi[tri(d, first)], dv[first]);
#pragma unroll
        for (int k = 0; k < D; k++)
            if (k != first) dv[k] = fmaf(delta, Mi[tri(d, k)], dv[k]);
    }
};

template <int D>
struct PgsState<D, true> {
    static constexpr int DP = (D + 1) / 2;
    float2 M[D][DP];
    float2 dv[DP];
    __device__ __forceinline__ void load(const float* __restrict__ minv, int ndof, size_t sN, int env) {
#pragma unroll
        for (int k = 0; k < DP; k++) dv[k] = make_float2(0.f, 0.f);
#pragma unroll
        for (int a = 0; a < D; a++)
#pragma unroll
            for (int k = 0; k < DP; k++) {
                const int b0 = 2 * k, b1 = 2 * k + 1;
                M[a][k].x = (a < ndof && b0 < ndof) ? minv[((size_t)a * ndof + b0) * sN + env] : 0.f;
                M[a][k].y = (a < ndof && b1 < ndof) ? minv[((size_t)a * ndof + b1) * sN + env] : 0.f;
            }
    }
    __device__ __forceinline__ float diag(int d) const { return (d & 1) ? M[d][d >> 1].y : M[d][d >> 1].x; }
    __device__ __forceinline__ float get(int d) const { return (d & 1) ? dv[d >> 1].y : dv[d >> 1].x; }
    __device__ __forceinline__ float mat(int a, int b) const { return (b & 1) ? M[a][b >> 1].y : M[a][b >> 1].x; }
    __device__ __forceinline__ void add(int d, float x) {
        if (d & 1) dv[d >> 1].y += x; else dv[d >> 1].x += x;
    }
    // contact rows: J and U = M^-1 J^T are stored as the same pairs as dv (zero beyond D), so the row's J dv is DP packed
    // FFMA2 in two chains and its update another DP - a third of the scalar instruction count of a row visit, which is what a
    // lone warp per SM pays for
    __device__ __forceinline__ float cdot(const float2* __restrict__ j) const {
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < DP; k++) {
            if (k & 1) a1 = __ffma2_rn(j[k], dv[k], a1); else a0 = __ffma2_rn(j[k], dv[k], a0);
        }
        return (a0.x + a0.y) + (a1.x + a1.y);
    }
    __device__ __forceinline__ void caxpy(float delta, const float2* __restrict__ u) {
        const float2 dl = make_float2(delta, delta);
#pragma unroll
        for (int k = 0; k < DP; k++) dv[k] = __ffma2_rn(dl, u[k], dv[k]);
    }
    // dv += delta * M^-1[d][:]; the pair holding the component the next row reads (`first`) is issued first
    __device__ __forceinline__ void axpy(int d, float delta, int first) {
        const float2 dl = make_float2(delta, delta);
        const int fp = (first >= 0 && first < D) ? (first >> 1) : -1;
        if (fp >= 0) dv[fp] = __ffma2_rn(dl, M[d][fp], dv[fp]);
#pragma unroll
        for (int k = 0; k < DP; k++)
            if (k != fp) dv[k] = __ffma2_rn(dl, M[d][k], dv[k]);
    }
};

// launch 3: thread = arm.  Environment.step (reference environment.py:453-485) from the free velocity on:
// motor / limit rows, projected Gauss-Seidel, integration, get_state, get_reward, is_terminal_state.
template <int D, int NLMAX, bool PACKED, bool EXACT, bool GJK>
__global__ void __launch_bounds__(kTpb)
sim_solve_kernel(const __grid_constant__ ModelDev M, SimArrays S, const __grid_constant__ StepCfgDev cfg, int mode,
                 const float* __restrict__ actions, const uint8_t* __restrict__ active, float* __restrict__ obs,
                 float* __restrict__ reward, uint8_t* __restrict__ done, uint8_t* __restrict__ valid_out) {
    extern __shared__ float smem[];
    const int tid = threadIdx.x;
    const int N = S.n_envs;
    const int env = blockIdx.x * kTpb + tid;
    const int so = 9 + 2 * M.n_obs;
    float* sm_obs = smem + tid * so;                    // [kTpb][so] staging for a coalesced obs write
    float* sm_lim = smem + kTpb * so + tid;             // [4][D][kTpb]: limit rows rhs_lo, app_lo, rhs_hi, app_hi
    bool run = env < N && (active == nullptr || active[env] != 0);
    // an env with pending reset sub-steps spends this launch on one of them (motors as begin_reset left
    // them) instead of an action step; it emits no transition (valid = 0)
    const int pending = run ? S.reset_left[env] : 0;
    if (mode == kModeResetOnly && pending == 0) run = false;
    const bool is_reset = pending > 0;
    bool staged = false;
    if (run) {
        const int ndof = EXACT ? D : M.ndof;       // EXACT: the instantiation matches the model, no per-row guards
        const size_t sN = (size_t)N;
        // ---- every global operand of the solve is requested up front, branch-free, so the whole batch costs one
        // memory round trip (the stores to the motor table below would otherwise fence the loads of the next dof) ----
        float q0[D], qs0[D], kp0[D], tp0[D], tv0[D], mi0[D], ac0[D];
#pragma unroll
        for (int d = 0; d < D; d++) {
            const int link = M.doflink[(EXACT || d < ndof) ? d : 0];
            const size_t at = (size_t)link * sN + env;
            q0[d] = S.q[at]; qs0[d] = S.qs[at];
            kp0[d] = S.kp[at]; tp0[d] = S.tpos[at]; tv0[d] = S.tvel[at]; mi0[d] = S.maximp[at];
            const int ai = cfg.act_index[link];
            ac0[d] = (ai >= 0 && actions != nullptr) ? actions[(size_t)env * cfg.n_act + ai] : 0.f;
        }
        PgsState<D, PACKED> P;
        P.load(S.minv, ndof, sN, env);

        // ---- constraint rows: joint limits first (created at import), then one motor row per dof ----
        float rhs[D], mx[D], jdi[D], app[D], csum[D];   // csum = app + rhs, kept current off the critical path
        unsigned mlo = 0u, mhi = 0u;
        const float inv_dt = M.inv_dt;
#pragma unroll
        for (int d = 0; d < D; d++) {
            rhs[d] = 0.f; mx[d] = 0.f; jdi[d] = 0.f; app[d] = 0.f; csum[d] = 0.f;
            if (EXACT || d < ndof) {
                const int link = M.doflink[d];
                const size_t at = (size_t)link * sN + env;
                const float q = q0[d], qs = qs0[d];
                // setJointMotorControl2: VELOCITY_CONTROL on involved joints (environment.py:464-469), then
                // POSITION_CONTROL target 0 on the fixed joints (:472-476); other joints keep their motor
                const int ai = cfg.act_index[link];
                const bool fixedj = (cfg.fixed_mask >> link) & 1u;
                float kp = kp0[d], tpos = tp0[d], tvel = tv0[d], maximp = mi0[d];
                if (!(is_reset || (!fixedj && ai < 0))) {
                    if (fixedj) {
                        kp = 0.1f; tpos = 0.f; tvel = 0.f; maximp = cfg.pos_maximp;
                    } else {
                        kp = 0.f; tpos = 0.f; tvel = ac0[d]; maximp = cfg.vel_maximp;
                    }
                    S.kp[at] = kp; S.tpos[at] = tpos; S.tvel[at] = tvel; S.maximp[at] = maximp;
                }
                const float j = 1.f / P.diag(d);
                jdi[d] = j;
                mx[d] = maximp;
                // btMultiBodyJointMotor: rhs = kp (q_des - q)/dt + qs + kd (qd_des - qs), kd = 1, erp = 1
                // (the row's right-hand side is rhs - qs; formed directly to avoid the cancellation)
                rhs[d] = fmaf(kp * (tpos - q), inv_dt, tvel - qs) * j;
                csum[d] = rhs[d];
                if (M.has_limit[link]) {
                    const float pen0 = q - M.lower[link], pen1 = M.upper[link] - q;
                    if (!(pen0 > 0.f)) {
                        mlo |= 1u << d;
                        sm_lim[(0 * D + d) * kTpb] = (-pen0 * M.erp * inv_dt - qs) * j;
                        sm_lim[(1 * D + d) * kTpb] = 0.f;
                    }
                    if (!(pen1 > 0.f)) {
                        mhi |= 1u << d;
                        sm_lim[(2 * D + d) * kTpb] = (-pen1 * M.erp * inv_dt + qs) * j;
                        sm_lim[(3 * D + d) * kTpb] = 0.f;
                    }
                }
            }
        }

        // ---- contact rows against the obstacle sphere / the target cube (environment.py:252-255 loads both
        // as collidable fixed bodies).  The rows come from the collision phase (sim_contacts_kernel); they live in local memory
        // here (dynamic contact index) ----
        int nc = 0;
        constexpr int kCP = kMaxDof / 2;
        __align__(16) float2 cJ2[kMaxContacts * kCP], cU2[kMaxContacts * kCP];        // pairs (dof 2k, 2k + 1) like dv
        float* cJ = reinterpret_cast<float*>(cJ2);
        float* cU = reinterpret_cast<float*>(cU2);
        float crhs[kMaxContacts], cjdi[kMaxContacts], capp[kMaxContacts], cjmj[kMaxContacts];
        const int nflag = cfg.contact_thr > 0.f ? S.near[env] : 0;
        if (nflag > 0 && !(cfg.contact_dbg & 2)) {
            float cd[kMaxContacts];
            {                       // rows left by the collision phase (sim_contacts_kernel, same pose)
                nc = min(nflag, kMaxContacts);
                const float* row = S.crow + (size_t)env * kMaxContacts * kContactRec;
                for (int c = 0; c < nc; c++) {
#pragma unroll
                    for (int a = 0; a < kMaxDof; a++) cJ[c * kMaxDof + a] = a < D ? row[c * kContactRec + a] : 0.f;
                    cd[c] = row[c * kContactRec + kMaxDof];
                }
            }
            if (cfg.contact_dbg & 1) nc = 0;          // diagnostics: pay for the collision phase only
            for (int c = 0; c < nc; c++) {
                float jmj = 0.f, rel = 0.f;
#pragma unroll
                for (int a = 0; a < kMaxDof; a++) {
                    float u = 0.f;
                    if (a < D) {
#pragma unroll
                        for (int b = 0; b < D; b++) u = fmaf(P.mat(a, b), cJ[c * kMaxDof + b], u);     // (M^-1 J^T)_a
                        jmj = fmaf(cJ[c * kMaxDof + a], u, jmj);
                        rel = fmaf(cJ[c * kMaxDof + a], qs0[a], rel);
                    }
                    cU[c * kMaxDof + a] = u;
                }
                const float jd = jmj > 1e-12f ? 1.f / jmj : 0.f;
                const float d = cd[c];
                // separated: the row only acts if the gap would close within the step; penetrating: erp pushes it out
                const float pos_err = d > 0.f ? 0.f : -d * M.erp * inv_dt;
                const float vel_err = -rel - (d > 0.f ? d * inv_dt : 0.f);
                cjdi[c] = jd;
                cjmj[c] = jmj;
                crhs[c] = (pos_err + vel_err) * jd;
                capp[c] = 0.f;
            }
        }

        // ---- projected Gauss-Seidel, M^-1 in registers; sweep direction alternates like Bullet's ----
        const float limit_hi = M.limit_max_imp, thresh = M.resid_thresh;
        const int iters = M.iters;
        int it = 0;
        float resid = 0.f;
        // Row update (btMultiBodyConstraintSolver::resolveSingleConstraintRowGeneric), branch-free and with the
        // shortest dependent chain from dv[d] to the dv the next row reads:
        //   sum = (applied + rhs) - dv[d] / A_dd   [Bullet: applied + (rhs - dv[d] / A_dd); differs by one rounding]
        //   applied' = clamp(sum) ; delta = applied' - applied ; dv += delta * M^-1[d][:]
        auto motor_row = [&](const int d, const int next) {
            const float sum = fmaf(-P.get(d), jdi[d], csum[d]);
            const float napp = fminf(fmaxf(sum, -mx[d]), mx[d]);
            const float delta = napp - app[d];
            P.axpy(d, delta, next);
            app[d] = napp;
            csum[d] = napp + rhs[d];
            const float dvel = delta * P.diag(d);
            resid = fmaxf(resid, dvel * dvel);
        };
        auto limit_row = [&](const int d, const int side) {
            const float sgn = side ? -1.f : 1.f;
            float* r = sm_lim + (2 * side * D + d) * kTpb;
            const float a0 = r[D * kTpb];
            const float sum = fmaf(-(sgn * P.get(d)), jdi[d], a0 + r[0]);
            const float napp = fminf(fmaxf(sum, 0.f), limit_hi);
            const float delta = napp - a0;
            r[D * kTpb] = napp;
            P.axpy(d, sgn * delta, -1);
            const float dvel = delta * P.diag(d);
            resid = fmaxf(resid, dvel * dvel);
        };
        for (it = 0; it < iters; it++) {
            resid = 0.f;
            if (it & 1) {
                if (mlo | mhi) {
#pragma unroll
                    for (int d = 0; d < D; d++) {
                        if ((mlo >> d) & 1u) limit_row(d, 0);
                        if ((mhi >> d) & 1u) limit_row(d, 1);
                    }
                }
#pragma unroll
                for (int d = 0; d < D; d++)
                    if (EXACT || d < ndof) motor_row(d, d + 1);
            } else {
#pragma unroll
                for (int d = D - 1; d >= 0; d--)
                    if (EXACT || d < ndof) motor_row(d, d - 1);
                if (mlo | mhi) {
#pragma unroll
                    for (int d = D - 1; d >= 0; d--) {
                        if ((mhi >> d) & 1u) limit_row(d, 1);
                        if ((mlo >> d) & 1u) limit_row(d, 0);
                    }
                }
            }
            if (nc > 0) {       // normal contact rows, after the non-contact rows of the sweep.  Unrolled over the (at most
                                // kMaxContacts) rows so that every local-memory address is static: the loads of row c + 1
                                // are issued while row c's dependent chain runs
#pragma unroll
                for (int c = 0; c < kMaxContacts; c++) {
                    if (c < nc && cjdi[c] != 0.f) {
                        const float jdv = P.cdot(cJ2 + c * kCP);
                        const float sum = fmaxf(fmaf(-jdv, cjdi[c], capp[c] + crhs[c]), 0.f);      // applied impulse >= 0
                        const float delta = sum - capp[c];
                        capp[c] = sum;
                        P.caxpy(delta, cU2 + c * kCP);
                        const float dvel = delta * cjmj[c];
                        resid = fmaxf(resid, dvel * dvel);
                    }
                }
            }
            if (!(resid > thresh) || it >= iters - 1) { it++; break; }
        }

        // ---- velocity and position update (stepPositionsMultiDof); q and the free velocity are re-read (one batch
        // of independent L2 hits) rather than held in 24 registers across the sweep ----
        const float max_vel = M.max_vel, dt = M.dt;
        float qe[D], qse[D];
#pragma unroll
        for (int d = 0; d < D; d++) {
            const size_t at = (size_t)M.doflink[(EXACT || d < ndof) ? d : 0] * sN + env;
            qe[d] = S.q[at];
            qse[d] = S.qs[at];
        }
#pragma unroll
        for (int d = 0; d < D; d++) {
            if (EXACT || d < ndof) {
                const size_t at = (size_t)M.doflink[d] * sN + env;
                const float x = fminf(fmaxf(qse[d] + P.get(d), -max_vel), max_vel);
                S.qd[at] = x;
                S.q[at] = fmaf(dt, x, qe[d]);
            }
        }
        const int nl = M.nl;
        for (int i = 0; i < nl; i++)
            if (M.dofidx[i] < 0) S.qd[(size_t)i * sN + env] = 0.f;

        // ---- state / reward / done from the post-step configuration ----
        const V3 tg = ld3(S.target + 3 * (size_t)env), ob = ld3(S.obstacle + 3 * (size_t)env);
        // the observed joint states are requested before the kinematics, so their round trip hides behind it
        const int n = M.n_obs;
        float oq[kMaxDof], oqd[kMaxDof];
#pragma unroll
        for (int i = 0; i < kMaxDof; i++) {
            oq[i] = (i < n && obs != nullptr) ? S.q[(size_t)i * sN + env] : 0.f;
            oqd[i] = (i < n && obs != nullptr) ? S.qd[(size_t)i * sN + env] : 0.f;
        }
        const ObsOut o = fk_and_distances<NLMAX, GJK>(M, S.q + env, N, ob, tg, cfg.obstacle_thr, !is_reset, nullptr, 0);
        if (obs != nullptr) {
#pragma unroll
            for (int i = 0; i < kMaxDof; i++)
                if (i < n) { sm_obs[i] = oq[i]; sm_obs[n + i] = oqd[i]; }
            for (int i = kMaxDof; i < n; i++) {
                sm_obs[i] = S.q[(size_t)i * sN + env];
                sm_obs[n + i] = S.qd[(size_t)i * sN + env];
            }
            float* t = sm_obs + 2 * n;
            t[0] = o.ee_pos.x; t[1] = o.ee_pos.y; t[2] = o.ee_pos.z;
            t[3] = tg.x; t[4] = tg.y; t[5] = tg.z;
            t[6] = ob.x; t[7] = ob.y; t[8] = ob.z;
            staged = true;
        }
        S.iters[env] = it;
        if (is_reset) {
            S.reset_left[env] = pending - 1;
            if (mode == kModeStep) {
                reward[env] = 0.f;
                done[env] = 0;
                if (valid_out != nullptr) valid_out[env] = 0;
            }
        } else {
            const bool goal = o.ee_target < cfg.target_thr;
            // get_reward (environment.py:364-371): goal first, then collision, else -(d - threshold)
            reward[env] = goal ? 250.f : (o.hit ? -1000.f : -(o.ee_target - cfg.target_thr));
            done[env] = (o.hit || goal) ? 1 : 0;         // is_terminal_state (environment.py:326-333)
            if (valid_out != nullptr) valid_out[env] = 1;
        }
    }
    // obs rows of a block are contiguous: write them coalesced when every arm of the block produced one
    if (obs == nullptr) return;
    const int nvalid = min(kTpb, N - blockIdx.x * kTpb);
    const unsigned all = __ballot_sync(0xffffffffu, staged || tid >= nvalid);
    __syncwarp();
    if (all == 0xffffffffu) {
        float* dst = obs + (size_t)blockIdx.x * kTpb * so;
        for (int k = tid; k < nvalid * so; k += kTpb) dst[k] = smem[k];
    } else if (staged) {
        float* dst = obs + (size_t)env * so;
        for (int k = 0; k < so; k++) dst[k] = sm_obs[k];
    }
}

// get_state + the distances behind get_reward / is_terminal_state, without stepping
template <int NLMAX, bool GJK>
__global__ void __launch_bounds__(kTpb)
sim_observe_kernel(const __grid_constant__ ModelDev M, SimArrays S, float* __restrict__ obs,
                   float* __restrict__ link_obst, float* __restrict__ ee_target) {
    const int env = blockIdx.x * kTpb + threadIdx.x;
    const int N = S.n_envs;
    if (env >= N) return;
    const V3 tg = ld3(S.target + 3 * (size_t)env), ob = ld3(S.obstacle + 3 * (size_t)env);
    const ObsOut o = fk_and_distances<NLMAX, GJK>(M, S.q + env, N, ob, tg, 0.f, true,
                                             link_obst ? link_obst + (size_t)env * M.nl : nullptr, 1);
    if (obs != nullptr) {
        const int n = M.n_obs;
        float* t = obs + (size_t)env * (9 + 2 * n);
        for (int i = 0; i < n; i++) {
            t[i] = S.q[(size_t)i * N + env];
            t[n + i] = S.qd[(size_t)i * N + env];
        }
        t += 2 * n;
        t[0] = o.ee_pos.x; t[1] = o.ee_pos.y; t[2] = o.ee_pos.z;
        t[3] = tg.x; t[4] = tg.y; t[5] = tg.z;
        t[6] = ob.x; t[7] = ob.y; t[8] = ob.z;
    }
    if (ee_target != nullptr) ee_target[env] = o.ee_target;
}

// get_manipulator_collisions_with_itself (environment.py:394-412, collision_detector.py:63-98): closest distance
// between every pair of non-adjacent links = min over their shape pairs of GJK(core A, core B) - radius A - radius B
// (sphere core = point, capsule core = segment, hull radius = margin); thread = arm.  Not on the step path.
template <int NLMAX>
__global__ void __launch_bounds__(kTpb)
sim_self_distances_kernel(const __grid_constant__ ModelDev M, SimArrays S, float* __restrict__ out) {
    const int env = blockIdx.x * kTpb + threadIdx.x;
    const int N = S.n_envs, nl = M.nl;
    if (env >= N) return;
    float Rw[NLMAX * 9], pw[NLMAX * 3];
    for (int i = 0; i < nl; i++) {
        const int par = M.parent[i];
        M3 Rl;
        V3 pl, dd;
        joint_transform(M, i, S.q[(size_t)i * N + env], Rl, pl, dd);
        M3 R = Rl;
        V3 p = pl;
        if (par >= 0) {
            M3 Rp;
#pragma unroll
            for (int k = 0; k < 9; k++) Rp.m[k] = Rw[par * 9 + k];
            p = ld3(pw + par * 3) + mul(Rp, pl);
            R = mul(Rp, Rl);
        }
#pragma unroll
        for (int k = 0; k < 9; k++) Rw[i * 9 + k] = R.m[k];
        st3(pw + i * 3, p);
    }
    float* o = out + (size_t)env * nl * nl;
    for (int k = 0; k < nl * nl; k++) o[k] = 10.f;
    auto posed = [&](int s, GjkShape& A) -> float {      // returns the radius around the core
        const int l = M.s_link[s], t = M.s_type[s];
        M3 Rl, sR;
#pragma unroll
        for (int k = 0; k < 9; k++) { Rl.m[k] = Rw[l * 9 + k]; sR.m[k] = M.s_R[s][k]; }
        A.R = mul(Rl, sR);
        A.p = ld3(pw + l * 3) + mul(Rl, v3(M.s_p[s][0], M.s_p[s][1], M.s_p[s][2]));
        A.verts = nullptr;
        A.nv = 0;
        A.half = v3(0.f, 0.f, 0.f);
        if (t == RLOA_SHAPE_SPHERE) return M.s_dim[s][0];
        if (t == RLOA_SHAPE_CAPSULE) { A.half.z = M.s_dim[s][1]; return M.s_dim[s][0]; }
        if (t == RLOA_SHAPE_BOX) { A.half = v3(M.s_dim[s][0], M.s_dim[s][1], M.s_dim[s][2]); return 0.f; }
        A.verts = M.verts + M.s_v0[s];
        A.nv = M.s_vn[s];
        return M.s_dim[s][0];
    };
    const int ns = M.ns;
    for (int sa = 0; sa < ns; sa++) {
        GjkShape A;
        const float ra = posed(sa, A);
        const int i = M.s_link[sa];
        for (int sb = sa + 1; sb < ns; sb++) {
            const int j = M.s_link[sb];
            if (i == j || i == j + 1 || j == i + 1) continue;
            GjkShape B;
            const float rb = posed(sb, B);
            const float d = gjk_distance_pair(A, B) - ra - rb;
            if (d < o[i * nl + j]) { o[i * nl + j] = d; o[j * nl + i] = d; }
        }
    }
}

// lock-step asynchronous Environment.reset: arm the POSITION_CONTROL motors of the masked envs and let
// the next n_substeps step launches run their reset sub-steps (environment.py:295-301)
__global__ void sim_begin_reset_kernel(SimArrays S, const uint8_t* __restrict__ mask,
                                       const float* __restrict__ init_targets, int n_init, int nsub, float pos_maximp) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= S.n_envs) return;
    if (mask != nullptr && mask[env] == 0) return;
    const size_t N = (size_t)S.n_envs;
    for (int j = 0; j < n_init; j++) {
        const size_t at = (size_t)j * N + env;
        S.kp[at] = 0.1f; S.tpos[at] = init_targets[(size_t)env * n_init + j]; S.tvel[at] = 0.f; S.maximp[at] = pos_maximp;
    }
    S.reset_left[env] = nsub;
}

// the same with the start pose drawn on the device: target_j = pos_j + var_j * U(-1, 1), Philox keyed by
// (seed, tick, env, j) — random.uniform(pos - var, pos + var) of environment.py:288 for every masked env
__global__ void sim_begin_reset_random_kernel(SimArrays S, const uint8_t* __restrict__ mask, const float* __restrict__ pos,
                                              const float* __restrict__ var, int n_init, int nsub, float pos_maximp,
                                              unsigned long long seed, const unsigned long long* __restrict__ tick) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= S.n_envs) return;
    if (mask != nullptr && mask[env] == 0) return;
    const size_t N = (size_t)S.n_envs;
    const unsigned long long t = tick != nullptr ? *tick : 0ull;
    for (int j0 = 0; j0 < n_init; j0 += 4) {
        uint32_t c[4] = {(uint32_t)env, (uint32_t)(j0 >> 2), (uint32_t)t, (uint32_t)(t >> 32)};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x52455345u);
        for (int k = 0; k < 4 && j0 + k < n_init; k++) {
            const int j = j0 + k;
            const size_t at = (size_t)j * N + env;
            const float v = var != nullptr ? var[j] : 0.f;
            S.kp[at] = 0.1f;
            S.tpos[at] = fmaf(v, 2.f * u01(c[k]) - 1.f, pos != nullptr ? pos[j] : 0.f);
            S.tvel[at] = 0.f;
            S.maximp[at] = pos_maximp;
        }
    }
    S.reset_left[env] = nsub;
}

// [N][nl] (C ABI layout) <-> [nl][N] (device layout)
__global__ void sim_scatter_kernel(int n_envs, int nl, const float* __restrict__ src, float* __restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_envs * nl) return;
    const int env = (int)(i / nl), j = (int)(i - (size_t)env * nl);
    dst[(size_t)j * n_envs + env] = src[i];
}
__global__ void sim_gather_kernel(int n_envs, int nl, const float* __restrict__ src, float* __restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_envs * nl) return;
    const int env = (int)(i / nl), j = (int)(i - (size_t)env * nl);
    dst[i] = src[(size_t)j * n_envs + env];
}

__global__ void sim_clear_kernel(SimArrays S) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)S.n_envs * S.nl;
    if (i < n) {
        S.q[i] = 0.f; S.qd[i] = 0.f;
        S.kp[i] = 0.f; S.tpos[i] = 0.f; S.tvel[i] = 0.f; S.maximp[i] = 1.f;   // load-time default motor
    }
    if (i < (size_t)S.n_envs) { S.iters[i] = 0; S.reset_left[i] = 0; }
}

// Episode bookkeeping of the vectorised rollout (naf_algorithm.py:263-277, rl_framework.py:343-354)
__global__ void episode_update_kernel(int n, int frames, const float* __restrict__ reward,
                                      const uint8_t* __restrict__ done, const uint8_t* __restrict__ active,
                                      float* __restrict__ score, int* __restrict__ frame, uint8_t* __restrict__ reset_mask,
                                      float* __restrict__ log_score, int* __restrict__ log_frame,
                                      float* __restrict__ log_last, int* __restrict__ log_env, int log_cap,
                                      int* __restrict__ log_count, unsigned long long* __restrict__ transitions,
                                      unsigned long long* __restrict__ tick) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e == 0 && tick != nullptr) *tick += 1ull;
    const bool live = e < n && (active == nullptr || active[e] != 0);
    const int nlive = __syncthreads_count(live);
    if (threadIdx.x == 0 && transitions != nullptr && nlive > 0) atomicAdd(transitions, (unsigned long long)nlive);
    if (e >= n) return;
    if (!live) {
        reset_mask[e] = 0;
        return;
    }
    const float sc = score[e] + reward[e];
    const int fr = frame[e] + 1;
    const bool fin = done[e] != 0 || fr >= frames;
    if (fin) {
        const int slot = atomicAdd(log_count, 1);
        if (slot < log_cap) {
            log_score[slot] = sc;
            log_frame[slot] = fr;
            log_last[slot] = reward[e];
            log_env[slot] = e;
        }
        score[e] = 0.f;
        frame[e] = 0;
        reset_mask[e] = 1;
    } else {
        score[e] = sc;
        frame[e] = fr;
        reset_mask[e] = 0;
    }
}

// Episode bookkeeping and the lock-step reset of the finished envs in ONE launch (the two are per-env and
// independent across envs).  The start-pose draw is keyed by the loop counter read at kernel entry; the counter is
// incremented by whichever block finishes last, so every block sees the same value whatever the scheduling order.
__global__ void episode_update_reset_kernel(int n, int frames, const float* __restrict__ reward,
                                            const uint8_t* __restrict__ done, const uint8_t* __restrict__ active,
                                            float* __restrict__ score, int* __restrict__ frame, uint8_t* __restrict__ reset_mask,
                                            float* __restrict__ log_score, int* __restrict__ log_frame,
                                            float* __restrict__ log_last, int* __restrict__ log_env, int log_cap,
                                            int* __restrict__ log_count, unsigned long long* __restrict__ transitions,
                                            unsigned long long* __restrict__ tick, unsigned long long* __restrict__ tick_next,
                                            SimArrays S, const float* __restrict__ pos,
                                            const float* __restrict__ var, int n_init, int nsub, float pos_maximp,
                                            unsigned long long seed, unsigned* __restrict__ ticket) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long t = tick != nullptr ? *tick + 1ull : 0ull;     // the value begin_reset would have seen
    const bool live = e < n && (active == nullptr || active[e] != 0);
    const int nlive = __syncthreads_count(live);
    if (threadIdx.x == 0 && transitions != nullptr && nlive > 0) atomicAdd(transitions, (unsigned long long)nlive);
    bool fin = false;
    if (e < n) {
        if (live) {
            const float sc = score[e] + reward[e];
            const int fr = frame[e] + 1;
            fin = done[e] != 0 || fr >= frames;
            if (fin) {
                const int slot = atomicAdd(log_count, 1);
                if (slot < log_cap) {
                    log_score[slot] = sc;
                    log_frame[slot] = fr;
                    log_last[slot] = reward[e];
                    log_env[slot] = e;
                }
                score[e] = 0.f;
                frame[e] = 0;
            } else {
                score[e] = sc;
                frame[e] = fr;
            }
        }
        reset_mask[e] = fin ? 1 : 0;
        if (fin) {        // sim_begin_reset_random_kernel for this env
            const size_t N = (size_t)S.n_envs;
            for (int j0 = 0; j0 < n_init; j0 += 4) {
                uint32_t c[4] = {(uint32_t)e, (uint32_t)(j0 >> 2), (uint32_t)t, (uint32_t)(t >> 32)};
                philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x52455345u);
                for (int k = 0; k < 4 && j0 + k < n_init; k++) {
                    const int j = j0 + k;
                    const size_t at = (size_t)j * N + e;
                    const float v = var != nullptr ? var[j] : 0.f;
                    S.kp[at] = 0.1f;
                    S.tpos[at] = fmaf(v, 2.f * u01(c[k]) - 1.f, pos != nullptr ? pos[j] : 0.f);
                    S.tvel[at] = 0.f;
                    S.maximp[at] = pos_maximp;
                }
            }
            S.reset_left[e] = nsub;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && tick != nullptr) {
        if (tick_next != nullptr && tick_next != tick) {
            if (blockIdx.x == 0) *tick_next = t;         // ping-pong counter: *tick is not written, readers may run beside
        } else if (atomicAdd(ticket, 1u) == gridDim.x - 1u) {
            *ticket = 0u;
            *tick += 1ull;
        }
    }
}

}  // namespace rloa

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace rloa;

struct rloa_model {
    ModelDev host;
    int device = 0;
    float4* verts = nullptr;     // device copy of the hull vertices (host.verts points here)
    ~rloa_model() {
        if (verts) cudaFree(verts);
    }
};

struct rloa_sim {
    const rloa_model* model = nullptr;
    SimArrays a{};
    float* block = nullptr;
    unsigned* ticket = nullptr;      // last-block-done counter of rloa_episode_update_reset (zero at rest)
    int device = 0;
    // software pipelining of the step (rloa_sim_prepare): the action-independent half (dynamics + M^-1) of the NEXT
    // step runs on this side stream while the caller's stream does something else
    cudaStream_t side = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    bool prepared = false;           // F / qs / M^-1 hold the current (q, qd): rloa_sim_step launches the solve only
    bool join_pending = false;       // the caller's stream has not yet waited for the side stream
    float contact_thr = 0.f;         // contact rows of rloa_sim_prepare / rloa_sim_reset (rloa_sim_set_contacts); steps carry theirs in the config
    float prepared_thr = -1.f;       // threshold the prepared collision phase ran with (-1: none)
};

// the caller's stream waits for an outstanding rloa_sim_prepare; with invalidate the prepared half is dropped
static int sim_sync_prepared(rloa_sim* s, cudaStream_t st, bool invalidate) {
    if (s->join_pending) {
        RLOA_CUDA(cudaStreamWaitEvent(st, s->join_ev, 0));
        s->join_pending = false;
    }
    if (invalidate) s->prepared = false;
    return RLOA_OK;
}

static void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j];
            C[3 * i + j] = s;
        }
}

extern "C" int rloa_model_create(const rloa_model_desc* d, rloa_model** out) {
    RLOA_REQUIRE(d != nullptr && out != nullptr, "rloa_model_create: null argument");
    RLOA_REQUIRE(d->nl >= 1 && d->nl <= kMaxLinks, "rloa_model_create: 1 <= nl <= 32 links supported");
    RLOA_REQUIRE(d->ns >= 0 && d->ns <= kMaxShapes, "rloa_model_create: at most 32 collision shapes supported");
    RLOA_REQUIRE(d->n_verts >= 0 && d->n_verts <= RLOA_MAX_HULL_VERTS, "rloa_model_create: too many convex-hull vertices");
    RLOA_REQUIRE(d->n_verts == 0 || (d->s_vert_first && d->s_vert_count && d->verts), "rloa_model_create: hull vertex arrays missing");
    RLOA_REQUIRE(d->ee_link >= 0 && d->ee_link < d->nl, "rloa_model_create: endeffector index out of range");
    RLOA_REQUIRE(d->n_obs_joints >= 0 && d->n_obs_joints <= d->nl, "rloa_model_create: n_obs_joints out of range");
    RLOA_REQUIRE(d->dt > 0 && d->iters >= 1, "rloa_model_create: dt > 0 and iters >= 1 required");
    rloa_model* m = new (std::nothrow) rloa_model();
    RLOA_REQUIRE(m != nullptr, "rloa_model_create: out of host memory");
    ModelDev& h = m->host;
    std::memset(&h, 0, sizeof(h));
    const int nl = d->nl;
    h.nl = nl; h.ns = d->ns; h.ee_link = d->ee_link; h.n_obs = d->n_obs_joints; h.iters = d->iters;
    h.dt = (float)d->dt; h.inv_dt = (float)(1.0 / d->dt);
    h.lin_damp = (float)d->lin_damp; h.ang_damp = (float)d->ang_damp;
    h.resid_thresh = (float)d->resid_thresh; h.erp = (float)d->erp; h.max_vel = (float)d->max_vel;
    h.limit_max_imp = (float)d->limit_max_impulse;
    for (int k = 0; k < 3; k++) { h.gravity[k] = (float)d->gravity[k]; h.target_half[k] = (float)d->target_half[k]; }
    h.obstacle_radius = (float)d->obstacle_radius;
    int ndof = 0, nslots = 0;
    std::vector<int> slot(nl, -1);
    for (int i = 0; i < nl; i++) {
        const int p = d->parent[i];
        if (!(p >= -1 && p < i)) { delete m; return fail(RLOA_ERR_INVALID, "rloa_model_create: links must be numbered depth first (parent < child)"); }
        h.parent[i] = p;
        h.jtype[i] = d->jtype[i];
        h.has_limit[i] = d->has_limit[i];
        h.dofidx[i] = -1;
        if (d->jtype[i] != RLOA_JOINT_FIXED) {
            if (ndof >= kMaxDof) { delete m; return fail(RLOA_ERR_INVALID, "rloa_model_create: at most 16 movable joints supported"); }
            if (!(d->mass[i] > 0)) { delete m; return fail(RLOA_ERR_INVALID, "rloa_model_create: movable link without mass"); }
            h.dofidx[i] = ndof;
            h.doflink[ndof++] = i;
        }
        // parent COM frame <- child COM frame at q = 0; the fixed base pose is folded into root links
        double E0T[9], ee[3];
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) E0T[3 * a + b] = d->E0[9 * i + 3 * b + a];
        for (int a = 0; a < 3; a++) ee[a] = d->e[3 * i + a];
        if (p < 0) {
            double T[9], e2[3];
            mat3_mul(d->base_R, E0T, T);
            for (int a = 0; a < 3; a++)
                e2[a] = d->base_p[a] + d->base_R[3 * a] * ee[0] + d->base_R[3 * a + 1] * ee[1] + d->base_R[3 * a + 2] * ee[2];
            std::memcpy(E0T, T, sizeof(T));
            std::memcpy(ee, e2, sizeof(e2));
        }
        for (int k = 0; k < 9; k++) h.E0T[i][k] = (float)E0T[k];
        for (int k = 0; k < 3; k++) {
            h.e[i][k] = (float)ee[k];
            h.d[i][k] = (float)d->d[3 * i + k];
            h.axis[i][k] = (float)d->axis[3 * i + k];
            h.inertia[i][k] = (float)d->inertia[3 * i + k];
        }
        h.mass[i] = (float)d->mass[i];
        h.damping[i] = (float)d->damping[i];
        h.lower[i] = (float)d->lower[i];
        h.upper[i] = (float)d->upper[i];
    }
    if (ndof < 1) { delete m; return fail(RLOA_ERR_INVALID, "rloa_model_create: the model has no movable joint"); }
    h.ndof = ndof;
    // where the serial sweeps find a link's parent value: the previous link (running register), or a slot
    // kept for links whose children are not numbered right after them
    for (int i = 0; i < nl; i++) {
        const int p = h.parent[i];
        h.bwsrc[i] = 0;
        if (p < 0) { h.fwsrc[i] = 0; h.bwdst[i] = -1; }
        else if (p == i - 1) { h.fwsrc[i] = 1; h.bwdst[i] = 0; }
        else {
            if (slot[p] < 0) {
                if (nslots >= kMaxSlots) { delete m; return fail(RLOA_ERR_INVALID, "rloa_model_create: too many branching links (max 3)"); }
                slot[p] = nslots++;
            }
            h.fwsrc[i] = 2 + slot[p];
            h.bwdst[i] = 1 + slot[p];
        }
    }
    h.nslots = nslots;
    for (int i = 0; i < nl; i++) {
        h.fwsave[i] = slot[i];
        if (i + 1 < nl && h.parent[i + 1] == i) h.bwsrc[i] |= 1;
        if (slot[i] >= 0) h.bwsrc[i] |= (slot[i] + 1) << 1;
    }
    for (int s = 0; s < d->ns; s++) {
        if (!(d->s_link[s] >= 0 && d->s_link[s] < nl)) { delete m; return fail(RLOA_ERR_INVALID, "rloa_model_create: shape link out of range"); }
        h.s_link[s] = d->s_link[s];
        h.s_type[s] = d->s_type[s];
        for (int k = 0; k < 9; k++) h.s_R[s][k] = (float)d->s_R[9 * s + k];
        for (int k = 0; k < 3; k++) { h.s_p[s][k] = (float)d->s_p[3 * s + k]; h.s_dim[s][k] = (float)d->s_dim[3 * s + k]; }
        const int t = d->s_type[s];
        if (!(t >= RLOA_SHAPE_SPHERE && t <= RLOA_SHAPE_HULL)) { delete m; return fail(RLOA_ERR_INVALID, "rloa_model_create: unknown shape type"); }
        // bounding sphere: centre c (shape frame) and radius, then the centre moved to the link frame
        double c[3] = {0, 0, 0}, rb = 0;
        if (t == RLOA_SHAPE_SPHERE) rb = d->s_dim[3 * s];
        else if (t == RLOA_SHAPE_CAPSULE) rb = d->s_dim[3 * s] + d->s_dim[3 * s + 1];
        else if (t == RLOA_SHAPE_BOX) rb = std::sqrt(d->s_dim[3 * s] * d->s_dim[3 * s] + d->s_dim[3 * s + 1] * d->s_dim[3 * s + 1] + d->s_dim[3 * s + 2] * d->s_dim[3 * s + 2]);
        else {
            const int v0 = d->n_verts ? d->s_vert_first[s] : 0, vn = d->n_verts ? d->s_vert_count[s] : 0;
            if (!(vn >= 1 && v0 >= 0 && v0 + vn <= d->n_verts)) { delete m; return fail(RLOA_ERR_INVALID, "rloa_model_create: hull shape without vertices / vertex range out of bounds"); }
            h.s_v0[s] = v0;
            h.s_vn[s] = vn;
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, r2 = 0;
            for (int i = v0; i < v0 + vn; i++)
                for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], d->verts[3 * i + k]); hi[k] = std::max(hi[k], d->verts[3 * i + k]); }
            for (int k = 0; k < 3; k++) c[k] = 0.5 * (lo[k] + hi[k]);
            for (int i = v0; i < v0 + vn; i++) {
                double q = 0;
                for (int k = 0; k < 3; k++) q += (d->verts[3 * i + k] - c[k]) * (d->verts[3 * i + k] - c[k]);
                r2 = std::max(r2, q);
            }
            rb = std::sqrt(r2) + d->s_dim[3 * s];         // + the collision margin
            h.use_gjk = 1;
        }
        for (int k = 0; k < 3; k++)
            h.s_bs[s][k] = (float)(d->s_p[3 * s + k] + d->s_R[9 * s + 3 * k] * c[0] + d->s_R[9 * s + 3 * k + 1] * c[1] + d->s_R[9 * s + 3 * k + 2] * c[2]);
        h.s_bs[s][3] = (float)(rb * (1.0 + 1e-5) + 1e-6);
        if (t == RLOA_SHAPE_BOX && d->s_link[s] == d->ee_link) h.use_gjk = 1;   // box vs target cube needs the narrow phase
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { delete m; cudaGetLastError(); return fail(RLOA_ERR_NO_DEVICE, "rloa_model_create: no CUDA device visible"); }
    if (cudaGetDevice(&m->device) != cudaSuccess) { delete m; return fail(RLOA_ERR_CUDA, "rloa_model_create: cudaGetDevice failed"); }
    if (d->n_verts > 0) {
        std::vector<float4> hv((size_t)d->n_verts);
        for (int i = 0; i < d->n_verts; i++)
            hv[i] = make_float4((float)d->verts[3 * i], (float)d->verts[3 * i + 1], (float)d->verts[3 * i + 2], 0.f);
        if (cudaMalloc(&m->verts, hv.size() * sizeof(float4)) != cudaSuccess ||
            cudaMemcpy(m->verts, hv.data(), hv.size() * sizeof(float4), cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("rloa_model_create: uploading the hull vertices failed: %s", cudaGetErrorString(cudaGetLastError()));
            delete m;
            return RLOA_ERR_CUDA;
        }
        h.verts = m->verts;
    }
    *out = m;
    return RLOA_OK;
}

extern "C" void rloa_model_destroy(rloa_model* m) { delete m; }

extern "C" int rloa_sim_create(const rloa_model* m, int32_t n_envs, rloa_sim** out) {
    RLOA_REQUIRE(m != nullptr && out != nullptr, "rloa_sim_create: null argument");
    RLOA_REQUIRE(n_envs >= 1, "rloa_sim_create: n_envs >= 1 required");
    rloa_sim* s = new (std::nothrow) rloa_sim();
    RLOA_REQUIRE(s != nullptr, "rloa_sim_create: out of host memory");
    s->model = m;
    s->device = m->device;
    const int nl = m->host.nl, ndof = m->host.ndof;
    const size_t N = (size_t)n_envs, n = N * nl;
    s->a.n_envs = n_envs;
    s->a.nl = nl;
    s->a.ndof = ndof;
    // one allocation: q qd kp tpos tvel maximp qs | F | minv | target obstacle | iters reset_left near
    const size_t floats = 7 * n + n * kFRec + N * ndof * ndof + 6 * N + 3 * N;
    if (cudaMalloc(&s->block, floats * sizeof(float)) != cudaSuccess) {
        set_error("rloa_sim_create: cudaMalloc of %zu bytes failed: %s", floats * sizeof(float), cudaGetErrorString(cudaGetLastError()));
        delete s;
        return RLOA_ERR_CUDA;
    }
    float* b = s->block;
    s->a.q = b; s->a.qd = b + n; s->a.kp = b + 2 * n; s->a.tpos = b + 3 * n;
    s->a.tvel = b + 4 * n; s->a.maximp = b + 5 * n; s->a.qs = b + 6 * n;
    s->a.F = b + 7 * n;
    s->a.minv = s->a.F + n * kFRec;
    s->a.target = s->a.minv + N * ndof * ndof; s->a.obstacle = s->a.target + 3 * N;
    s->a.iters = reinterpret_cast<int*>(s->a.obstacle + 3 * N);
    s->a.reset_left = s->a.iters + n_envs;
    s->a.near = s->a.reset_left + n_envs;
    if (cudaMalloc(&s->a.crow, N * kMaxContacts * kContactRec * sizeof(float)) != cudaSuccess) {
        set_error("rloa_sim_create: cudaMalloc of the contact rows failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(s->block);
        delete s;
        return RLOA_ERR_CUDA;
    }
    if (cudaMalloc(&s->ticket, sizeof(unsigned)) == cudaSuccess) cudaMemset(s->ticket, 0, sizeof(unsigned));
    cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&s->fork_ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s->join_ev, cudaEventDisableTiming);
    cudaFuncSetAttribute(sim_minv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxLinks * kFRec * 32 * (int)sizeof(float));
    cudaMemset(s->block, 0, floats * sizeof(float));
    cudaMemset(s->a.near, 0, N * sizeof(int));
    sim_clear_kernel<<<(unsigned)((n + 255) / 256), 256>>>(s->a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        set_error("rloa_sim_create: initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(s->block);
        delete s;
        return RLOA_ERR_CUDA;
    }
    *out = s;
    return RLOA_OK;
}

extern "C" void rloa_sim_destroy(rloa_sim* s) {
    if (s == nullptr) return;
    if (s->block) cudaFree(s->block);
    if (s->a.crow) cudaFree(s->a.crow);
    if (s->ticket) cudaFree(s->ticket);
    if (s->side) cudaStreamDestroy(s->side);
    if (s->fork_ev) cudaEventDestroy(s->fork_ev);
    if (s->join_ev) cudaEventDestroy(s->join_ev);
    delete s;
}

extern "C" int rloa_sim_num_envs(const rloa_sim* s) { return s ? s->a.n_envs : RLOA_ERR_INVALID; }
extern "C" int rloa_sim_obs_size(const rloa_sim* s) { return s ? 9 + 2 * s->model->host.n_obs : RLOA_ERR_INVALID; }

extern "C" int rloa_sim_set_task(rloa_sim* s, const float* target, const float* obstacle, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_set_task: null sim");
    const size_t bytes = 3 * (size_t)s->a.n_envs * sizeof(float);
    if (target) RLOA_CUDA(cudaMemcpyAsync(s->a.target, target, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    if (obstacle) RLOA_CUDA(cudaMemcpyAsync(s->a.obstacle, obstacle, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return RLOA_OK;
}

// [N][nl] caller layout -> [nl][N] device layout (and back)
static int scatter(const rloa_sim* s, const float* src, float* dst, void* stream) {
    const size_t n = (size_t)s->a.n_envs * s->a.nl;
    sim_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(s->a.n_envs, s->a.nl, src, dst);
    RLOA_LAUNCHED();
    return RLOA_OK;
}
static int gather(const rloa_sim* s, const float* src, float* dst, void* stream) {
    const size_t n = (size_t)s->a.n_envs * s->a.nl;
    sim_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(s->a.n_envs, s->a.nl, src, dst);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

extern "C" int rloa_sim_set_state(rloa_sim* s, const float* q, const float* qd, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_set_state: null sim");
    int rc = sim_sync_prepared(s, as_stream(stream), true);
    if (rc != RLOA_OK) return rc;
    if (q && (rc = scatter(s, q, s->a.q, stream)) != RLOA_OK) return rc;
    if (qd && (rc = scatter(s, qd, s->a.qd, stream)) != RLOA_OK) return rc;
    return RLOA_OK;
}

extern "C" int rloa_sim_get_state(const rloa_sim* s, float* q, float* qd, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_get_state: null sim");
    int rc = RLOA_OK;
    if (q && (rc = gather(s, s->a.q, q, stream)) != RLOA_OK) return rc;
    if (qd && (rc = gather(s, s->a.qd, qd, stream)) != RLOA_OK) return rc;
    return RLOA_OK;
}

extern "C" int rloa_sim_set_motors(rloa_sim* s, const float* kp, const float* tpos, const float* tvel,
                                   const float* maximp, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_set_motors: null sim");
    int rc = RLOA_OK;
    if (kp && (rc = scatter(s, kp, s->a.kp, stream)) != RLOA_OK) return rc;
    if (tpos && (rc = scatter(s, tpos, s->a.tpos, stream)) != RLOA_OK) return rc;
    if (tvel && (rc = scatter(s, tvel, s->a.tvel, stream)) != RLOA_OK) return rc;
    if (maximp && (rc = scatter(s, maximp, s->a.maximp, stream)) != RLOA_OK) return rc;
    return RLOA_OK;
}

extern "C" int rloa_sim_clear(rloa_sim* s, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_clear: null sim");
    {
        const int rcp = sim_sync_prepared(s, as_stream(stream), true);
        if (rcp != RLOA_OK) return rcp;
    }
    const size_t n = (size_t)s->a.n_envs * s->a.nl;
    sim_clear_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(s->a);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

static int make_step_cfg(const rloa_sim* s, const rloa_step_config* c, StepCfgDev* out) {
    const ModelDev& h = s->model->host;
    RLOA_REQUIRE(c->n_act >= 0 && c->n_act <= kMaxLinks && c->n_fixed >= 0 && c->n_fixed <= kMaxLinks,
                 "rloa_sim_step: joint list too long");
    RLOA_REQUIRE(c->n_act == h.n_obs, "rloa_sim_step: n_act must equal the model's n_obs_joints");
    std::memset(out, 0, sizeof(*out));
    for (int i = 0; i < kMaxLinks; i++) out->act_index[i] = -1;
    out->fixed_mask = 0;
    for (int k = 0; k < c->n_act; k++) {
        const int j = c->act_joint[k];
        RLOA_REQUIRE(j >= 0 && j < h.nl, "rloa_sim_step: involved joint index out of range");
        out->act_index[j] = (signed char)k;
    }
    for (int k = 0; k < c->n_fixed; k++) {
        const int j = c->fixed_joint[k];
        RLOA_REQUIRE(j >= 0 && j < h.nl, "rloa_sim_step: fixed joint index out of range");
        out->fixed_mask |= 1u << j;
    }
    out->n_act = c->n_act;
    out->vel_maximp = c->max_force * h.dt;
    out->pos_maximp = 100000.f * h.dt;
    out->target_thr = c->target_threshold;
    out->obstacle_thr = c->obstacle_threshold;
    out->contact_thr = c->contact_threshold > 0.f ? c->contact_threshold : 0.f;
    {
        const char* e = getenv("RLOA_CONTACT_DEBUG");
        out->contact_dbg = e != nullptr ? atoi(e) : 0;
    }
    return RLOA_OK;
}

template <int NLMAX>
static int launch_solve(const rloa_sim* s, const StepCfgDev& c, int mode, const float* actions, const uint8_t* active,
                        float* obs, float* reward, uint8_t* done, uint8_t* valid, cudaStream_t st) {
    const ModelDev& h = s->model->host;
    const unsigned blocks = (unsigned)((s->a.n_envs + kTpb - 1) / kTpb);
    const int so = 9 + 2 * h.n_obs;
#define RLOA_SOLVE_G(D, EXACT, GJK)                                                                                \
    sim_solve_kernel<D, NLMAX, (D <= 12), EXACT, GJK><<<blocks, kTpb, (size_t)(kTpb * so + 4 * D * kTpb) * sizeof(float), st>>>( \
        h, s->a, c, mode, actions, active, obs, reward, done, valid)
#define RLOA_SOLVE(D, EXACT) RLOA_SOLVE_G(D, EXACT, false)
    if (h.use_gjk) {               // mesh-hull shapes or box end-effector shapes: the GJK narrow phase is compiled in
        if (h.ndof <= 8) RLOA_SOLVE_G(8, false, true);
        else if (h.ndof <= 12) RLOA_SOLVE_G(12, false, true);
        else RLOA_SOLVE_G(16, false, true);
    } else
    if (h.ndof == 12 && NLMAX == 16) RLOA_SOLVE(12, true);         // KUKA iiwa + gripper
    else if (h.ndof == 9 && NLMAX == 16) RLOA_SOLVE(9, true);      // Panda
    else if (h.ndof == 7 && NLMAX == 16) RLOA_SOLVE(7, true);      // bare 7-dof arm
    else if (h.ndof <= 4) RLOA_SOLVE(4, false);
    else if (h.ndof <= 8) RLOA_SOLVE(8, false);
    else if (h.ndof <= 10) RLOA_SOLVE(10, false);
    else if (h.ndof <= 12) RLOA_SOLVE(12, false);
    else RLOA_SOLVE(16, false);
#undef RLOA_SOLVE
#undef RLOA_SOLVE_G
    RLOA_LAUNCHED();
    return RLOA_OK;
}

// one stepSimulation for every env = three launches on the caller's stream
// the action-independent half of a stepSimulation: dynamics + M^-1 columns for the current (q, qd)
// the collision phase for the current pose (contact rows -> crow / near); thr <= 0: none, the solve sees no rows
static int launch_contacts(const rloa_sim* s, float thr, cudaStream_t st) {
    if (!(thr > 0.f)) return RLOA_OK;       // the solve ignores `near` without a threshold
    const ModelDev& h = s->model->host;
    const int tpb = s->a.n_envs <= 16384 ? kDynWideTpb : kTpb;
    const unsigned blocks = (unsigned)((s->a.n_envs + tpb - 1) / tpb);
    if (h.nl <= 16) sim_contacts_kernel<16><<<blocks, tpb, 0, st>>>(h, s->a, thr);
    else sim_contacts_kernel<32><<<blocks, tpb, 0, st>>>(h, s->a, thr);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

static int launch_dynamics(const rloa_sim* s, cudaStream_t st) {
    const ModelDev& h = s->model->host;
    const unsigned blocks = (unsigned)((s->a.n_envs + kTpb - 1) / kTpb);
    const int dtpb = s->a.n_envs <= 16384 ? kDynWideTpb : kTpb;
    const unsigned dblocks = (unsigned)((s->a.n_envs + dtpb - 1) / dtpb);
    if (h.nl <= 16) sim_dynamics_kernel<16><<<dblocks, dtpb, 0, st>>>(h, s->a);
    else sim_dynamics_kernel<32><<<dblocks, dtpb, 0, st>>>(h, s->a);
    RLOA_LAUNCHED();
    sim_minv_kernel<<<blocks, dim3(32, h.ndof), (size_t)h.nl * kFRec * 32 * sizeof(float), st>>>(h, s->a);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

// one stepSimulation for every env = three launches on the caller's stream
static int launch_substep(const rloa_sim* s, const StepCfgDev& c, int mode, const float* actions, const uint8_t* active,
                          float* obs, float* reward, uint8_t* done, uint8_t* valid, cudaStream_t st, bool skip_dynamics = false,
                          float prepared_thr = -1.f) {
    const ModelDev& h = s->model->host;
    if (!skip_dynamics) {
        const int rc = launch_dynamics(s, st);
        if (rc != RLOA_OK) return rc;
    }
    if (!skip_dynamics || prepared_thr != c.contact_thr) {        // collision phase on the pose the step starts from
        const int rc = launch_contacts(s, c.contact_thr, st);
        if (rc != RLOA_OK) return rc;
    }
    if (h.nl <= 16) return launch_solve<16>(s, c, mode, actions, active, obs, reward, done, valid, st);
    return launch_solve<32>(s, c, mode, actions, active, obs, reward, done, valid, st);
}

extern "C" int rloa_sim_set_contacts(rloa_sim* s, float contact_threshold) {
    RLOA_REQUIRE(s != nullptr && contact_threshold >= 0.f, "rloa_sim_set_contacts: bad argument");
    s->contact_thr = contact_threshold;
    return RLOA_OK;
}

extern "C" int rloa_sim_contact_counts(rloa_sim* s, int32_t* counts, void* stream) {
    RLOA_REQUIRE(s != nullptr && counts != nullptr, "rloa_sim_contact_counts: bad argument");
    RLOA_CUDA(cudaMemcpyAsync(counts, s->a.near, sizeof(int32_t) * (size_t)s->a.n_envs, cudaMemcpyDeviceToDevice,
                              as_stream(stream)));
    return RLOA_OK;
}

extern "C" int rloa_sim_prepare(rloa_sim* s, void* stream) {
    RLOA_REQUIRE(s != nullptr && s->side != nullptr, "rloa_sim_prepare: null sim");
    cudaStream_t st = as_stream(stream);
    int rc = sim_sync_prepared(s, st, true);
    if (rc != RLOA_OK) return rc;
    RLOA_CUDA(cudaEventRecord(s->fork_ev, st));
    RLOA_CUDA(cudaStreamWaitEvent(s->side, s->fork_ev, 0));
    rc = launch_dynamics(s, s->side);
    if (rc != RLOA_OK) return rc;
    rc = launch_contacts(s, s->contact_thr, s->side);
    if (rc != RLOA_OK) return rc;
    s->prepared_thr = s->contact_thr;
    RLOA_CUDA(cudaEventRecord(s->join_ev, s->side));
    s->prepared = true;
    s->join_pending = true;
    return RLOA_OK;
}

extern "C" int rloa_sim_join(rloa_sim* s, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_join: null sim");
    return sim_sync_prepared(s, as_stream(stream), false);
}

extern "C" int rloa_sim_step(rloa_sim* s, const rloa_step_config* cfg, const float* actions, const uint8_t* active,
                             float* obs, float* reward, uint8_t* done, uint8_t* valid, void* stream) {
    RLOA_REQUIRE(s && cfg && actions && obs && reward && done, "rloa_sim_step: null argument");
    StepCfgDev c;
    int rc = make_step_cfg(s, cfg, &c);
    if (rc != RLOA_OK) return rc;
    const bool prepared = s->prepared;
    rc = sim_sync_prepared(s, as_stream(stream), true);     // the step consumes (and ends) the prepared state
    if (rc != RLOA_OK) return rc;
    return launch_substep(s, c, kModeStep, actions, active, obs, reward, done, valid, as_stream(stream), prepared,
                          prepared ? s->prepared_thr : -1.f);
}

extern "C" int rloa_sim_begin_reset(rloa_sim* s, const uint8_t* mask, const float* init_targets, int32_t n_init,
                                    int32_t n_substeps, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_begin_reset: null sim");
    RLOA_REQUIRE(n_init >= 0 && n_init <= s->a.nl, "rloa_sim_begin_reset: n_init out of range");
    RLOA_REQUIRE(n_init == 0 || init_targets != nullptr, "rloa_sim_begin_reset: init_targets missing");
    RLOA_REQUIRE(n_substeps >= 0, "rloa_sim_begin_reset: n_substeps < 0");
    sim_begin_reset_kernel<<<(unsigned)((s->a.n_envs + 255) / 256), 256, 0, as_stream(stream)>>>(
        s->a, mask, init_targets, n_init, n_substeps, 100000.f * s->model->host.dt);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

extern "C" int rloa_sim_begin_reset_random(rloa_sim* s, const uint8_t* mask, const float* pos, const float* var,
                                           int32_t n_init, int32_t n_substeps, uint64_t seed, const uint64_t* tick,
                                           void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_begin_reset_random: null sim");
    RLOA_REQUIRE(n_init >= 0 && n_init <= s->a.nl, "rloa_sim_begin_reset_random: n_init out of range");
    RLOA_REQUIRE(n_substeps >= 0, "rloa_sim_begin_reset_random: n_substeps < 0");
    sim_begin_reset_random_kernel<<<(unsigned)((s->a.n_envs + 255) / 256), 256, 0, as_stream(stream)>>>(
        s->a, mask, pos, var, n_init, n_substeps, 100000.f * s->model->host.dt, seed,
        reinterpret_cast<const unsigned long long*>(tick));
    RLOA_LAUNCHED();
    return RLOA_OK;
}

extern "C" int rloa_sim_reset(rloa_sim* s, const uint8_t* mask, const float* init_targets, int32_t n_init,
                              int32_t n_substeps, float* obs, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_reset: null sim");
    int rc = sim_sync_prepared(s, as_stream(stream), true);
    if (rc != RLOA_OK) return rc;
    rc = rloa_sim_begin_reset(s, mask, init_targets, n_init, n_substeps, stream);
    if (rc != RLOA_OK) return rc;
    StepCfgDev c;
    std::memset(&c, 0, sizeof(c));
    for (int i = 0; i < kMaxLinks; i++) c.act_index[i] = -1;
    c.contact_thr = s->contact_thr;          // Environment.reset steps the same world: contacts included
    // the masked envs run their n_substeps reset sub-steps back to back; the last one writes their state
    for (int k = 0; k < n_substeps; k++) {
        const int r = launch_substep(s, c, kModeResetOnly, nullptr, nullptr, k == n_substeps - 1 ? obs : nullptr, nullptr,
                                     nullptr, nullptr, as_stream(stream));
        if (r != RLOA_OK) return r;
    }
    if (n_substeps == 0 && obs != nullptr) return rloa_sim_observe(s, obs, nullptr, nullptr, stream);
    return RLOA_OK;
}

extern "C" int rloa_sim_observe(const rloa_sim* s, float* obs, float* link_obstacle, float* ee_target, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_observe: null sim");
    const ModelDev& h = s->model->host;
    const unsigned blocks = (unsigned)((s->a.n_envs + kTpb - 1) / kTpb);
    if (h.use_gjk) {
        if (h.nl <= 16) sim_observe_kernel<16, true><<<blocks, kTpb, 0, as_stream(stream)>>>(h, s->a, obs, link_obstacle, ee_target);
        else sim_observe_kernel<32, true><<<blocks, kTpb, 0, as_stream(stream)>>>(h, s->a, obs, link_obstacle, ee_target);
    } else if (h.nl <= 16) sim_observe_kernel<16, false><<<blocks, kTpb, 0, as_stream(stream)>>>(h, s->a, obs, link_obstacle, ee_target);
    else sim_observe_kernel<32, false><<<blocks, kTpb, 0, as_stream(stream)>>>(h, s->a, obs, link_obstacle, ee_target);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

extern "C" int rloa_sim_self_distances(const rloa_sim* s, float* link_link, void* stream) {
    RLOA_REQUIRE(s != nullptr && link_link != nullptr, "rloa_sim_self_distances: null argument");
    const ModelDev& h = s->model->host;
    const unsigned blocks = (unsigned)((s->a.n_envs + kTpb - 1) / kTpb);
    if (h.nl <= 16) sim_self_distances_kernel<16><<<blocks, kTpb, 0, as_stream(stream)>>>(h, s->a, link_link);
    else sim_self_distances_kernel<32><<<blocks, kTpb, 0, as_stream(stream)>>>(h, s->a, link_link);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

extern "C" int rloa_sim_last_iterations(const rloa_sim* s, int32_t* iters, void* stream) {
    RLOA_REQUIRE(s != nullptr && iters != nullptr, "rloa_sim_last_iterations: null argument");
    RLOA_CUDA(cudaMemcpyAsync(iters, s->a.iters, (size_t)s->a.n_envs * sizeof(int), cudaMemcpyDeviceToDevice,
                              as_stream(stream)));
    return RLOA_OK;
}

extern "C" int rloa_episode_update(int32_t n_envs, int32_t frames, const float* reward, const uint8_t* done,
                                   const uint8_t* active, float* score, int32_t* frame, uint8_t* reset_mask,
                                   float* log_score, int32_t* log_frame, float* log_last_reward, int32_t* log_env,
                                   int32_t log_capacity, int32_t* log_count, int64_t* transitions, uint64_t* tick,
                                   void* stream) {
    RLOA_REQUIRE(reward && done && score && frame && reset_mask && log_score && log_frame && log_last_reward &&
                     log_env && log_count, "rloa_episode_update: null argument");
    RLOA_REQUIRE(n_envs >= 1, "rloa_episode_update: n_envs >= 1 required");
    episode_update_kernel<<<(unsigned)((n_envs + 255) / 256), 256, 0, as_stream(stream)>>>(
        n_envs, frames, reward, done, active, score, frame, reset_mask, log_score, log_frame, log_last_reward, log_env,
        log_capacity, log_count, reinterpret_cast<unsigned long long*>(transitions),
        reinterpret_cast<unsigned long long*>(tick));
    RLOA_LAUNCHED();
    return RLOA_OK;
}

extern "C" int rloa_episode_update_reset(rloa_sim* s, int32_t frames, const float* reward, const uint8_t* done,
                                         const uint8_t* active, float* score, int32_t* frame, uint8_t* reset_mask,
                                         float* log_score, int32_t* log_frame, float* log_last_reward, int32_t* log_env,
                                         int32_t log_capacity, int32_t* log_count, int64_t* transitions, uint64_t* tick,
                                         uint64_t* tick_next, const float* pos, const float* var, int32_t n_init,
                                         int32_t n_substeps, uint64_t seed, void* stream) {
    RLOA_REQUIRE(s != nullptr && s->ticket != nullptr, "rloa_episode_update_reset: null sim");
    RLOA_REQUIRE(reward && done && score && frame && reset_mask && log_score && log_frame && log_last_reward &&
                     log_env && log_count, "rloa_episode_update_reset: null argument");
    RLOA_REQUIRE(n_init >= 0 && n_init <= s->a.nl && n_substeps >= 0, "rloa_episode_update_reset: n_init / n_substeps out of range");
    const int n = s->a.n_envs;
    // with a ping-pong counter nothing this kernel writes is read by the caller's stream before rloa_sim_join, so while
    // a rloa_sim_prepare is in flight it runs behind it on the simulator's side stream (beside the NAF update)
    const bool beside = tick != nullptr && tick_next != nullptr && tick_next != tick && s->join_pending;
    cudaStream_t st = beside ? s->side : as_stream(stream);
    episode_update_reset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        n, frames, reward, done, active, score, frame, reset_mask, log_score, log_frame, log_last_reward, log_env,
        log_capacity, log_count, reinterpret_cast<unsigned long long*>(transitions),
        reinterpret_cast<unsigned long long*>(tick), reinterpret_cast<unsigned long long*>(tick_next), s->a, pos, var,
        n_init, n_substeps, 100000.f * s->model->host.dt, seed, s->ticket);
    RLOA_LAUNCHED();
    if (beside) RLOA_CUDA(cudaEventRecord(s->join_ev, s->side));
    return RLOA_OK;
}
