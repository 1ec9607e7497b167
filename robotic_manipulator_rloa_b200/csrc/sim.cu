// librloa_b200: batched manipulator simulator — kernels and C ABI (include/rloa_b200.h).
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "sim_device.cuh"

namespace rloa {

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
struct SimArrays {
    int n_envs, nl;
    float *q, *qd;                          // [N][nl]
    float *kp, *tpos, *tvel, *maximp;       // [N][nl] motor table
    float *target, *obstacle;               // [N][3]
    int* iters;                             // [N]
    int* reset_left;                        // [N] pending reset sub-steps (lock-step asynchronous reset)
};

struct StepCfgDev {
    signed char act_index[kMaxLinks];       // index into the action vector, -1 = not involved
    unsigned fixed_mask;                    // joints held with POSITION_CONTROL target 0
    int n_act;
    float vel_maximp;                       // max_force * dt
    float pos_maximp;                       // 1e5 * dt (pybullet POSITION_CONTROL default force)
    float target_thr, obstacle_thr;
};

__device__ __forceinline__ void write_obs(const ModelDev* __restrict__ M, int lane, float q, float qd, V3 ee, V3 tg,
                                          V3 ob, float* __restrict__ o) {
    const int n = M->n_obs;
    if (lane < n) {
        o[lane] = q;
        o[n + lane] = qd;
    }
    if (lane == 0) {
        float* t = o + 2 * n;
        t[0] = ee.x; t[1] = ee.y; t[2] = ee.z;
        t[3] = tg.x; t[4] = tg.y; t[5] = tg.z;
        t[6] = ob.x; t[7] = ob.y; t[8] = ob.z;
    }
}

__device__ __forceinline__ V3 load3(const float* p) { return v3(p[0], p[1], p[2]); }

// Environment.step (reference environment.py:453-485) for one env per warp
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sim_step_kernel(const ModelDev* __restrict__ M, SimArrays S, StepCfgDev cfg, const float* __restrict__ actions,
                const uint8_t* __restrict__ active, float* __restrict__ obs, float* __restrict__ reward,
                uint8_t* __restrict__ done, uint8_t* __restrict__ valid_out) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * kWarpsPerBlock + warp;
    if (env >= S.n_envs) return;
    if (active != nullptr && active[env] == 0) return;
    const int nl = S.nl;
    float* sm = smem + warp * sim_smem_floats_per_warp(nl);
    const bool valid = lane < nl;
    const size_t base = (size_t)env * nl + (valid ? lane : 0);
    float q = 0.f, qd = 0.f;
    Motor mot{0.f, 0.f, 0.f, 0.f};
    // an env with pending reset sub-steps spends this launch on one of them (motors as begin_reset left
    // them) instead of an action step; it emits no transition (valid = 0)
    const int pending = S.reset_left[env];
    if (pending > 0) {
        if (valid) {
            q = S.q[base];
            qd = S.qd[base];
            mot = Motor{S.kp[base], S.tpos[base], S.tvel[base], S.maximp[base]};
        }
        const int it = warp_substep(M, sm, lane, q, qd, mot);
        if (valid) {
            S.q[base] = q;
            S.qd[base] = qd;
        }
        M3 R;
        V3 p;
        lane_fk(M, valid ? lane : 0, valid, q, R, p);
        const V3 tg = load3(S.target + 3 * (size_t)env), ob = load3(S.obstacle + 3 * (size_t)env);
        write_obs(M, lane, q, qd, sh3(p, M->ee_link), tg, ob, obs + (size_t)env * (9 + 2 * M->n_obs));
        if (lane == 0) {
            S.reset_left[env] = pending - 1;
            reward[env] = 0.f;
            done[env] = 0;
            S.iters[env] = it;
            if (valid_out != nullptr) valid_out[env] = 0;
        }
        return;
    }
    if (valid) {
        q = S.q[base];
        qd = S.qd[base];
        // setJointMotorControl2: VELOCITY_CONTROL on involved joints (environment.py:464-469), then
        // POSITION_CONTROL target 0 on the fixed joints (:472-476); other joints keep their motor
        const int ai = cfg.act_index[lane];
        const bool fixedj = (cfg.fixed_mask >> lane) & 1u;
        if (fixedj) {
            mot = Motor{0.1f, 0.f, 0.f, cfg.pos_maximp};
        } else if (ai >= 0) {
            mot = Motor{0.f, 0.f, actions[(size_t)env * cfg.n_act + ai], cfg.vel_maximp};
        } else {
            mot = Motor{S.kp[base], S.tpos[base], S.tvel[base], S.maximp[base]};
        }
        if (fixedj || ai >= 0) {
            S.kp[base] = mot.kp; S.tpos[base] = mot.tpos; S.tvel[base] = mot.tvel; S.maximp[base] = mot.maximp;
        }
    }
    const int it = warp_substep(M, sm, lane, q, qd, mot);
    if (valid) {
        S.q[base] = q;
        S.qd[base] = qd;
    }
    // state / reward / done from the post-step configuration
    M3 R;
    V3 p;
    lane_fk(M, valid ? lane : 0, valid, q, R, p);
    const V3 tg = load3(S.target + 3 * (size_t)env), ob = load3(S.obstacle + 3 * (size_t)env);
    const ObsOut o = lane_distances(M, lane, R, p, ob, tg, cfg.obstacle_thr, false);
    write_obs(M, lane, q, qd, o.ee_pos, tg, ob, obs + (size_t)env * (9 + 2 * M->n_obs));
    if (lane == 0) {
        const bool goal = o.ee_target < cfg.target_thr;
        // get_reward (environment.py:364-371): goal first, then collision, else -(d - threshold)
        reward[env] = goal ? 250.f : (o.hit ? -1000.f : -(o.ee_target - cfg.target_thr));
        done[env] = (o.hit || goal) ? 1 : 0;         // is_terminal_state (environment.py:326-333)
        S.iters[env] = it;
        if (valid_out != nullptr) valid_out[env] = 1;
    }
}

// lock-step asynchronous Environment.reset: arm the POSITION_CONTROL motors of the masked envs and let
// the next n_substeps step launches run their reset sub-steps (environment.py:295-301)
__global__ void sim_begin_reset_kernel(SimArrays S, const uint8_t* __restrict__ mask,
                                       const float* __restrict__ init_targets, int n_init, int nsub, float pos_maximp) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)S.n_envs * S.nl;
    if (i >= n) return;
    const int env = (int)(i / S.nl), j = (int)(i - (size_t)env * S.nl);
    if (mask != nullptr && mask[env] == 0) return;
    if (j < n_init) {
        S.kp[i] = 0.1f; S.tpos[i] = init_targets[(size_t)env * n_init + j]; S.tvel[i] = 0.f; S.maximp[i] = pos_maximp;
    }
    if (j == 0) S.reset_left[env] = nsub;
}

// Environment.reset (reference environment.py:264-309) for the masked envs
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sim_reset_kernel(const ModelDev* __restrict__ M, SimArrays S, const uint8_t* __restrict__ mask,
                 const float* __restrict__ init_targets, int n_init, int nsub, float pos_maximp,
                 float* __restrict__ obs) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * kWarpsPerBlock + warp;
    if (env >= S.n_envs) return;
    if (mask != nullptr && mask[env] == 0) return;
    const int nl = S.nl;
    float* sm = smem + warp * sim_smem_floats_per_warp(nl);
    const bool valid = lane < nl;
    const size_t base = (size_t)env * nl + (valid ? lane : 0);
    float q = 0.f, qd = 0.f;
    Motor mot{0.f, 0.f, 0.f, 0.f};
    if (valid) {
        q = S.q[base];
        qd = S.qd[base];
        if (lane < n_init) {   // POSITION_CONTROL defaults: kp 0.1, kd 1, force 1e5 (environment.py:295-298)
            mot = Motor{0.1f, init_targets[(size_t)env * n_init + lane], 0.f, pos_maximp};
            S.kp[base] = mot.kp; S.tpos[base] = mot.tpos; S.tvel[base] = mot.tvel; S.maximp[base] = mot.maximp;
        } else {
            mot = Motor{S.kp[base], S.tpos[base], S.tvel[base], S.maximp[base]};
        }
    }
    int it = 0;
    for (int s = 0; s < nsub; s++) it = warp_substep(M, sm, lane, q, qd, mot);
    if (valid) {
        S.q[base] = q;
        S.qd[base] = qd;
    }
    if (lane == 0) {
        S.iters[env] = it;
        S.reset_left[env] = 0;
    }
    if (obs != nullptr) {
        M3 R;
        V3 p;
        lane_fk(M, valid ? lane : 0, valid, q, R, p);
        const V3 tg = load3(S.target + 3 * (size_t)env), ob = load3(S.obstacle + 3 * (size_t)env);
        const V3 ee = sh3(p, M->ee_link);
        write_obs(M, lane, q, qd, ee, tg, ob, obs + (size_t)env * (9 + 2 * M->n_obs));
    }
}

// get_state + the distances behind get_reward / is_terminal_state, without stepping
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sim_observe_kernel(const ModelDev* __restrict__ M, SimArrays S, float* __restrict__ obs, float* __restrict__ link_obst,
                   float* __restrict__ ee_target) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * kWarpsPerBlock + warp;
    if (env >= S.n_envs) return;
    const int nl = S.nl;
    const bool valid = lane < nl;
    const size_t base = (size_t)env * nl + (valid ? lane : 0);
    const float q = valid ? S.q[base] : 0.f, qd = valid ? S.qd[base] : 0.f;
    M3 R;
    V3 p;
    lane_fk(M, valid ? lane : 0, valid, q, R, p);
    const V3 tg = load3(S.target + 3 * (size_t)env), ob = load3(S.obstacle + 3 * (size_t)env);
    const ObsOut o = lane_distances(M, lane, R, p, ob, tg, 0.f, link_obst != nullptr);
    if (obs != nullptr) write_obs(M, lane, q, qd, o.ee_pos, tg, ob, obs + (size_t)env * (9 + 2 * M->n_obs));
    if (link_obst != nullptr && valid) link_obst[base] = o.link_dist;
    if (ee_target != nullptr && lane == 0) ee_target[env] = o.ee_target;
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
                                      int* __restrict__ log_count, unsigned long long* __restrict__ transitions) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
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

}  // namespace rloa

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace rloa;

struct rloa_model {
    ModelDev host;
    ModelDev* dev = nullptr;
    int device = 0;
};

struct rloa_sim {
    const rloa_model* model = nullptr;
    SimArrays a{};
    int device = 0;
    size_t smem_bytes = 0;
};

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
    RLOA_REQUIRE(d->ns >= 0 && d->ns <= kMaxShapes, "rloa_model_create: at most 32 collision primitives supported");
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
    int ndof = 0, maxdepth = 0, nslots = 0;
    std::vector<int> slot(nl, -1);
    for (int i = 0; i < nl; i++) {
        const int p = d->parent[i];
        if (!(p >= -1 && p < i)) { delete m; return fail(RLOA_ERR_INVALID, "rloa_model_create: links must be numbered depth first (parent < child)"); }
        h.parent[i] = p;
        h.jtype[i] = d->jtype[i];
        h.depth[i] = p < 0 ? 0 : h.depth[p] + 1;
        if (h.depth[i] > maxdepth) maxdepth = h.depth[i];
        h.has_limit[i] = d->has_limit[i];
        h.dofidx[i] = -1;
        if (d->jtype[i] != RLOA_JOINT_FIXED) {
            if (ndof >= kMaxDof) { delete m; return fail(RLOA_ERR_INVALID, "rloa_model_create: at most 16 movable joints supported"); }
            if (!(d->mass[i] > 0)) { delete m; return fail(RLOA_ERR_INVALID, "rloa_model_create: movable link without mass"); }
            h.dofidx[i] = ndof;
            h.doflink[ndof++] = i;
        }
        if (p >= 0) {
            if (h.nch[p] >= kMaxChildren) { delete m; return fail(RLOA_ERR_INVALID, "rloa_model_create: at most 4 children per link supported"); }
            h.child[p][h.nch[p]++] = i;
        }
        h.anc_mask[i] = (1u << i) | (p >= 0 ? h.anc_mask[p] : 0u);
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
    h.ndof = ndof; h.maxdepth = maxdepth;
    h.nrounds = 0;
    while ((1 << h.nrounds) < maxdepth + 1) h.nrounds++;
    for (int i = 0; i < nl; i++) {          // where the unit-response sweep finds the parent's acceleration
        const int p = h.parent[i];
        h.accsave[i] = -1;
        if (p < 0) h.accsrc[i] = 0;
        else if (p == i - 1) h.accsrc[i] = 1;
        else {
            if (slot[p] < 0) {
                if (nslots >= kMaxSlots) { delete m; return fail(RLOA_ERR_INVALID, "rloa_model_create: too many branching links (max 3)"); }
                slot[p] = nslots++;
            }
            h.accsrc[i] = 2 + slot[p];
        }
    }
    for (int i = 0; i < nl; i++) h.accsave[i] = slot[i];
    for (int i = 0; i < nl; i++)
        if (h.nch[i] > h.lvl_maxch[h.depth[i] + 1]) h.lvl_maxch[h.depth[i] + 1] = h.nch[i];
    for (int s = 0; s < d->ns; s++) {
        if (!(d->s_link[s] >= 0 && d->s_link[s] < nl)) { delete m; return fail(RLOA_ERR_INVALID, "rloa_model_create: shape link out of range"); }
        h.s_link[s] = d->s_link[s];
        h.s_type[s] = d->s_type[s];
        for (int k = 0; k < 9; k++) h.s_R[s][k] = (float)d->s_R[9 * s + k];
        for (int k = 0; k < 3; k++) { h.s_p[s][k] = (float)d->s_p[3 * s + k]; h.s_dim[s][k] = (float)d->s_dim[3 * s + k]; }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { delete m; cudaGetLastError(); return fail(RLOA_ERR_NO_DEVICE, "rloa_model_create: no CUDA device visible"); }
    if (cudaGetDevice(&m->device) != cudaSuccess || cudaMalloc(&m->dev, sizeof(ModelDev)) != cudaSuccess ||
        cudaMemcpy(m->dev, &h, sizeof(ModelDev), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("rloa_model_create: CUDA allocation/copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete m;
        return RLOA_ERR_CUDA;
    }
    *out = m;
    return RLOA_OK;
}

extern "C" void rloa_model_destroy(rloa_model* m) {
    if (m == nullptr) return;
    if (m->dev) cudaFree(m->dev);
    delete m;
}

extern "C" int rloa_sim_create(const rloa_model* m, int32_t n_envs, rloa_sim** out) {
    RLOA_REQUIRE(m != nullptr && out != nullptr, "rloa_sim_create: null argument");
    RLOA_REQUIRE(n_envs >= 1, "rloa_sim_create: n_envs >= 1 required");
    rloa_sim* s = new (std::nothrow) rloa_sim();
    RLOA_REQUIRE(s != nullptr, "rloa_sim_create: out of host memory");
    s->model = m;
    s->device = m->device;
    const int nl = m->host.nl;
    const size_t n = (size_t)n_envs * nl;
    s->a.n_envs = n_envs;
    s->a.nl = nl;
    float* block = nullptr;
    // one allocation: q qd kp tpos tvel maximp | target obstacle | iters
    const size_t floats = 6 * n + 6 * (size_t)n_envs + 2 * (size_t)n_envs;
    if (cudaMalloc(&block, floats * sizeof(float)) != cudaSuccess) {
        set_error("rloa_sim_create: cudaMalloc of %zu bytes failed: %s", floats * sizeof(float), cudaGetErrorString(cudaGetLastError()));
        delete s;
        return RLOA_ERR_CUDA;
    }
    s->a.q = block; s->a.qd = block + n; s->a.kp = block + 2 * n; s->a.tpos = block + 3 * n;
    s->a.tvel = block + 4 * n; s->a.maximp = block + 5 * n;
    s->a.target = block + 6 * n; s->a.obstacle = s->a.target + 3 * (size_t)n_envs;
    s->a.iters = reinterpret_cast<int*>(s->a.obstacle + 3 * (size_t)n_envs);
    s->a.reset_left = s->a.iters + n_envs;
    s->smem_bytes = (size_t)kWarpsPerBlock * sim_smem_floats_per_warp(nl) * sizeof(float);
    if (s->smem_bytes > 48 * 1024) {
        cudaFuncSetAttribute(sim_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes);
        cudaFuncSetAttribute(sim_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes);
    }
    cudaMemset(s->a.target, 0, 6 * (size_t)n_envs * sizeof(float));
    const int threads = 256;
    sim_clear_kernel<<<(unsigned)((n + threads - 1) / threads), threads>>>(s->a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        set_error("rloa_sim_create: initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(block);
        delete s;
        return RLOA_ERR_CUDA;
    }
    *out = s;
    return RLOA_OK;
}

extern "C" void rloa_sim_destroy(rloa_sim* s) {
    if (s == nullptr) return;
    if (s->a.q) cudaFree(s->a.q);
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

extern "C" int rloa_sim_set_state(rloa_sim* s, const float* q, const float* qd, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_set_state: null sim");
    const size_t bytes = (size_t)s->a.n_envs * s->a.nl * sizeof(float);
    if (q) RLOA_CUDA(cudaMemcpyAsync(s->a.q, q, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    if (qd) RLOA_CUDA(cudaMemcpyAsync(s->a.qd, qd, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return RLOA_OK;
}

extern "C" int rloa_sim_get_state(const rloa_sim* s, float* q, float* qd, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_get_state: null sim");
    const size_t bytes = (size_t)s->a.n_envs * s->a.nl * sizeof(float);
    if (q) RLOA_CUDA(cudaMemcpyAsync(q, s->a.q, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    if (qd) RLOA_CUDA(cudaMemcpyAsync(qd, s->a.qd, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return RLOA_OK;
}

extern "C" int rloa_sim_set_motors(rloa_sim* s, const float* kp, const float* tpos, const float* tvel,
                                   const float* maximp, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_set_motors: null sim");
    const size_t bytes = (size_t)s->a.n_envs * s->a.nl * sizeof(float);
    if (kp) RLOA_CUDA(cudaMemcpyAsync(s->a.kp, kp, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    if (tpos) RLOA_CUDA(cudaMemcpyAsync(s->a.tpos, tpos, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    if (tvel) RLOA_CUDA(cudaMemcpyAsync(s->a.tvel, tvel, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    if (maximp) RLOA_CUDA(cudaMemcpyAsync(s->a.maximp, maximp, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return RLOA_OK;
}

extern "C" int rloa_sim_clear(rloa_sim* s, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_clear: null sim");
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
    return RLOA_OK;
}

extern "C" int rloa_sim_step(rloa_sim* s, const rloa_step_config* cfg, const float* actions, const uint8_t* active,
                             float* obs, float* reward, uint8_t* done, uint8_t* valid, void* stream) {
    RLOA_REQUIRE(s && cfg && actions && obs && reward && done, "rloa_sim_step: null argument");
    StepCfgDev c;
    const int rc = make_step_cfg(s, cfg, &c);
    if (rc != RLOA_OK) return rc;
    const unsigned blocks = (unsigned)((s->a.n_envs + kWarpsPerBlock - 1) / kWarpsPerBlock);
    sim_step_kernel<<<blocks, kWarpsPerBlock * 32, s->smem_bytes, as_stream(stream)>>>(s->model->dev, s->a, c, actions,
                                                                                     active, obs, reward, done, valid);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

extern "C" int rloa_sim_begin_reset(rloa_sim* s, const uint8_t* mask, const float* init_targets, int32_t n_init,
                                    int32_t n_substeps, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_begin_reset: null sim");
    RLOA_REQUIRE(n_init >= 0 && n_init <= s->a.nl, "rloa_sim_begin_reset: n_init out of range");
    RLOA_REQUIRE(n_init == 0 || init_targets != nullptr, "rloa_sim_begin_reset: init_targets missing");
    RLOA_REQUIRE(n_substeps >= 0, "rloa_sim_begin_reset: n_substeps < 0");
    const size_t n = (size_t)s->a.n_envs * s->a.nl;
    sim_begin_reset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
        s->a, mask, init_targets, n_init, n_substeps, 100000.f * s->model->host.dt);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

extern "C" int rloa_sim_reset(rloa_sim* s, const uint8_t* mask, const float* init_targets, int32_t n_init,
                              int32_t n_substeps, float* obs, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_reset: null sim");
    RLOA_REQUIRE(n_init >= 0 && n_init <= s->a.nl, "rloa_sim_reset: n_init out of range");
    RLOA_REQUIRE(n_init == 0 || init_targets != nullptr, "rloa_sim_reset: init_targets missing");
    RLOA_REQUIRE(n_substeps >= 0, "rloa_sim_reset: n_substeps < 0");
    const unsigned blocks = (unsigned)((s->a.n_envs + kWarpsPerBlock - 1) / kWarpsPerBlock);
    sim_reset_kernel<<<blocks, kWarpsPerBlock * 32, s->smem_bytes, as_stream(stream)>>>(
        s->model->dev, s->a, mask, init_targets, n_init, n_substeps, 100000.f * s->model->host.dt, obs);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

extern "C" int rloa_sim_observe(const rloa_sim* s, float* obs, float* link_obstacle, float* ee_target, void* stream) {
    RLOA_REQUIRE(s != nullptr, "rloa_sim_observe: null sim");
    const unsigned blocks = (unsigned)((s->a.n_envs + kWarpsPerBlock - 1) / kWarpsPerBlock);
    sim_observe_kernel<<<blocks, kWarpsPerBlock * 32, 0, as_stream(stream)>>>(s->model->dev, s->a, obs, link_obstacle,
                                                                            ee_target);
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
                                   int32_t log_capacity, int32_t* log_count, int64_t* transitions, void* stream) {
    RLOA_REQUIRE(reward && done && score && frame && reset_mask && log_score && log_frame && log_last_reward &&
                     log_env && log_count, "rloa_episode_update: null argument");
    RLOA_REQUIRE(n_envs >= 1, "rloa_episode_update: n_envs >= 1 required");
    episode_update_kernel<<<(unsigned)((n_envs + 255) / 256), 256, 0, as_stream(stream)>>>(
        n_envs, frames, reward, done, active, score, frame, reset_mask, log_score, log_frame, log_last_reward, log_env,
        log_capacity, log_count, reinterpret_cast<unsigned long long*>(transitions));
    RLOA_LAUNCHED();
    return RLOA_OK;
}
