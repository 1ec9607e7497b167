// Device side of the batched manipulator simulator: one THREAD per arm, struct-of-arrays state.
//
// What it replaces: the Bullet3 multibody step that runs under the reference's
// Environment.step / Environment.reset (reference environment/environment.py:264-309, 453-485) —
// btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof, calcAccelerationDeltasMultiDof,
// btMultiBodyJointMotor / btMultiBodyJointLimitConstraint rows, the PGS sweep of
// btMultiBodyConstraintSolver and stepPositionsMultiDof — plus the getClosestPoints queries of
// utils/collision_detector.py:33-61 and the reward / terminal logic of environment.py:311-371.
//
// Mapping (B200-first, not Bullet's).  A first version gave every arm a warp (lane = link); ncu showed
// 21k warp-instructions per env-step, 52 % of them in a row-space Gauss-Seidel sweep that used one lane
// per row and the rest in 13 serial tree levels that all 32 lanes executed redundantly
// (profiles/r1a_sim_step_kernel_ncu_full.md).  The serial chain of a manipulator cannot be spread over
// lanes, so lanes are spread over ARMS instead: every quantity is [component][env] in HBM (coalesced),
// per-arm temporaries live in registers / L1-resident local memory, and one env-step is three launches
//   1. sim_dynamics_kernel   thread = arm:           FK, velocities, bias forces, ABA factorisation,
//                                                     free accelerations -> factor records F, qs
//   2. sim_minv_kernel       warp = (32 arms, dof j): column j of M^-1 by a unit impulse through F
//   3. sim_solve_kernel<D>   thread = arm:           constraint rows, PGS with M^-1 in registers,
//                                                     integration, FK, distances, reward / done, obs
// All spatial quantities are in WORLD orientation with the origin at the owning link's centre of mass,
// so the transform between a link and its parent is a pure translation r = p_child - p_parent.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rloa_b200.h"

namespace rloa {

constexpr int kMaxLinks = 32;
constexpr int kMaxShapes = 32;
constexpr int kMaxDof = 16;
constexpr int kMaxSlots = 3;
constexpr int kFRec = 16;                // factor record per link: s_a s_l h_a h_l r invD (F[link][k][env])
constexpr int kLRec = 28;                // per-link local record of the dynamics kernel (forward-sweep results + u)
constexpr int kTpb = 32;                 // threads (= arms) per block of the thread-per-arm kernels
constexpr int kMaxContacts = 4;          // normal contact rows per arm (obstacle sphere / target cube vs the link shapes), in
                                         // shape order; every row costs ~80 dependent cycles per sweep of the thread-per-arm PGS

// fp32 device copy of rloa_model_desc plus the tree tables the kernels need; passed BY VALUE as a
// __grid_constant__ kernel parameter so every access is a uniform constant-bank load
struct ModelDev {
    int nl, ns, ndof, ee_link, n_obs, iters, nslots, pad0;
    float dt, inv_dt, lin_damp, ang_damp, resid_thresh, erp, max_vel, limit_max_imp;
    float gravity[3];
    float obstacle_radius;
    float target_half[3];
    float pad1;
    int parent[kMaxLinks], jtype[kMaxLinks], has_limit[kMaxLinks];
    int dofidx[kMaxLinks];               // compact dof index of a movable link, -1 otherwise
    int doflink[kMaxDof];                // link of dof j
    int fwsrc[kMaxLinks];                // root->leaf sweeps: parent value is 0 zero (root) | 1 running | 2+k slot k
    int fwsave[kMaxLinks];               // slot to save the link's value into, -1 none
    int bwdst[kMaxLinks];                // leaf->root fold: -1 discard (root) | 0 running | 1+k slot k
    int bwsrc[kMaxLinks];                // bit 0: take the running accumulator; bits 1..: 1 + slot to take (0 none)
    float E0T[kMaxLinks][9];             // parent COM frame <- child COM frame at q = 0 (base folded into roots)
    float e[kMaxLinks][3];               // parent COM -> pivot in the parent frame (base folded into roots)
    float d[kMaxLinks][3], axis[kMaxLinks][3];
    float mass[kMaxLinks], inertia[kMaxLinks][3], damping[kMaxLinks], lower[kMaxLinks], upper[kMaxLinks];
    int s_link[kMaxShapes], s_type[kMaxShapes];
    float s_R[kMaxShapes][9], s_p[kMaxShapes][3], s_dim[kMaxShapes][3];
    // bounding sphere of every shape: centre in the LINK frame, radius (margins / radii included, rounded up)
    float s_bs[kMaxShapes][4];
    // convex-hull shapes (mesh / cylinder collision geometry): vertex range in verts
    int s_v0[kMaxShapes], s_vn[kMaxShapes];
    const float4* verts;                 // device, owned by rloa_model; xyz = vertex in the shape frame
    int use_gjk;                         // the model has hull shapes or box shapes on the end-effector link
    int pad2;
};

struct SimArrays {
    int n_envs, nl, ndof, pad;
    float *q, *qd;                          // [nl][N]
    float *kp, *tpos, *tvel, *maximp;       // [nl][N] motor table (what setJointMotorControl2 left behind)
    float *target, *obstacle;               // [N][3]
    int* iters;                             // [N]
    int* reset_left;                        // [N] pending reset sub-steps (lock-step asynchronous reset)
    int* near;                              // [N] contact rows stored in crow by the latest collision phase (sim_contacts_kernel)
    float* crow;                            // [N][kMaxContacts][kMaxDof + 1] J row | signed distance (NULL-safe: only with contacts)
    float* F;                               // [nl][kFRec][N] ABA factor records (dynamics -> minv)
    float* qs;                              // [nl][N] free velocity qd + dt*qdd, clamped
    float* minv;                            // [ndof][ndof][N]
};

struct StepCfgDev {
    signed char act_index[kMaxLinks];       // index into the action vector, -1 = not involved
    unsigned fixed_mask;                    // joints held with POSITION_CONTROL target 0
    int n_act;
    float vel_maximp;                       // max_force * dt
    float pos_maximp;                       // 1e5 * dt (pybullet POSITION_CONTROL default force)
    float target_thr, obstacle_thr;
    float contact_thr;                      // contact breaking threshold (Bullet: 0.02); 0 = contact rows off
    int contact_dbg;                        // RLOA_CONTACT_DEBUG bits: 1 = drop the rows after loading them, 2 = do not load them
                                            // (collision phase only): timing splits, tools/prof_contacts.py
};

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ V3 fma3(float s, V3 a, V3 b) {   // s*a + b
    return v3(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z));
}
__device__ __forceinline__ V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

// row-major 3x3
struct M3 {
    float m[9];
};
__device__ __forceinline__ V3 mul(const M3& A, V3 x) {
    return v3(fmaf(A.m[0], x.x, fmaf(A.m[1], x.y, A.m[2] * x.z)), fmaf(A.m[3], x.x, fmaf(A.m[4], x.y, A.m[5] * x.z)),
              fmaf(A.m[6], x.x, fmaf(A.m[7], x.y, A.m[8] * x.z)));
}
__device__ __forceinline__ V3 mulT(const M3& A, V3 x) {
    return v3(fmaf(A.m[0], x.x, fmaf(A.m[3], x.y, A.m[6] * x.z)), fmaf(A.m[1], x.x, fmaf(A.m[4], x.y, A.m[7] * x.z)),
              fmaf(A.m[2], x.x, fmaf(A.m[5], x.y, A.m[8] * x.z)));
}
__device__ __forceinline__ M3 mul(const M3& A, const M3& B) {
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C.m[3 * i + j] = fmaf(A.m[3 * i], B.m[j], fmaf(A.m[3 * i + 1], B.m[3 + j], A.m[3 * i + 2] * B.m[6 + j]));
    return C;
}

// symmetric 3x3: xx xy xz yy yz zz
struct S3 {
    float xx, xy, xz, yy, yz, zz;
};
__device__ __forceinline__ V3 mul(const S3& A, V3 v) {
    return v3(fmaf(A.xx, v.x, fmaf(A.xy, v.y, A.xz * v.z)), fmaf(A.xy, v.x, fmaf(A.yy, v.y, A.yz * v.z)),
              fmaf(A.xz, v.x, fmaf(A.yz, v.y, A.zz * v.z)));
}

// ------------------------------------------------------------------------------------------------
// Joint transform of link i at coordinate q: rotation parent COM frame <- link COM frame and the link
// COM in the parent COM frame (Bullet: btQuaternion(axis,-q) * zeroRotParentToThis, rVector = E e + d)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void joint_transform(const ModelDev& M, int i, float q, M3& Rl, V3& pl, V3& dd) {
    M3 E0T;
#pragma unroll
    for (int k = 0; k < 9; k++) E0T.m[k] = M.E0T[i][k];
    const V3 ax = v3(M.axis[i][0], M.axis[i][1], M.axis[i][2]);
    dd = v3(M.d[i][0], M.d[i][1], M.d[i][2]);
    const V3 ee = v3(M.e[i][0], M.e[i][1], M.e[i][2]);
    const int jt = M.jtype[i];
    if (jt == RLOA_JOINT_REVOLUTE) {
        float s, c;
        sincosf(q, &s, &c);
        const float t = 1.f - c;
        M3 Rq;   // Rodrigues(axis, +q)
        Rq.m[0] = fmaf(t * ax.x, ax.x, c);          Rq.m[1] = fmaf(t * ax.x, ax.y, -s * ax.z);  Rq.m[2] = fmaf(t * ax.x, ax.z, s * ax.y);
        Rq.m[3] = fmaf(t * ax.y, ax.x, s * ax.z);   Rq.m[4] = fmaf(t * ax.y, ax.y, c);          Rq.m[5] = fmaf(t * ax.y, ax.z, -s * ax.x);
        Rq.m[6] = fmaf(t * ax.z, ax.x, -s * ax.y);  Rq.m[7] = fmaf(t * ax.z, ax.y, s * ax.x);   Rq.m[8] = fmaf(t * ax.z, ax.z, c);
        Rl = mul(E0T, Rq);
    } else {
        Rl = E0T;
        if (jt == RLOA_JOINT_PRISMATIC) dd = fma3(q, ax, dd);
    }
    pl = ee + mul(Rl, dd);
}

// ------------------------------------------------------------------------------------------------
// closest-distance helpers (getClosestPoints restated for convex primitives; no margins)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float point_box_signed(V3 p, V3 h) {
    float ox = fabsf(p.x) - h.x, oy = fabsf(p.y) - h.y, oz = fabsf(p.z) - h.z;
    float mx = fmaxf(ox, fmaxf(oy, oz));
    float s = 0.f;
    if (ox > 0.f) s = fmaf(ox, ox, s);
    if (oy > 0.f) s = fmaf(oy, oy, s);
    if (oz > 0.f) s = fmaf(oz, oz, s);
    return s > 0.f ? sqrtf(s) : mx;
}
__device__ __forceinline__ float seg_box_f(V3 a, V3 dir, V3 h, float t) {
    float ox = fabsf(fmaf(t, dir.x, a.x)) - h.x, oy = fabsf(fmaf(t, dir.y, a.y)) - h.y,
          oz = fabsf(fmaf(t, dir.z, a.z)) - h.z;
    float s = 0.f;
    if (ox > 0.f) s = fmaf(ox, ox, s);
    if (oy > 0.f) s = fmaf(oy, oy, s);
    if (oz > 0.f) s = fmaf(oz, oz, s);
    return s;
}
// exact distance between segment a..b and the origin-centred box h: the squared distance along the
// segment is convex piecewise quadratic in t with breakpoints where a coordinate crosses a face plane.
__device__ __noinline__ float segment_box(V3 a, V3 b, V3 h) {
    V3 dir = b - a;
    float bp[8];
    int nb = 0;
    bp[nb++] = 0.f;
    bp[nb++] = 1.f;
    const float av[3] = {a.x, a.y, a.z}, dv[3] = {dir.x, dir.y, dir.z}, hv[3] = {h.x, h.y, h.z};
#pragma unroll
    for (int x = 0; x < 3; x++) {
        if (fabsf(dv[x]) < 1e-30f) continue;
        float t1 = (hv[x] - av[x]) / dv[x], t2 = (-hv[x] - av[x]) / dv[x];
        if (t1 > 0.f && t1 < 1.f) bp[nb++] = t1;
        if (t2 > 0.f && t2 < 1.f) bp[nb++] = t2;
    }
    for (int i = 1; i < nb; i++) {
        float v = bp[i];
        int j = i - 1;
        while (j >= 0 && bp[j] > v) {
            bp[j + 1] = bp[j];
            j--;
        }
        bp[j + 1] = v;
    }
    float best = 3.0e38f;
    for (int i = 0; i < nb; i++) {
        best = fminf(best, seg_box_f(a, dir, h, bp[i]));
        if (i + 1 < nb) {
            float tm = 0.5f * (bp[i] + bp[i + 1]), A = 0.f, B = 0.f;
#pragma unroll
            for (int x = 0; x < 3; x++) {
                float xm = fmaf(tm, dv[x], av[x]);
                if (fabsf(xm) > hv[x]) {
                    float sg = xm > 0.f ? 1.f : -1.f;
                    A = fmaf(dv[x], dv[x], A);
                    B = fmaf(2.f * (sg * av[x] - hv[x]) * sg, dv[x], B);
                }
            }
            if (A > 0.f) {
                float ts = -B / (2.f * A);
                if (ts > bp[i] && ts < bp[i + 1]) best = fminf(best, seg_box_f(a, dir, h, ts));
            }
        }
    }
    return sqrtf(best);
}

#include "gjk.cuh"

// ---- contact rows -------------------------------------------------------------
// Candidates of shape s (world pose Rs, ps) against the obstacle sphere and the axis-aligned target cube: the core point x of
// the shape (sphere centre / closest point of the capsule axis / closest point of the box) against the closest point of the
// fixed body; add(link, x, r_shape, from, r_body) keeps the pair when it is closer than the threshold.  Same pairs, same
// order as oracle/bullet_restatement.c::find_contacts.
__device__ __noinline__ float segment_box_argmin(V3 a, V3 b, V3 h);

// sphere / capsule (position ps, world axis ax) against the axis-aligned target cube.  Needs the shape's axis only, not its
// whole frame.  Cheap exact pre-test first: the cube lies inside its bounding sphere, so a shape farther than that from the
// cube's centre cannot be a contact (for a capsule: distance from the centre to its axis) - the piecewise segment / box routine
// then runs only for shapes within a few centimetres of the cube.
template <class Add>
__device__ __forceinline__ void cube_contact(int type, int l, V3 dim, V3 ps, V3 ax, V3 target, V3 th, float thr, Add&& add) {
    const float r_s = dim.x;
    {
        const V3 t = target - ps;
        const float tz = dot(t, ax);
        const float ez = type == RLOA_SHAPE_CAPSULE ? tz - fminf(fmaxf(tz, -dim.y), dim.y) : tz;
        const float r2 = fmaxf(fmaf(ez, ez, fmaf(-tz, tz, dot(t, t))), 0.f);
        const float dc = sqrtf(r2) - r_s - sqrtf(dot(th, th));
        if (!(dc < thr + 1e-5f)) return;         // 1e-5: the pre-test is a filter, rounding must not decide a contact
    }
    V3 x = ps;
    if (type == RLOA_SHAPE_CAPSULE) {
        const V3 a1 = ps - target - dim.y * ax, b1 = ps - target + dim.y * ax;
        const float tb = segment_box_argmin(a1, b1, th);
        x = target + a1 + tb * (b1 - a1);
    }
    const V3 c = x - target;
    const V3 y = target + v3(fminf(fmaxf(c.x, -th.x), th.x), fminf(fmaxf(c.y, -th.y), th.y), fminf(fmaxf(c.z, -th.z), th.z));
    add(l, x, r_s, y, 0.f);
}

template <class Add>
__device__ __forceinline__ void shape_contacts(const ModelDev& M, int s, const M3& Rs, V3 ps, V3 obstacle, V3 target, V3 th,
                                               bool near_o, bool near_t, float M_contact_thr, Add&& add) {
    const int l = M.s_link[s], type = M.s_type[s];
    if (type != RLOA_SHAPE_SPHERE && type != RLOA_SHAPE_CAPSULE && type != RLOA_SHAPE_BOX) return;
    const V3 dim = v3(M.s_dim[s][0], M.s_dim[s][1], M.s_dim[s][2]);
    const V3 ax = v3(Rs.m[2], Rs.m[5], Rs.m[8]);
    const float r_s = type == RLOA_SHAPE_BOX ? 0.f : dim.x;
    if (near_o) {   // obstacle sphere
        V3 x = ps;
        if (type == RLOA_SHAPE_CAPSULE) {
            const float tz = fminf(fmaxf(dot(obstacle - ps, ax), -dim.y), dim.y);
            x = fma3(tz, ax, ps);
        } else if (type == RLOA_SHAPE_BOX) {
            V3 loc = mulT(Rs, obstacle - ps);
            loc = v3(fminf(fmaxf(loc.x, -dim.x), dim.x), fminf(fmaxf(loc.y, -dim.y), dim.y), fminf(fmaxf(loc.z, -dim.z), dim.z));
            x = ps + mul(Rs, loc);
        }
        add(l, x, r_s, obstacle, M.obstacle_radius);
    }
    if (type != RLOA_SHAPE_BOX && near_t) cube_contact(type, l, dim, ps, ax, target, th, M_contact_thr, add);
}

constexpr int kContactRec = kMaxDof + 1;      // floats per stored contact row: J[kMaxDof] | signed distance

// The distance pass of a step records its contact CANDIDATES (link, normal, point on the link, distance) while it walks the
// shapes and writes their Jacobian rows afterwards, in one walk over the links shared by all candidates: the world axis and
// pivot of a joint are formed once and serve every candidate whose chain holds the link.  Arms without a candidate - nearly
// all of them, nearly always - pay nothing for the rows.
struct ContactCands {
    int n;
    int link[kMaxContacts];
    V3 nrm[kMaxContacts], pA[kMaxContacts];
    float dist[kMaxContacts];
};

__device__ __forceinline__ void contact_candidate(ContactCands& cc, int link, V3 x, float r_shape, V3 from, float r_body, float thr) {
    if (cc.n >= kMaxContacts) return;
    const V3 v = x - from;
    const float L = sqrtf(dot(v, v));
    if (L < 1e-9f) return;
    const float d = L - r_shape - r_body;
    if (!(d < thr)) return;
    const V3 n = (1.f / L) * v;
    cc.link[cc.n] = link;
    cc.nrm[cc.n] = n;
    cc.pA[cc.n] = x - r_shape * n;
    cc.dist[cc.n] = d;
    cc.n++;
}

__device__ __noinline__ void contact_rows_from_candidates(const ModelDev& M, const float* Rw, const float* pw, const ContactCands& cc,
                                                          float* __restrict__ crow) {
    unsigned anc[kMaxContacts], any = 0;
#pragma unroll
    for (int c = 0; c < kMaxContacts; c++) {
        anc[c] = 0;
        if (c < cc.n) {
            for (int k = cc.link[c]; k >= 0; k = M.parent[k]) anc[c] |= 1u << k;
            for (int a = 0; a < kMaxDof; a++) crow[c * kContactRec + a] = 0.f;
            crow[c * kContactRec + kMaxDof] = cc.dist[c];
        }
        any |= anc[c];
    }
    const int nl = M.nl;
    for (int k = 0; k < nl; k++) {
        const int di = M.dofidx[k];
        if (di < 0 || !((any >> k) & 1u)) continue;
        M3 Rk;
#pragma unroll
        for (int e = 0; e < 9; e++) Rk.m[e] = Rw[k * 9 + e];
        const V3 aw = mul(Rk, v3(M.axis[k][0], M.axis[k][1], M.axis[k][2]));
        const V3 pivot = ld3(pw + k * 3) - mul(Rk, v3(M.d[k][0], M.d[k][1], M.d[k][2]));
        const bool prismatic = M.jtype[k] == RLOA_JOINT_PRISMATIC;
#pragma unroll
        for (int c = 0; c < kMaxContacts; c++)
            if ((anc[c] >> k) & 1u)
                crow[c * kContactRec + di] = prismatic ? dot(cc.nrm[c], aw) : dot(cc.nrm[c], cross(aw, cc.pA[c] - pivot));
    }
}


struct ObsOut {
    float ee_target;     // closest distance end-effector link <-> target cube (10 when no shape)
    bool hit;            // any link <-> obstacle distance < obstacle_threshold
    int ncontacts;       // contact rows of this pose written to crow (want_near): the NEXT step's collision phase
    V3 ee_pos;           // COM of the end-effector link (getLinkState()[0])
};

// FK of the whole tree for one arm (q read from the SoA column qcol[i * N]) followed by the distance
// queries.  Rw / pw: per-thread scratch [NLMAX][9] / [NLMAX][3].  link_dist (optional, [nl] with stride
// ld_stride) receives the per-link minimum over the link's shapes (diagnostics path).
template <int NLMAX, bool GJK>
__device__ __forceinline__ ObsOut fk_and_distances(const ModelDev& M, const float* __restrict__ qcol, int N,
                                                   V3 obstacle, V3 target, float obstacle_thr, bool want_dist,
                                                   float* __restrict__ link_dist, int ld_stride, float contact_thr = 0.f,
                                                   bool want_near = false, float* __restrict__ crow = nullptr) {
    float Rw[NLMAX * 9], pw[NLMAX * 3];
    const int nl = M.nl;
    // every joint coordinate is requested before the serial walk: one memory round trip instead of one per link
    float qv[NLMAX];
#pragma unroll
    for (int i = 0; i < NLMAX; i++) qv[i] = i < nl ? qcol[(size_t)i * N] : 0.f;
    for (int i = 0; i < nl; i++) {
        const int par = M.parent[i];
        M3 Rl;
        V3 pl, dd;
        joint_transform(M, i, qv[i], Rl, pl, dd);
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
    ObsOut o;
    o.ee_pos = ld3(pw + M.ee_link * 3);
    o.ee_target = 10.f;
    o.hit = false;
    o.ncontacts = 0;
    if (!want_dist && !want_near) return o;
    if (link_dist != nullptr)
        for (int i = 0; i < nl; i++) link_dist[i * ld_stride] = 10.f;
    const int ns = M.ns;
    const V3 th = v3(M.target_half[0], M.target_half[1], M.target_half[2]);
    const float reach = M.obstacle_radius + fmaxf(obstacle_thr, 0.f);
    ContactCands cc;
    cc.n = 0;
    auto candidate = [&](int link, V3 x, float r_shape, V3 from, float r_body) {
        contact_candidate(cc, link, x, r_shape, from, r_body, contact_thr);
    };
    for (int s = 0; s < ns; s++) {
        const int l = M.s_link[s];
        M3 Rl, sR;
#pragma unroll
        for (int k = 0; k < 9; k++) Rl.m[k] = Rw[l * 9 + k];
        const V3 pl = ld3(pw + l * 3);
        // broad phase, exact for the decision the step consumes (d < threshold): the shape lies inside its bounding
        // sphere (centre given in the link frame), so a sphere that far from the obstacle cannot be a hit
        // (SURVEY.md 7.3-4).  The diagnostics path (link_dist) and the end-effector link keep the narrow phase.
        bool far = false, near_t = false;
        if (link_dist == nullptr) {
            const V3 bc = pl + mul(Rl, v3(M.s_bs[s][0], M.s_bs[s][1], M.s_bs[s][2]));
            const V3 rel = obstacle - bc;
            const float rr = M.s_bs[s][3] + reach + (want_near ? contact_thr : 0.f);
            far = dot(rel, rel) >= rr * rr;
            if (want_near) {        // bounding sphere against the cube itself (point - box distance); refined below
                const V3 rt = bc - target;
                const V3 ex = v3(fmaxf(fabsf(rt.x) - th.x, 0.f), fmaxf(fabsf(rt.y) - th.y, 0.f), fmaxf(fabsf(rt.z) - th.z, 0.f));
                const float rc = M.s_bs[s][3] + contact_thr;
                near_t = dot(ex, ex) < rc * rc;
            }
            if (!want_dist && far && !near_t) continue;
            if (want_dist && far && l != M.ee_link && !near_t) continue;
        }
        const V3 sp = v3(M.s_p[s][0], M.s_p[s][1], M.s_p[s][2]);
        const V3 dim = v3(M.s_dim[s][0], M.s_dim[s][1], M.s_dim[s][2]);
        const int type = M.s_type[s];
        const V3 ps = pl + mul(Rl, sp);
        if (far && !(want_dist && l == M.ee_link)) {
            // here only because the cube is close: the pair needs the shape's position and axis, not its whole frame
            if (type == RLOA_SHAPE_SPHERE || type == RLOA_SHAPE_CAPSULE)
                cube_contact(type, l, dim, ps, mul(Rl, v3(M.s_R[s][2], M.s_R[s][5], M.s_R[s][8])), target, th, contact_thr, candidate);
            continue;
        }
#pragma unroll
        for (int k = 0; k < 9; k++) sR.m[k] = M.s_R[s][k];
        const M3 Rs = mul(Rl, sR);               // world <- shape
        const V3 o_s = mulT(Rs, obstacle - ps);  // obstacle centre in the shape frame
        if (want_near && (near_t || !far))
            // contact candidates of THIS pose for the next step's collision phase (stepSimulation detects collisions on the
            // pose it starts from): found while the world frames are at hand, so that the step does not redo the kinematics
            shape_contacts(M, s, Rs, ps, obstacle, target, th, !far, near_t, contact_thr, candidate);
        float dist;
        if (far) {
            dist = reach + (want_near ? contact_thr : 0.f);     // not a hit and not a contact candidate
        } else if (type == RLOA_SHAPE_SPHERE) {
            dist = sqrtf(dot(o_s, o_s)) - dim.x;
        } else if (type == RLOA_SHAPE_CAPSULE) {
            const float tz = fminf(fmaxf(o_s.z, -dim.y), dim.y);   // closest point of the local-z segment
            const float ez = o_s.z - tz;
            dist = sqrtf(fmaf(o_s.x, o_s.x, fmaf(o_s.y, o_s.y, ez * ez))) - dim.x;
        } else if (!GJK || type == RLOA_SHAPE_BOX) {
            dist = point_box_signed(o_s, dim);
        } else {
            const GjkShape A{M.verts + M.s_v0[s], M.s_vn[s], dim, Rs, ps};
            dist = gjk_distance(A, obstacle, v3(0.f, 0.f, 0.f)) - dim.x;
        }
        const float d_obst = dist - M.obstacle_radius;
        if (!want_dist) continue;
        o.hit = o.hit || (d_obst < obstacle_thr);
        if (link_dist != nullptr) link_dist[l * ld_stride] = fminf(link_dist[l * ld_stride], d_obst);
        if (l == M.ee_link) {                    // vs the axis-aligned target cube, in the cube frame
            const V3 c_t = ps - target;
            float d_tgt = 10.f;
            if (type == RLOA_SHAPE_SPHERE) {
                d_tgt = point_box_signed(c_t, th) - dim.x;
            } else if (type == RLOA_SHAPE_CAPSULE) {
                const V3 axw = v3(Rs.m[2], Rs.m[5], Rs.m[8]);
                d_tgt = segment_box(c_t - dim.y * axw, c_t + dim.y * axw, th) - dim.x;
            } else if (GJK) {                    // box or hull vs the cube: GJK (instantiated only for such models)
                const bool box = type == RLOA_SHAPE_BOX;
                const GjkShape A{box ? nullptr : M.verts + M.s_v0[s], M.s_vn[s], dim, Rs, ps};
                d_tgt = gjk_distance(A, target, th) - (box ? 0.f : dim.x);
            }
            o.ee_target = fminf(o.ee_target, d_tgt);
        }
    }
    if (want_near) {
        o.ncontacts = cc.n;
        if (cc.n > 0) contact_rows_from_candidates(M, Rw, pw, cc, crow);
    }
    return o;
}

// ------------------------------------------------------------------------------------------------
// Kernel 1 body: kinematics, bias forces, articulated-body factorisation and free accelerations of
// one arm.  Writes the factor records F and the free velocity qs = clamp(qd + dt * qdd).
// ------------------------------------------------------------------------------------------------
template <int NLMAX>
__device__ __forceinline__ void arm_dynamics(const ModelDev& M, const SimArrays& S, int env) {
    const int N = S.n_envs, nl = M.nl;
    float L[NLMAX * kLRec];
    float slot[kMaxSlots * 27];          // forward sweeps use the first 18 / 6 floats of a slot
    // all joint coordinates are requested before the serial sweep starts (one memory round trip instead of one
    // per link); the sweep then reads them from L1-resident local memory
    float qv[NLMAX], qdv[NLMAX];
#pragma unroll
    for (int i = 0; i < NLMAX; i++) {
        qv[i] = i < nl ? S.q[(size_t)i * N + env] : 0.f;
        qdv[i] = i < nl ? S.qd[(size_t)i * N + env] : 0.f;
    }

    // ---- root -> leaves: poses, velocities, Coriolis terms, zero-acceleration forces ----
    {
        M3 R;
        V3 p = v3(0.f, 0.f, 0.f), w = p, v = p;
#pragma unroll
        for (int k = 0; k < 9; k++) R.m[k] = 0.f;
        const V3 g = v3(M.gravity[0], M.gravity[1], M.gravity[2]);
        const float ka = M.ang_damp, kl = M.lin_damp;
        for (int i = 0; i < nl; i++) {
            const int src = M.fwsrc[i];
            M3 Rp = R;
            V3 pp = p, wp = w, vp = v;
            if (src >= 2) {
                const float* sl = slot + (src - 2) * 27;
#pragma unroll
                for (int k = 0; k < 9; k++) Rp.m[k] = sl[k];
                pp = ld3(sl + 9); wp = ld3(sl + 12); vp = ld3(sl + 15);
            }
            const float q = qv[i], qd = qdv[i];
            M3 Rl;
            V3 pl, dd;
            joint_transform(M, i, q, Rl, pl, dd);
            V3 r = v3(0.f, 0.f, 0.f);
            if (src == 0) {
                R = Rl;
                p = pl;
                wp = vp = v3(0.f, 0.f, 0.f);
            } else {
                r = mul(Rp, pl);
                p = pp + r;
                R = mul(Rp, Rl);
            }
            const int jt = M.jtype[i];
            V3 s_a = v3(0.f, 0.f, 0.f), s_l = s_a;
            {
                const V3 aw = mul(R, v3(M.axis[i][0], M.axis[i][1], M.axis[i][2]));
                if (jt == RLOA_JOINT_REVOLUTE) {
                    s_a = aw;
                    s_l = cross(aw, mul(R, dd));      // Bullet m_bottomVec = axis x dVector
                } else if (jt == RLOA_JOINT_PRISMATIC) {
                    s_l = aw;
                }
            }
            const V3 vj_a = qd * s_a, vj_l = qd * s_l;
            w = wp + vj_a;
            v = vp + cross(wp, r) + vj_l;             // linear velocity of the link COM
            // Coriolis term (spatVel x spatJointVel) and zero-acceleration force (gyroscopic, drag, gravity)
            const V3 c_a = cross(w, vj_a);
            const V3 c_l = cross(w, vj_l) + cross(v, vj_a);
            const V3 I = v3(M.inertia[i][0], M.inertia[i][1], M.inertia[i][2]);
            const float mass = M.mass[i];
            const V3 wl = mulT(R, w);
            const V3 Iw = mul(R, v3(I.x * wl.x, I.y * wl.y, I.z * wl.z));
            const float wn = sqrtf(dot(w, w)), vn = sqrtf(dot(v, v));
            const V3 pA_a = fma3(fmaf(ka, wn, ka), Iw, cross(w, Iw));
            const V3 pA_l = mass * (cross(w, v) - g + fmaf(kl, vn, kl) * v);
            float* Li = L + i * kLRec;
            st3(Li, s_a); st3(Li + 3, s_l); st3(Li + 6, r); st3(Li + 9, c_a); st3(Li + 12, c_l);
            st3(Li + 15, pA_a); st3(Li + 18, pA_l);
            Li[21] = fmaf(R.m[0] * R.m[0], I.x, fmaf(R.m[1] * R.m[1], I.y, R.m[2] * R.m[2] * I.z));
            Li[22] = fmaf(R.m[0] * R.m[3], I.x, fmaf(R.m[1] * R.m[4], I.y, R.m[2] * R.m[5] * I.z));
            Li[23] = fmaf(R.m[0] * R.m[6], I.x, fmaf(R.m[1] * R.m[7], I.y, R.m[2] * R.m[8] * I.z));
            Li[24] = fmaf(R.m[3] * R.m[3], I.x, fmaf(R.m[4] * R.m[4], I.y, R.m[5] * R.m[5] * I.z));
            Li[25] = fmaf(R.m[3] * R.m[6], I.x, fmaf(R.m[4] * R.m[7], I.y, R.m[5] * R.m[8] * I.z));
            Li[26] = fmaf(R.m[6] * R.m[6], I.x, fmaf(R.m[7] * R.m[7], I.y, R.m[8] * R.m[8] * I.z));
            const int sv = M.fwsave[i];
            if (sv >= 0) {
                float* sl = slot + sv * 27;
#pragma unroll
                for (int k = 0; k < 9; k++) sl[k] = R.m[k];
                st3(sl + 9, p); st3(sl + 12, w); st3(sl + 15, v);
            }
        }
    }

    // ---- leaves -> root: articulated inertias (A ang-ang, B ang-lin, C lin-lin) and bias forces ----
    {
        float run[27];
#pragma unroll
        for (int k = 0; k < 27; k++) run[k] = 0.f;
        const int nslots = M.nslots;
        for (int k = 0; k < nslots * 27; k++) slot[k] = 0.f;
        for (int i = nl - 1; i >= 0; i--) {
            float* Li = L + i * kLRec;
            const V3 s_a = ld3(Li), s_l = ld3(Li + 3), r = ld3(Li + 6), c_a = ld3(Li + 9), c_l = ld3(Li + 12);
            V3 pA_a = ld3(Li + 15), pA_l = ld3(Li + 18);
            S3 A{Li[21], Li[22], Li[23], Li[24], Li[25], Li[26]};
            const float mass = M.mass[i];
            const float tau = -M.damping[i] * qdv[i];       // explicit joint damping (PhysicsServerCommandProcessor)
            S3 C{mass, 0.f, 0.f, mass, 0.f, mass};
            M3 B;
#pragma unroll
            for (int k = 0; k < 9; k++) B.m[k] = 0.f;
            const int bsrc = M.bwsrc[i];
            if (bsrc & 1) {
                A.xx += run[0]; A.xy += run[1]; A.xz += run[2]; A.yy += run[3]; A.yz += run[4]; A.zz += run[5];
#pragma unroll
                for (int k = 0; k < 9; k++) B.m[k] += run[6 + k];
                C.xx += run[15]; C.xy += run[16]; C.xz += run[17]; C.yy += run[18]; C.yz += run[19]; C.zz += run[20];
                pA_a = pA_a + v3(run[21], run[22], run[23]);
                pA_l = pA_l + v3(run[24], run[25], run[26]);
            }
            if (bsrc >> 1) {
                const float* G = slot + ((bsrc >> 1) - 1) * 27;
                A.xx += G[0]; A.xy += G[1]; A.xz += G[2]; A.yy += G[3]; A.yz += G[4]; A.zz += G[5];
#pragma unroll
                for (int k = 0; k < 9; k++) B.m[k] += G[6 + k];
                C.xx += G[15]; C.xy += G[16]; C.xz += G[17]; C.yy += G[18]; C.yz += G[19]; C.zz += G[20];
                pA_a = pA_a + v3(G[21], G[22], G[23]);
                pA_l = pA_l + v3(G[24], G[25], G[26]);
            }
            const V3 h_a = mul(A, s_a) + mul(B, s_l);
            const V3 h_l = mulT(B, s_a) + mul(C, s_l);
            const float D = dot(s_a, h_a) + dot(s_l, h_l);
            const float invD = (M.jtype[i] != RLOA_JOINT_FIXED) ? 1.f / D : 0.f;
            const float u = tau - (dot(s_a, pA_a) + dot(s_l, pA_l)) - (dot(c_a, h_a) + dot(c_l, h_l));
            Li[27] = u;         // h, 1/D, s and r are read back from the factor record F by the third sweep
            {
                float* Fi = S.F + (size_t)i * kFRec * N + env;
                Fi[0] = s_a.x; Fi[(size_t)N] = s_a.y; Fi[(size_t)2 * N] = s_a.z;
                Fi[(size_t)3 * N] = s_l.x; Fi[(size_t)4 * N] = s_l.y; Fi[(size_t)5 * N] = s_l.z;
                Fi[(size_t)6 * N] = h_a.x; Fi[(size_t)7 * N] = h_a.y; Fi[(size_t)8 * N] = h_a.z;
                Fi[(size_t)9 * N] = h_l.x; Fi[(size_t)10 * N] = h_l.y; Fi[(size_t)11 * N] = h_l.z;
                Fi[(size_t)12 * N] = r.x; Fi[(size_t)13 * N] = r.y; Fi[(size_t)14 * N] = r.z;
                Fi[(size_t)15 * N] = invD;
            }
            const int dst = M.bwdst[i];
            if (dst < 0) continue;
            // Ia = IA - h h^T / D ; pa = pA + IA c + h u / D
            const V3 ha = invD * h_a, hl = invD * h_l;
            const S3 A1{fmaf(-ha.x, h_a.x, A.xx), fmaf(-ha.x, h_a.y, A.xy), fmaf(-ha.x, h_a.z, A.xz),
                        fmaf(-ha.y, h_a.y, A.yy), fmaf(-ha.y, h_a.z, A.yz), fmaf(-ha.z, h_a.z, A.zz)};
            const S3 C1{fmaf(-hl.x, h_l.x, C.xx), fmaf(-hl.x, h_l.y, C.xy), fmaf(-hl.x, h_l.z, C.xz),
                        fmaf(-hl.y, h_l.y, C.yy), fmaf(-hl.y, h_l.z, C.yz), fmaf(-hl.z, h_l.z, C.zz)};
            M3 B1;
            B1.m[0] = fmaf(-ha.x, h_l.x, B.m[0]); B1.m[1] = fmaf(-ha.x, h_l.y, B.m[1]); B1.m[2] = fmaf(-ha.x, h_l.z, B.m[2]);
            B1.m[3] = fmaf(-ha.y, h_l.x, B.m[3]); B1.m[4] = fmaf(-ha.y, h_l.y, B.m[4]); B1.m[5] = fmaf(-ha.y, h_l.z, B.m[5]);
            B1.m[6] = fmaf(-ha.z, h_l.x, B.m[6]); B1.m[7] = fmaf(-ha.z, h_l.y, B.m[7]); B1.m[8] = fmaf(-ha.z, h_l.z, B.m[8]);
            const float ud = u * invD;
            const V3 n = pA_a + mul(A, c_a) + mul(B, c_l) + ud * h_a;
            const V3 f = pA_l + mulT(B, c_a) + mul(C, c_l) + ud * h_l;
            // shift to the parent's COM: C'' = C ; B'' = B + [r]x C ; A'' = A - B [r]x + [r]x B''^T
            const V3 c0 = cross(r, v3(C1.xx, C1.xy, C1.xz)), c1 = cross(r, v3(C1.xy, C1.yy, C1.yz)),
                     c2 = cross(r, v3(C1.xz, C1.yz, C1.zz));                    // columns of [r]x C
            M3 B2;
            B2.m[0] = B1.m[0] + c0.x; B2.m[1] = B1.m[1] + c1.x; B2.m[2] = B1.m[2] + c2.x;
            B2.m[3] = B1.m[3] + c0.y; B2.m[4] = B1.m[4] + c1.y; B2.m[5] = B1.m[5] + c2.y;
            B2.m[6] = B1.m[6] + c0.z; B2.m[7] = B1.m[7] + c1.z; B2.m[8] = B1.m[8] + c2.z;
            const V3 br0 = cross(v3(B1.m[0], B1.m[1], B1.m[2]), r), br1 = cross(v3(B1.m[3], B1.m[4], B1.m[5]), r),
                     br2 = cross(v3(B1.m[6], B1.m[7], B1.m[8]), r);             // rows of B [r]x
            const V3 rb0 = cross(r, v3(B2.m[0], B2.m[1], B2.m[2])), rb1 = cross(r, v3(B2.m[3], B2.m[4], B2.m[5])),
                     rb2 = cross(r, v3(B2.m[6], B2.m[7], B2.m[8]));             // columns of [r]x B''^T
            float T[27];
            T[0] = A1.xx - br0.x + rb0.x;
            T[1] = A1.xy - br0.y + rb1.x;
            T[2] = A1.xz - br0.z + rb2.x;
            T[3] = A1.yy - br1.y + rb1.y;
            T[4] = A1.yz - br1.z + rb2.y;
            T[5] = A1.zz - br2.z + rb2.z;
#pragma unroll
            for (int k = 0; k < 9; k++) T[6 + k] = B2.m[k];
            T[15] = C1.xx; T[16] = C1.xy; T[17] = C1.xz; T[18] = C1.yy; T[19] = C1.yz; T[20] = C1.zz;
            const V3 n2 = n + cross(r, f);
            T[21] = n2.x; T[22] = n2.y; T[23] = n2.z; T[24] = f.x; T[25] = f.y; T[26] = f.z;
            if (dst == 0) {
#pragma unroll
                for (int k = 0; k < 27; k++) run[k] = T[k];
            } else {
                float* G = slot + (dst - 1) * 27;
#pragma unroll
                for (int k = 0; k < 27; k++) G[k] += T[k];
            }
        }
    }

    // ---- root -> leaves: free accelerations -> free velocity ----
    {
        V3 a_a = v3(0.f, 0.f, 0.f), a_l = a_a;
        const float dt = M.dt, max_vel = M.max_vel;
        for (int i = 0; i < nl; i++) {
            const float* Li = L + i * kLRec;
            const int src = M.fwsrc[i];
            V3 p_a = a_a, p_l = a_l;
            if (src == 0) {
                p_a = p_l = v3(0.f, 0.f, 0.f);
            } else if (src >= 2) {
                p_a = ld3(slot + (src - 2) * 27);
                p_l = ld3(slot + (src - 2) * 27 + 3);
            }
            const float* Fi = S.F + (size_t)i * kFRec * N + env;      // written by this thread in the second sweep
            const V3 s_a = v3(Fi[0], Fi[(size_t)N], Fi[(size_t)2 * N]);
            const V3 s_l = v3(Fi[(size_t)3 * N], Fi[(size_t)4 * N], Fi[(size_t)5 * N]);
            const V3 h_a = v3(Fi[(size_t)6 * N], Fi[(size_t)7 * N], Fi[(size_t)8 * N]);
            const V3 h_l = v3(Fi[(size_t)9 * N], Fi[(size_t)10 * N], Fi[(size_t)11 * N]);
            const V3 r = v3(Fi[(size_t)12 * N], Fi[(size_t)13 * N], Fi[(size_t)14 * N]);
            const float invD = Fi[(size_t)15 * N];
            const V3 x_a = p_a, x_l = p_l + cross(p_a, r);
            const float qdd = (Li[27] - (dot(h_a, x_a) + dot(h_l, x_l))) * invD;
            a_a = fma3(qdd, s_a, x_a + ld3(Li + 9));
            a_l = fma3(qdd, s_l, x_l + ld3(Li + 12));
            const int sv = M.fwsave[i];
            if (sv >= 0) {
                st3(slot + sv * 27, a_a);
                st3(slot + sv * 27 + 3, a_l);
            }
            const float qd = qdv[i];
            const bool hasdof = M.jtype[i] != RLOA_JOINT_FIXED;
            S.qs[(size_t)i * N + env] = hasdof ? fminf(fmaxf(fmaf(dt, qdd, qd), -max_vel), max_vel) : 0.f;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel 2 body: column j of M^-1 for one arm (calcAccelerationDeltasMultiDof with a unit impulse on
// dof j): test force up the ancestors of the dof's link, accelerations down the whole tree.
// ------------------------------------------------------------------------------------------------
// F: this arm's factor records, component k of link i at F[i * link_stride + k * k_stride] (shared-memory stage)
__device__ __forceinline__ void arm_minv_column(const ModelDev& M, const SimArrays& S, int env, int j,
                                                const float* __restrict__ F, int k_stride, int link_stride) {
    const int N = S.n_envs, nl = M.nl, ndof = M.ndof;
    const int lj = M.doflink[j];
    float Y[kMaxLinks];
#pragma unroll
    for (int i = 0; i < kMaxLinks; i++) Y[i] = 0.f;
    const int ks = k_stride;
    {
        V3 zf_a = v3(0.f, 0.f, 0.f), zf_l = zf_a;
        int cur = lj;
        while (cur >= 0) {
            const float* Fi = F + cur * link_stride;
            const V3 sa = v3(Fi[0], Fi[ks], Fi[2 * ks]), sl = v3(Fi[3 * ks], Fi[4 * ks], Fi[5 * ks]);
            const V3 ha = v3(Fi[6 * ks], Fi[7 * ks], Fi[8 * ks]), hl = v3(Fi[9 * ks], Fi[10 * ks], Fi[11 * ks]);
            const V3 rr = v3(Fi[12 * ks], Fi[13 * ks], Fi[14 * ks]);
            const float iD = Fi[15 * ks];
            float y = (cur == lj ? 1.f : 0.f) - (dot(sa, zf_a) + dot(sl, zf_l));
            if (iD == 0.f) y = 0.f;
            Y[cur] = y;
            const float t = y * iD;
            const V3 f_a = fma3(t, ha, zf_a), f_l = fma3(t, hl, zf_l);
            zf_a = f_a + cross(rr, f_l);
            zf_l = f_l;
            cur = M.parent[cur];
        }
    }
    {
        V3 a_a = v3(0.f, 0.f, 0.f), a_l = a_a;
        float sv[kMaxSlots * 6];
        for (int i = 0; i < nl; i++) {
            const float* Fi = F + i * link_stride;
            const int src = M.fwsrc[i];
            V3 p_a = a_a, p_l = a_l;
            if (src == 0) {
                p_a = p_l = v3(0.f, 0.f, 0.f);
            } else if (src >= 2) {
                p_a = ld3(sv + (src - 2) * 6);
                p_l = ld3(sv + (src - 2) * 6 + 3);
            }
            const V3 rr = v3(Fi[12 * ks], Fi[13 * ks], Fi[14 * ks]);
            a_a = p_a;
            a_l = p_l + cross(p_a, rr);
            const int di = M.dofidx[i];
            if (di >= 0) {
                const V3 sa = v3(Fi[0], Fi[ks], Fi[2 * ks]), sl = v3(Fi[3 * ks], Fi[4 * ks], Fi[5 * ks]);
                const V3 ha = v3(Fi[6 * ks], Fi[7 * ks], Fi[8 * ks]), hl = v3(Fi[9 * ks], Fi[10 * ks], Fi[11 * ks]);
                const float x = (Y[i] - (dot(ha, a_a) + dot(hl, a_l))) * Fi[15 * ks];
                a_a = fma3(x, sa, a_a);
                a_l = fma3(x, sl, a_l);
                S.minv[((size_t)di * ndof + j) * N + env] = x;
            }
            const int save = M.fwsave[i];
            if (save >= 0) {
                st3(sv + save * 6, a_a);
                st3(sv + save * 6 + 3, a_l);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Contact rows against the two collidable fixed bodies of the reference (environment.py:252-255: obstacle sphere and
// target cube, useFixedBase): collision detection on the pre-step pose, one NORMAL row per (shape, body) pair closer than
// the contact breaking threshold (Bullet gContactBreakingThreshold = 0.02), restated from
// btMultiBodyConstraintSolver::setupMultiBodyContactConstraint.  Same model as oracle/bullet_restatement.c
// (find_contacts / contact_jacobian): sphere, capsule and box shapes against the sphere, sphere and capsule shapes against
// the cube; no friction rows, one point per pair.  Returns the number of contacts; J[c][d] = d(n . p_contact)/dq_d per dof.
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ float segment_box_argmin(V3 a, V3 b, V3 h) {
    V3 dir = b - a;
    float bp[8];
    int nb = 0;
    bp[nb++] = 0.f;
    bp[nb++] = 1.f;
    const float av[3] = {a.x, a.y, a.z}, dv[3] = {dir.x, dir.y, dir.z}, hv[3] = {h.x, h.y, h.z};
#pragma unroll
    for (int x = 0; x < 3; x++) {
        if (fabsf(dv[x]) < 1e-30f) continue;
        float t1 = (hv[x] - av[x]) / dv[x], t2 = (-hv[x] - av[x]) / dv[x];
        if (t1 > 0.f && t1 < 1.f) bp[nb++] = t1;
        if (t2 > 0.f && t2 < 1.f) bp[nb++] = t2;
    }
    for (int i = 1; i < nb; i++) {
        float v = bp[i];
        int j = i - 1;
        while (j >= 0 && bp[j] > v) {
            bp[j + 1] = bp[j];
            j--;
        }
        bp[j + 1] = v;
    }
    float best = 3.0e38f, tb = 0.f;
    for (int i = 0; i < nb; i++) {
        const float f = seg_box_f(a, dir, h, bp[i]);
        if (f < best) { best = f; tb = bp[i]; }
        if (i + 1 < nb) {
            float tm = 0.5f * (bp[i] + bp[i + 1]), A = 0.f, B = 0.f;
#pragma unroll
            for (int x = 0; x < 3; x++) {
                float xm = fmaf(tm, dv[x], av[x]);
                if (fabsf(xm) > hv[x]) {
                    float sg = xm > 0.f ? 1.f : -1.f;
                    A = fmaf(dv[x], dv[x], A);
                    B = fmaf(2.f * (sg * av[x] - hv[x]) * sg, dv[x], B);
                }
            }
            if (A > 0.f) {
                float ts = -B / (2.f * A);
                if (ts > bp[i] && ts < bp[i + 1]) {
                    const float f2 = seg_box_f(a, dir, h, ts);
                    if (f2 < best) { best = f2; tb = ts; }
                }
            }
        }
    }
    return tb;
}

}  // namespace rloa
