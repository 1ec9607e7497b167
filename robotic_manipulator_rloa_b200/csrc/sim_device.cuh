// Device side of the batched manipulator simulator: one warp per arm.
//
// What it replaces: the Bullet3 multibody step that runs under the reference's
// Environment.step / Environment.reset (reference environment/environment.py:264-309, 453-485) —
// btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof, calcAccelerationDeltasMultiDof,
// btMultiBodyJointMotor / btMultiBodyJointLimitConstraint rows, the PGS sweep of
// btMultiBodyConstraintSolver and stepPositionsMultiDof — plus the getClosestPoints queries of
// utils/collision_detector.py:33-61 and the reward / terminal logic of environment.py:311-371.
//
// Mapping (B200-first, not Bullet's): lane l of the warp owns link l.  All spatial quantities are kept
// in WORLD orientation with the origin at the owning link's centre of mass, so the transform between
// a link and its parent is a pure translation r = p_child - p_parent:
//   - link poses and link velocities are tree prefix products / sums -> pointer jumping over the
//     parent map with warp shuffles (ceil(log2(depth)) rounds instead of a serial chain walk);
//   - the articulated-inertia recursion folds children into parents level by level, 27 shuffled
//     floats per child (symmetric 6x6 as blocks A (ang-ang), B (ang-lin), C (lin-lin));
//   - the columns of M^-1 (Bullet's per-row "unit impulse responses") are computed with lane j
//     owning column j, per-link factors broadcast from shared memory;
//   - the projected Gauss-Seidel sweep runs in ROW space: lane k owns constraint row k and keeps
//     w_k = J_k dv; one row update is a candidate impulse on every lane, one shuffle, one FMA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rloa {

constexpr int kMaxLinks = 32;
constexpr int kMaxShapes = 32;
constexpr int kMaxDof = 16;
constexpr int kMaxChildren = 4;
constexpr int kMaxSlots = 3;
constexpr int kWarpsPerBlock = 4;
constexpr int kLinkStride = 17;          // 16 floats of per-link factors + 1 pad (bank spread)
constexpr int kMinvStride = kMaxDof + 1;
constexpr unsigned kFull = 0xffffffffu;

// fp32 device copy of rloa_model_desc plus the tree tables the kernels need
struct ModelDev {
    int nl, ns, ndof, maxdepth, nrounds, ee_link, n_obs, iters;
    float dt, inv_dt, lin_damp, ang_damp, resid_thresh, erp, max_vel, limit_max_imp;
    float gravity[3];
    float obstacle_radius;
    float target_half[3];
    int parent[kMaxLinks], jtype[kMaxLinks], depth[kMaxLinks], nch[kMaxLinks];
    int child[kMaxLinks][kMaxChildren];
    int has_limit[kMaxLinks];
    int dofidx[kMaxLinks];               // compact dof index of a movable link, -1 otherwise
    int doflink[kMaxDof];                // link of dof j
    int accsrc[kMaxLinks];               // unit-response sweep: 0 zero (root) | 1 running | 2+k slot k
    int accsave[kMaxLinks];              // slot to save into after the link, -1 none
    int lvl_maxch[kMaxLinks + 1];        // max #children of any link at depth d-1 having children at depth d
    unsigned anc_mask[kMaxLinks];        // bit i set: link i is on the path link..root (inclusive)
    float E0T[kMaxLinks][9];             // parent COM frame <- child COM frame at q = 0 (base folded into roots)
    float e[kMaxLinks][3];               // parent COM -> pivot in the parent frame (base folded into roots)
    float d[kMaxLinks][3], axis[kMaxLinks][3];
    float mass[kMaxLinks], inertia[kMaxLinks][3], damping[kMaxLinks], lower[kMaxLinks], upper[kMaxLinks];
    int s_link[kMaxShapes], s_type[kMaxShapes];
    float s_R[kMaxShapes][9], s_p[kMaxShapes][3], s_dim[kMaxShapes][3];
};

__host__ __device__ inline int sim_smem_floats_per_warp(int nl) {
    // link factors | link ints | Y | Minv | rows (6 arrays of 32)
    return nl * kLinkStride + nl + nl * kMaxDof + kMaxDof * kMinvStride + 6 * 32;
}

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
__device__ __forceinline__ float shf(float v, int src) { return __shfl_sync(kFull, v, src); }
__device__ __forceinline__ int shi(int v, int src) { return __shfl_sync(kFull, v, src); }
__device__ __forceinline__ V3 sh3(V3 v, int src) { return v3(shf(v.x, src), shf(v.y, src), shf(v.z, src)); }

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
// Forward kinematics for the whole tree: world <- link COM frame rotation R and COM position p of
// the lane's link.  Local transforms come from the model (Bullet: btQuaternion(axis,-q) *
// zeroRotParentToThis, rVector = E e + d) and are composed by pointer jumping.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lane_fk(const ModelDev* __restrict__ M, int link, bool valid, float q, M3& R, V3& p) {
    V3 ax = v3(0.f, 0.f, 1.f), dd = v3(0.f, 0.f, 0.f), ee = dd;
    M3 E0T;
#pragma unroll
    for (int k = 0; k < 9; k++) E0T.m[k] = (k % 4 == 0) ? 1.f : 0.f;
    int jt = RLOA_JOINT_FIXED, anc = -1;
    if (valid) {
#pragma unroll
        for (int k = 0; k < 9; k++) E0T.m[k] = __ldg(&M->E0T[link][k]);
        ax = v3(__ldg(&M->axis[link][0]), __ldg(&M->axis[link][1]), __ldg(&M->axis[link][2]));
        dd = v3(__ldg(&M->d[link][0]), __ldg(&M->d[link][1]), __ldg(&M->d[link][2]));
        ee = v3(__ldg(&M->e[link][0]), __ldg(&M->e[link][1]), __ldg(&M->e[link][2]));
        jt = __ldg(&M->jtype[link]);
        anc = __ldg(&M->parent[link]);
    }
    if (jt == RLOA_JOINT_REVOLUTE) {
        float s, c;
        sincosf(q, &s, &c);
        float t = 1.f - c;
        M3 Rq;   // Rodrigues(axis, +q)
        Rq.m[0] = fmaf(t * ax.x, ax.x, c);          Rq.m[1] = fmaf(t * ax.x, ax.y, -s * ax.z);  Rq.m[2] = fmaf(t * ax.x, ax.z, s * ax.y);
        Rq.m[3] = fmaf(t * ax.y, ax.x, s * ax.z);   Rq.m[4] = fmaf(t * ax.y, ax.y, c);          Rq.m[5] = fmaf(t * ax.y, ax.z, -s * ax.x);
        Rq.m[6] = fmaf(t * ax.z, ax.x, -s * ax.y);  Rq.m[7] = fmaf(t * ax.z, ax.y, s * ax.x);   Rq.m[8] = fmaf(t * ax.z, ax.z, c);
        R = mul(E0T, Rq);
    } else {
        R = E0T;
        if (jt == RLOA_JOINT_PRISMATIC) dd = fma3(q, ax, dd);
    }
    p = ee + mul(R, dd);
    const int rounds = M->nrounds;
    for (int k = 0; k < rounds; k++) {
        int src = anc < 0 ? (int)(threadIdx.x & 31) : anc;
        M3 Ra;
#pragma unroll
        for (int i = 0; i < 9; i++) Ra.m[i] = shf(R.m[i], src);
        V3 pa = sh3(p, src);
        int aa = shi(anc, src);
        if (anc >= 0) {
            p = pa + mul(Ra, p);
            R = mul(Ra, R);
            anc = aa;
        }
    }
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

struct ObsOut {
    float ee_target;     // closest distance end-effector link <-> target cube (10 when no shape)
    bool hit;            // any link <-> obstacle distance < obstacle_threshold
    V3 ee_pos;           // COM of the end-effector link (getLinkState()[0])
    float link_dist;     // lane = link: min distance link <-> obstacle (10 when no shape)
};

// lane = link on entry (R, p of the lane's link); lanes are re-used as lane = shape for the queries
__device__ __forceinline__ ObsOut lane_distances(const ModelDev* __restrict__ M, int lane, const M3& R, V3 p,
                                                 V3 obstacle, V3 target, float obstacle_thr, bool want_link_dist) {
    const int ns = M->ns;
    const bool sv = lane < ns;
    const int l = sv ? __ldg(&M->s_link[lane]) : 0;
    M3 Rl;
#pragma unroll
    for (int i = 0; i < 9; i++) Rl.m[i] = shf(R.m[i], l);
    V3 pl = sh3(p, l);
    float d_obst = 10.f, d_tgt = 10.f;
    if (sv) {
        M3 sR;
#pragma unroll
        for (int i = 0; i < 9; i++) sR.m[i] = __ldg(&M->s_R[lane][i]);
        V3 sp = v3(__ldg(&M->s_p[lane][0]), __ldg(&M->s_p[lane][1]), __ldg(&M->s_p[lane][2]));
        V3 dim = v3(__ldg(&M->s_dim[lane][0]), __ldg(&M->s_dim[lane][1]), __ldg(&M->s_dim[lane][2]));
        const int type = __ldg(&M->s_type[lane]);
        M3 Rs = mul(Rl, sR);               // world <- shape
        V3 ps = pl + mul(Rl, sp);
        V3 o_s = mulT(Rs, obstacle - ps);  // obstacle centre in the shape frame
        float dist;
        if (type == RLOA_SHAPE_SPHERE) {
            dist = sqrtf(dot(o_s, o_s)) - dim.x;
        } else if (type == RLOA_SHAPE_CAPSULE) {
            float tz = fminf(fmaxf(o_s.z, -dim.y), dim.y);   // closest point of the local-z segment
            float ez = o_s.z - tz;
            dist = sqrtf(fmaf(o_s.x, o_s.x, fmaf(o_s.y, o_s.y, ez * ez))) - dim.x;
        } else {
            dist = point_box_signed(o_s, dim);
        }
        d_obst = dist - M->obstacle_radius;
        if (l == M->ee_link) {             // vs the axis-aligned target cube, in the cube frame
            V3 th = v3(M->target_half[0], M->target_half[1], M->target_half[2]);
            V3 c_t = ps - target;
            if (type == RLOA_SHAPE_SPHERE) {
                d_tgt = point_box_signed(c_t, th) - dim.x;
            } else if (type == RLOA_SHAPE_CAPSULE) {
                V3 axw = v3(Rs.m[2], Rs.m[5], Rs.m[8]);
                d_tgt = segment_box(c_t - dim.y * axw, c_t + dim.y * axw, th) - dim.x;
            }                               // box end-effector shapes: no narrow phase -> saturate (10)
        }
    }
    ObsOut o;
    o.hit = __any_sync(kFull, d_obst < obstacle_thr);
    float t = d_tgt;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t = fminf(t, __shfl_xor_sync(kFull, t, off));
    o.ee_target = t;
    o.ee_pos = sh3(p, M->ee_link);
    o.link_dist = 10.f;
    if (want_link_dist) {                   // diagnostics path only: per-link minimum over its shapes
        for (int s = 0; s < ns; s++) {
            float ds = shf(d_obst, s);
            int ls = shi(l, s);
            if (ls == lane) o.link_dist = fminf(o.link_dist, ds);
        }
    }
    return o;
}

// per-lane motor settings (btMultiBodyJointMotor state left behind by setJointMotorControl2; kd == 1)
struct Motor {
    float kp, tpos, tvel, maximp;
};

// ------------------------------------------------------------------------------------------------
// One stepSimulation for the warp's arm.  lane = link; q, qd in/out; sm = this warp's shared slice.
// Returns the number of PGS iterations used.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_substep(const ModelDev* __restrict__ M, float* __restrict__ sm, int lane,
                                            float& q, float& qd, const Motor& mot) {
    const int nl = M->nl;
    const bool valid = lane < nl;
    const int link = valid ? lane : 0;
    const int maxdepth = M->maxdepth;

    float* sm_link = sm;                                   // [nl][kLinkStride]
    int* sm_li = reinterpret_cast<int*>(sm_link + nl * kLinkStride);   // [nl] packed tree ints
    float* sm_Y = reinterpret_cast<float*>(sm_li + nl);    // [nl][kMaxDof]
    float* sm_Minv = sm_Y + nl * kMaxDof;                  // [kMaxDof][kMinvStride]
    float* sm_row = sm_Minv + kMaxDof * kMinvStride;       // 6 x [32]

    // ---- kinematics ----
    M3 R;
    V3 p;
    lane_fk(M, link, valid, q, R, p);
    int parent = -1, jt = RLOA_JOINT_FIXED, depth = -1, nch = 0, mydof = -1;
    int ch[kMaxChildren];
#pragma unroll
    for (int k = 0; k < kMaxChildren; k++) ch[k] = 0;
    float mass = 0.f, damping = 0.f;
    V3 I = v3(0.f, 0.f, 0.f), s_a = I, s_l = I;
    if (valid) {
        parent = __ldg(&M->parent[link]);
        jt = __ldg(&M->jtype[link]);
        depth = __ldg(&M->depth[link]);
        nch = __ldg(&M->nch[link]);
#pragma unroll
        for (int k = 0; k < kMaxChildren; k++) ch[k] = __ldg(&M->child[link][k]);
        mydof = __ldg(&M->dofidx[link]);
        mass = __ldg(&M->mass[link]);
        damping = __ldg(&M->damping[link]);
        I = v3(__ldg(&M->inertia[link][0]), __ldg(&M->inertia[link][1]), __ldg(&M->inertia[link][2]));
        V3 ax = v3(__ldg(&M->axis[link][0]), __ldg(&M->axis[link][1]), __ldg(&M->axis[link][2]));
        V3 aw = mul(R, ax);
        if (jt == RLOA_JOINT_REVOLUTE) {
            V3 dd = v3(__ldg(&M->d[link][0]), __ldg(&M->d[link][1]), __ldg(&M->d[link][2]));
            s_a = aw;
            s_l = cross(aw, mul(R, dd));      // Bullet m_bottomVec = axis x dVector
        } else if (jt == RLOA_JOINT_PRISMATIC) {
            s_l = aw;
        }
    }
    const bool hasdof = jt != RLOA_JOINT_FIXED;
    V3 r = v3(0.f, 0.f, 0.f);
    {
        V3 pp = sh3(p, parent < 0 ? lane : parent);
        if (parent >= 0) r = p - pp;
    }

    // ---- link velocities: tree prefix sum of joint twists taken at the world origin ----
    V3 vj_a = qd * s_a, vj_l = qd * s_l;
    V3 w = vj_a;                               // angular
    V3 v0 = vj_l - cross(vj_a, p);             // linear, at the world origin
    {
        int anc = parent;
        const int rounds = M->nrounds;
        for (int k = 0; k < rounds; k++) {
            int src = anc < 0 ? lane : anc;
            V3 wa = sh3(w, src), va = sh3(v0, src);
            int aa = shi(anc, src);
            if (anc >= 0) {
                w = w + wa;
                v0 = v0 + va;
                anc = aa;
            }
        }
    }
    V3 v = v0 + cross(w, p);                   // linear velocity of the link COM

    // Coriolis term (spatVel x spatJointVel) and zero-acceleration force (gyroscopic, drag, gravity)
    V3 c_a = cross(w, vj_a);
    V3 c_l = cross(w, vj_l) + cross(v, vj_a);
    V3 pA_a, pA_l;
    S3 A, C;
    M3 B;
    {
        V3 wl = mulT(R, w);
        V3 Iw = mul(R, v3(I.x * wl.x, I.y * wl.y, I.z * wl.z));
        float wn = sqrtf(dot(w, w)), vn = sqrtf(dot(v, v));
        float ka = M->ang_damp, kl = M->lin_damp;
        V3 g = v3(M->gravity[0], M->gravity[1], M->gravity[2]);
        pA_a = fma3(fmaf(ka, wn, ka), Iw, cross(w, Iw));
        pA_l = mass * (cross(w, v) - g + fmaf(kl, vn, kl) * v);
        A.xx = fmaf(R.m[0] * R.m[0], I.x, fmaf(R.m[1] * R.m[1], I.y, R.m[2] * R.m[2] * I.z));
        A.xy = fmaf(R.m[0] * R.m[3], I.x, fmaf(R.m[1] * R.m[4], I.y, R.m[2] * R.m[5] * I.z));
        A.xz = fmaf(R.m[0] * R.m[6], I.x, fmaf(R.m[1] * R.m[7], I.y, R.m[2] * R.m[8] * I.z));
        A.yy = fmaf(R.m[3] * R.m[3], I.x, fmaf(R.m[4] * R.m[4], I.y, R.m[5] * R.m[5] * I.z));
        A.yz = fmaf(R.m[3] * R.m[6], I.x, fmaf(R.m[4] * R.m[7], I.y, R.m[5] * R.m[8] * I.z));
        A.zz = fmaf(R.m[6] * R.m[6], I.x, fmaf(R.m[7] * R.m[7], I.y, R.m[8] * R.m[8] * I.z));
        C = S3{mass, 0.f, 0.f, mass, 0.f, mass};
#pragma unroll
        for (int i = 0; i < 9; i++) B.m[i] = 0.f;
    }

    // ---- articulated inertias, leaves -> root, one tree level per iteration ----
    V3 h_a = v3(0.f, 0.f, 0.f), h_l = h_a;
    float invD = 0.f, u = 0.f;
    const float tau = -damping * qd;           // explicit joint damping (PhysicsServerCommandProcessor)
    for (int dlev = maxdepth; dlev >= 0; dlev--) {
        // factors of the lane's own link (final once all its children have been folded in)
        h_a = mul(A, s_a) + mul(B, s_l);
        h_l = mulT(B, s_a) + mul(C, s_l);
        float D = dot(s_a, h_a) + dot(s_l, h_l);
        invD = hasdof ? 1.f / D : 0.f;
        u = tau - (dot(s_a, pA_a) + dot(s_l, pA_l)) - (dot(c_a, h_a) + dot(c_l, h_l));
        if (dlev == 0) break;
        // Ia = IA - h h^T / D ; pa = pA + IA c + h u / D
        V3 ha = invD * h_a, hl = invD * h_l;
        S3 A1{fmaf(-ha.x, h_a.x, A.xx), fmaf(-ha.x, h_a.y, A.xy), fmaf(-ha.x, h_a.z, A.xz),
              fmaf(-ha.y, h_a.y, A.yy), fmaf(-ha.y, h_a.z, A.yz), fmaf(-ha.z, h_a.z, A.zz)};
        S3 C1{fmaf(-hl.x, h_l.x, C.xx), fmaf(-hl.x, h_l.y, C.xy), fmaf(-hl.x, h_l.z, C.xz),
              fmaf(-hl.y, h_l.y, C.yy), fmaf(-hl.y, h_l.z, C.yz), fmaf(-hl.z, h_l.z, C.zz)};
        M3 B1;
        B1.m[0] = fmaf(-ha.x, h_l.x, B.m[0]); B1.m[1] = fmaf(-ha.x, h_l.y, B.m[1]); B1.m[2] = fmaf(-ha.x, h_l.z, B.m[2]);
        B1.m[3] = fmaf(-ha.y, h_l.x, B.m[3]); B1.m[4] = fmaf(-ha.y, h_l.y, B.m[4]); B1.m[5] = fmaf(-ha.y, h_l.z, B.m[5]);
        B1.m[6] = fmaf(-ha.z, h_l.x, B.m[6]); B1.m[7] = fmaf(-ha.z, h_l.y, B.m[7]); B1.m[8] = fmaf(-ha.z, h_l.z, B.m[8]);
        float ud = u * invD;
        V3 n = pA_a + mul(A, c_a) + mul(B, c_l) + ud * h_a;
        V3 f = pA_l + mulT(B, c_a) + mul(C, c_l) + ud * h_l;
        // shift to the parent's COM: C'' = C ; B'' = B + [r]x C ; A'' = A - B [r]x + [r]x B''^T
        V3 c0 = cross(r, v3(C1.xx, C1.xy, C1.xz)), c1 = cross(r, v3(C1.xy, C1.yy, C1.yz)),
           c2 = cross(r, v3(C1.xz, C1.yz, C1.zz));                    // columns of [r]x C
        M3 B2;
        B2.m[0] = B1.m[0] + c0.x; B2.m[1] = B1.m[1] + c1.x; B2.m[2] = B1.m[2] + c2.x;
        B2.m[3] = B1.m[3] + c0.y; B2.m[4] = B1.m[4] + c1.y; B2.m[5] = B1.m[5] + c2.y;
        B2.m[6] = B1.m[6] + c0.z; B2.m[7] = B1.m[7] + c1.z; B2.m[8] = B1.m[8] + c2.z;
        V3 br0 = cross(v3(B1.m[0], B1.m[1], B1.m[2]), r), br1 = cross(v3(B1.m[3], B1.m[4], B1.m[5]), r),
           br2 = cross(v3(B1.m[6], B1.m[7], B1.m[8]), r);             // rows of B [r]x
        V3 rb0 = cross(r, v3(B2.m[0], B2.m[1], B2.m[2])), rb1 = cross(r, v3(B2.m[3], B2.m[4], B2.m[5])),
           rb2 = cross(r, v3(B2.m[6], B2.m[7], B2.m[8]));             // columns of [r]x B''^T
        float T[27];
        T[0] = A1.xx - br0.x + rb0.x;
        T[1] = A1.xy - br0.y + rb1.x;
        T[2] = A1.xz - br0.z + rb2.x;
        T[3] = A1.yy - br1.y + rb1.y;
        T[4] = A1.yz - br1.z + rb2.y;
        T[5] = A1.zz - br2.z + rb2.z;
#pragma unroll
        for (int i = 0; i < 9; i++) T[6 + i] = B2.m[i];
        T[15] = C1.xx; T[16] = C1.xy; T[17] = C1.xz; T[18] = C1.yy; T[19] = C1.yz; T[20] = C1.zz;
        V3 n2 = n + cross(r, f);
        T[21] = n2.x; T[22] = n2.y; T[23] = n2.z; T[24] = f.x; T[25] = f.y; T[26] = f.z;
        // parents one level up pull from their children
        const int maxch = M->lvl_maxch[dlev];
        const bool is_parent = depth == dlev - 1;
#pragma unroll
        for (int k = 0; k < kMaxChildren; k++) {
            if (k < maxch) {
                const bool take = is_parent && k < nch;
                const int src = take ? ch[k] : lane;
                float G[27];
#pragma unroll
                for (int i = 0; i < 27; i++) G[i] = shf(T[i], src);
                if (take) {
                    A.xx += G[0]; A.xy += G[1]; A.xz += G[2]; A.yy += G[3]; A.yz += G[4]; A.zz += G[5];
#pragma unroll
                    for (int i = 0; i < 9; i++) B.m[i] += G[6 + i];
                    C.xx += G[15]; C.xy += G[16]; C.xz += G[17]; C.yy += G[18]; C.yz += G[19]; C.zz += G[20];
                    pA_a = pA_a + v3(G[21], G[22], G[23]);
                    pA_l = pA_l + v3(G[24], G[25], G[26]);
                }
            }
        }
    }

    // ---- accelerations, root -> leaves ----
    float qdd = 0.f;
    {
        V3 a_a = v3(0.f, 0.f, 0.f), a_l = a_a;
        for (int dlev = 0; dlev <= maxdepth; dlev++) {
            const int src = parent < 0 ? lane : parent;
            V3 pa_a = sh3(a_a, src), pa_l = sh3(a_l, src);
            if (depth == dlev) {
                if (parent < 0) pa_a = pa_l = v3(0.f, 0.f, 0.f);
                V3 x_a = pa_a, x_l = pa_l + cross(pa_a, r);
                qdd = (u - (dot(h_a, x_a) + dot(h_l, x_l))) * invD;
                a_a = x_a + c_a + qdd * s_a;
                a_l = x_l + c_l + qdd * s_l;
            }
        }
    }
    const float max_vel = M->max_vel;
    float qs = hasdof ? fminf(fmaxf(fmaf(M->dt, qdd, qd), -max_vel), max_vel) : 0.f;

    // ---- columns of M^-1: lane j owns dof j (calcAccelerationDeltasMultiDof with a unit impulse) ----
    __syncwarp();
    if (valid) {
        float* L = sm_link + link * kLinkStride;
        L[0] = s_a.x; L[1] = s_a.y; L[2] = s_a.z; L[3] = s_l.x; L[4] = s_l.y; L[5] = s_l.z;
        L[6] = h_a.x; L[7] = h_a.y; L[8] = h_a.z; L[9] = h_l.x; L[10] = h_l.y; L[11] = h_l.z;
        L[12] = r.x; L[13] = r.y; L[14] = r.z; L[15] = invD;
        // packed ints: parent+1 (8b) | accsrc (4b) | accsave+1 (4b) | dofidx+1 (8b)
        sm_li[link] = (parent + 1) | (__ldg(&M->accsrc[link]) << 8) | ((__ldg(&M->accsave[link]) + 1) << 12) |
                      ((mydof + 1) << 16);
    }
    __syncwarp();
    const int ndof = M->ndof;
    const bool colv = lane < ndof;
    const int lj = colv ? __ldg(&M->doflink[lane]) : 0;
    const unsigned ancm = colv ? __ldg(&M->anc_mask[lj]) : 0u;
    {
        V3 zf_a = v3(0.f, 0.f, 0.f), zf_l = zf_a;
        int cur = colv ? lj : -1;
        for (int step = 0; step <= maxdepth; step++) {
            if (cur >= 0) {
                const float* L = sm_link + cur * kLinkStride;
                V3 sa = v3(L[0], L[1], L[2]), sl = v3(L[3], L[4], L[5]);
                V3 ha = v3(L[6], L[7], L[8]), hl = v3(L[9], L[10], L[11]);
                V3 rr = v3(L[12], L[13], L[14]);
                float iD = L[15];
                float Y = (cur == lj ? 1.f : 0.f) - (dot(sa, zf_a) + dot(sl, zf_l));
                if (iD == 0.f) Y = 0.f;
                sm_Y[cur * kMaxDof + lane] = Y;
                float t = Y * iD;
                V3 f_a = fma3(t, ha, zf_a), f_l = fma3(t, hl, zf_l);
                zf_a = f_a + cross(rr, f_l);
                zf_l = f_l;
                cur = (sm_li[cur] & 0xff) - 1;
            }
        }
    }
    __syncwarp();
    {
        V3 a_a = v3(0.f, 0.f, 0.f), a_l = a_a;
        V3 sv_a[kMaxSlots], sv_l[kMaxSlots];
#pragma unroll
        for (int k = 0; k < kMaxSlots; k++) sv_a[k] = sv_l[k] = v3(0.f, 0.f, 0.f);
        for (int i = 0; i < nl; i++) {
            const float* L = sm_link + i * kLinkStride;
            const int li = sm_li[i];
            const int src = (li >> 8) & 0xf, save = ((li >> 12) & 0xf) - 1, di = ((li >> 16) & 0xff) - 1;
            V3 p_a, p_l;
            if (src == 0) p_a = p_l = v3(0.f, 0.f, 0.f);
            else if (src == 1) { p_a = a_a; p_l = a_l; }
            else if (src == 2) { p_a = sv_a[0]; p_l = sv_l[0]; }
            else if (src == 3) { p_a = sv_a[1]; p_l = sv_l[1]; }
            else { p_a = sv_a[2]; p_l = sv_l[2]; }
            V3 rr = v3(L[12], L[13], L[14]);
            a_a = p_a;
            a_l = p_l + cross(p_a, rr);
            if (di >= 0) {
                V3 sa = v3(L[0], L[1], L[2]), sl = v3(L[3], L[4], L[5]);
                V3 ha = v3(L[6], L[7], L[8]), hl = v3(L[9], L[10], L[11]);
                float Y = ((ancm >> i) & 1u) ? sm_Y[i * kMaxDof + lane] : 0.f;
                float x = (Y - (dot(ha, a_a) + dot(hl, a_l))) * L[15];
                a_a = fma3(x, sa, a_a);
                a_l = fma3(x, sl, a_l);
                if (colv) sm_Minv[di * kMinvStride + lane] = x;
            }
            if (save == 0) { sv_a[0] = a_a; sv_l[0] = a_l; }
            else if (save == 1) { sv_a[1] = a_a; sv_l[1] = a_l; }
            else if (save == 2) { sv_a[2] = a_a; sv_l[2] = a_l; }
        }
    }
    __syncwarp();

    // ---- constraint rows: joint limits first (created at import), then one motor row per dof ----
    float* r_dof = sm_row;            // dof index of the row's joint (as float bits)
    float* r_sign = sm_row + 32;
    float* r_rhs = sm_row + 64;       // already multiplied by jacDiagABInv
    float* r_lo = sm_row + 96;
    float* r_hi = sm_row + 128;
    float* r_jdi = sm_row + 160;
    int nrows;
    {
        const float lower = valid ? __ldg(&M->lower[link]) : 0.f, upper = valid ? __ldg(&M->upper[link]) : 0.f;
        const bool lim = hasdof && valid && __ldg(&M->has_limit[link]) != 0;
        const float pen0 = q - lower, pen1 = upper - q;
        const bool l0 = lim && !(pen0 > 0.f), l1 = lim && !(pen1 > 0.f);
        const unsigned b0 = __ballot_sync(kFull, l0), b1 = __ballot_sync(kFull, l1), bm = __ballot_sync(kFull, hasdof);
        const unsigned lt = (1u << lane) - 1u;
        const int nlim = __popc(b0) + __popc(b1);
        const int i0 = __popc(b0 & lt) + __popc(b1 & lt);
        const int i1 = i0 + (l0 ? 1 : 0);
        const int im = nlim + __popc(bm & lt);
        nrows = nlim + __popc(bm);
        if (hasdof) {
            const float jdi = 1.f / sm_Minv[mydof * kMinvStride + mydof];
            const float inv_dt = M->inv_dt;
            if (l0) {
                r_dof[i0] = __int_as_float(mydof); r_sign[i0] = 1.f; r_jdi[i0] = jdi;
                r_rhs[i0] = (-pen0 * M->erp * inv_dt - qs) * jdi;
                r_lo[i0] = 0.f; r_hi[i0] = M->limit_max_imp;
            }
            if (l1) {
                r_dof[i1] = __int_as_float(mydof); r_sign[i1] = -1.f; r_jdi[i1] = jdi;
                r_rhs[i1] = (-pen1 * M->erp * inv_dt + qs) * jdi;
                r_lo[i1] = 0.f; r_hi[i1] = M->limit_max_imp;
            }
            // btMultiBodyJointMotor: rhs = kp (q_des - q)/dt + qs + kd (qd_des - qs), kd = 1, erp = 1
            // (the row's right-hand side is rhs - qs; formed directly to avoid the cancellation)
            const float rhs_rel = fmaf(mot.kp * (mot.tpos - q), inv_dt, mot.tvel - qs);
            r_dof[im] = __int_as_float(mydof); r_sign[im] = 1.f; r_jdi[im] = jdi;
            r_rhs[im] = rhs_rel * jdi;
            r_lo[im] = -mot.maximp; r_hi[im] = mot.maximp;
        }
    }
    __syncwarp();

    // ---- projected Gauss-Seidel in row space: lane k owns row k ----
    const bool rowv = lane < nrows;
    const int rd = rowv ? __float_as_int(r_dof[lane]) : 0;
    const float rsg = rowv ? r_sign[lane] : 0.f, rrhs = rowv ? r_rhs[lane] : 0.f;
    const float rlo = rowv ? r_lo[lane] : 0.f, rhi = rowv ? r_hi[lane] : 0.f;
    const float rjdi = rowv ? r_jdi[lane] : 0.f;
    const float rinvj = rowv ? sm_Minv[rd * kMinvStride + rd] : 0.f;
    float wk = 0.f, app = 0.f;
    int it = 0;
    const int iters = M->iters;
    const float thresh = M->resid_thresh;
    for (it = 0; it < iters; it++) {
        float resid = 0.f;
        for (int jj = 0; jj < nrows; jj++) {
            const int rr = (it & 1) ? jj : nrows - 1 - jj;
            float delta = fmaf(-wk, rjdi, rrhs);
            const float sum = app + delta;
            float napp = sum;
            if (sum < rlo) { delta = rlo - app; napp = rlo; }
            else if (sum > rhi) { delta = rhi - app; napp = rhi; }
            const float dr = shf(delta, rr);
            const int dofr = __float_as_int(r_dof[rr]);
            const float sgr = r_sign[rr];
            if (lane == rr) {
                app = napp;
                const float dvel = delta * rinvj;
                resid = fmaxf(resid, dvel * dvel);
            }
            wk = fmaf(dr * (rsg * sgr), sm_Minv[dofr * kMinvStride + rd], wk);
        }
        const bool more = __any_sync(kFull, resid > thresh);
        if (!more || it >= iters - 1) { it++; break; }
    }

    // ---- velocity and position update (lane = link again) ----
    float dv = 0.f;
    {
        const float as = app * rsg;
        for (int k = 0; k < nrows; k++) {
            const float ak = shf(as, k);
            const int dk = shi(rd, k);
            if (hasdof) dv = fmaf(ak, sm_Minv[dk * kMinvStride + mydof], dv);
        }
    }
    if (hasdof) {
        const float x = fminf(fmaxf(qs + dv, -max_vel), max_vel);
        qd = x;
        q = fmaf(M->dt, x, q);
    } else {
        qd = 0.f;
    }
    __syncwarp();
    return it;
}

}  // namespace rloa
