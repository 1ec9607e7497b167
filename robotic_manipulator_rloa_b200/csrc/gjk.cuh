// GJK closest distance between a posed convex shape (hull of a vertex cloud, or a box) and an axis-aligned box or a
// point, one query per thread, fp32.  This is what p.getClosestPoints does for mesh collision shapes and box pairs
// (reference utils/collision_detector.py:47-52 -> btGjkPairDetector with btVoronoiSimplexSolver); no EPA: every
// consumer on the path compares the distance with a threshold, so overlapping cores report 0 and the margin /
// sphere radius subtracted by the caller make the result negative.
// Included by sim_device.cuh inside namespace rloa, after the V3 / M3 helpers.
#pragma once

struct GjkShape {
    const float4* verts;   // hull: shape-frame vertices (uniform address across the warp: broadcast loads); null = box
    int nv;
    V3 half;               // box half extents
    M3 R;                  // world <- shape
    V3 p;
};

__device__ __forceinline__ V3 gjk_support(const GjkShape& A, V3 dir_w) {
    const V3 dl = mulT(A.R, dir_w);
    V3 loc;
    if (A.verts == nullptr) {
        loc = v3(dl.x >= 0.f ? A.half.x : -A.half.x, dl.y >= 0.f ? A.half.y : -A.half.y, dl.z >= 0.f ? A.half.z : -A.half.z);
    } else {
        float best = -3.0e38f;
        loc = v3(0.f, 0.f, 0.f);
        for (int i = 0; i < A.nv; i++) {
            const float4 v = __ldg(A.verts + i);
            const float d = fmaf(dl.x, v.x, fmaf(dl.y, v.y, dl.z * v.z));
            if (d > best) { best = d; loc = v3(v.x, v.y, v.z); }
        }
    }
    return A.p + mul(A.R, loc);
}

// closest point of triangle (a, b, c) to the origin and the mask of vertices that support it (Voronoi regions)
__device__ __forceinline__ V3 gjk_tri_closest(V3 a, V3 b, V3 c, int& mask) {
    const V3 ab = b - a, ac = c - a;
    const float d1 = -dot(ab, a), d2 = -dot(ac, a);
    if (d1 <= 0.f && d2 <= 0.f) { mask = 1; return a; }
    const float d3 = -dot(ab, b), d4 = -dot(ac, b);
    if (d3 >= 0.f && d4 <= d3) { mask = 2; return b; }
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { mask = 3; return fma3(d1 / (d1 - d3), ab, a); }
    const float d5 = -dot(ab, c), d6 = -dot(ac, c);
    if (d6 >= 0.f && d5 <= d6) { mask = 4; return c; }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { mask = 5; return fma3(d2 / (d2 - d6), ac, a); }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        mask = 6;
        return fma3((d4 - d3) / ((d4 - d3) + (d5 - d6)), c - b, b);
    }
    const float den = 1.f / (va + vb + vc);
    mask = 7;
    return fma3(vb * den, ab, fma3(vc * den, ac, a));
}

// closest point of the simplex P[0..n) to the origin; P shrinks to the supporting sub-simplex.  true = the origin is
// inside a tetrahedron (overlap).
__device__ __forceinline__ bool gjk_simplex_closest(V3* P, int& n, V3& v) {
    if (n == 1) { v = P[0]; return false; }
    if (n == 2) {
        const V3 ab = P[1] - P[0];
        const float t = -dot(P[0], ab), L2 = dot(ab, ab);
        if (t <= 0.f || L2 <= 0.f) { v = P[0]; n = 1; return false; }
        if (t >= L2) { v = P[1]; P[0] = P[1]; n = 1; return false; }
        v = fma3(t / L2, ab, P[0]);
        return false;
    }
    int mask = 0;
    if (n == 3) {
        v = gjk_tri_closest(P[0], P[1], P[2], mask);
    } else {
        // the closest point lies on a face whose plane separates the origin from the fourth vertex
        float best = 3.0e38f;
        bool any = false;
#pragma unroll
        for (int f = 0; f < 4; f++) {
            const int i0 = f == 3 ? 1 : 0, i1 = f == 0 ? 1 : (f == 1 ? 2 : 3), i2 = f == 0 ? 2 : (f == 1 ? 3 : (f == 2 ? 1 : 2)),
                      i3 = f == 0 ? 3 : (f == 1 ? 1 : (f == 2 ? 2 : 0));
            const V3 a = P[i0], b = P[i1], c = P[i2], d = P[i3];
            const V3 nrm = cross(b - a, c - a), ad = d - a;
            const float sp = -dot(a, nrm), sd = dot(ad, nrm);
            if (sp * sd < 0.f || sd * sd <= 1e-12f * dot(nrm, nrm) * dot(ad, ad)) {   // outside, or a flat tetrahedron
                int m3;
                const V3 c3 = gjk_tri_closest(a, b, c, m3);
                const float dd = dot(c3, c3);
                if (dd < best) {
                    best = dd;
                    any = true;
                    v = c3;
                    mask = ((m3 & 1) ? 1 << i0 : 0) | ((m3 & 2) ? 1 << i1 : 0) | ((m3 & 4) ? 1 << i2 : 0);
                }
            }
        }
        if (!any) { v = v3(0.f, 0.f, 0.f); return true; }
    }
    int k = 0;
    for (int i = 0; i < n; i++)
        if ((mask >> i) & 1) { P[k] = P[i]; k++; }
    n = k;
    return false;
}

// support mapping of the second shape: an axis-aligned box / point (obstacle sphere centre, target cube) ...
struct GjkAabb {
    V3 c, h;
    __device__ __forceinline__ V3 operator()(V3 d) const {
        return v3(c.x + (d.x >= 0.f ? h.x : -h.x), c.y + (d.y >= 0.f ? h.y : -h.y), c.z + (d.z >= 0.f ? h.z : -h.z));
    }
    __device__ __forceinline__ V3 centre() const { return c; }
};
// ... or another posed shape (link vs link)
struct GjkPosed {
    const GjkShape& S;
    __device__ __forceinline__ V3 operator()(V3 d) const { return gjk_support(S, d); }
    __device__ __forceinline__ V3 centre() const { return S.p; }
};

// distance between the cores of shape A and shape B (given by its support mapping); 0 on overlap
template <class SupportB>
__device__ __forceinline__ float gjk_run(const GjkShape& A, const SupportB& B) {
    V3 P[4];
    int n = 0;
    V3 v = A.p - B.centre();
    if (dot(v, v) < 1e-20f) v = v3(1.f, 0.f, 0.f);
    for (int it = 0; it < 32; it++) {
        const V3 w = gjk_support(A, v3(-v.x, -v.y, -v.z)) - B(v);      // support of A - B along -v
        if (it > 0) {
            const float vv = dot(v, v);
            if (vv - dot(v, w) <= 2e-6f * vv) break;      // no vertex of A - B is closer: v is the closest point
            bool dup = false;
            for (int i = 0; i < n; i++) dup = dup || (P[i].x == w.x && P[i].y == w.y && P[i].z == w.z);
            if (dup) break;
        }
        P[n++] = w;
        if (gjk_simplex_closest(P, n, v) || dot(v, v) < 1e-14f) return 0.f;
    }
    return sqrtf(dot(v, v));
}

// shape A against the axis-aligned box (centre bc, half extents bh; bh = 0: a point)
__device__ __noinline__ float gjk_distance(const GjkShape& A, V3 bc, V3 bh) { return gjk_run(A, GjkAabb{bc, bh}); }
// shape A against shape B
__device__ __noinline__ float gjk_distance_pair(const GjkShape& A, const GjkShape& B) { return gjk_run(A, GjkPosed{B}); }
