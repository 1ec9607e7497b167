/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see bullet_restatement.h for the scope statement).
 * PARITY UNPINNED against PyBullet itself: no PyBullet, no assets, no golden physics vectors here.
 *
 * Double-precision, single-env-at-a-time restatement of the Bullet3 multibody step that runs
 * under the reference's Environment.step / Environment.reset
 * (/root/reference/robotic_manipulator_rloa/environment/environment.py:264-309, 453-485).
 * Deliberately plain: dense 6x6 spatial algebra, one function per Bullet routine.
 */
#include "bullet_restatement.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------- small linear algebra ---------- */
static void m3v(const double* A, const double* x, double* y) {
    for (int i = 0; i < 3; i++) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}
static void m3tv(const double* A, const double* x, double* y) {
    for (int i = 0; i < 3; i++) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2];
}
static void m3m(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j];
            C[3 * i + j] = s;
        }
}
static void m3mt(const double* A, const double* B, double* C) { /* C = A * B^T */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * j + k];
            C[3 * i + j] = s;
        }
}
static void cross3(const double* a, const double* b, double* c) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    c[0] = x; c[1] = y; c[2] = z;
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double dot6(const double* a, const double* b) {
    double s = 0;
    for (int i = 0; i < 6; i++) s += a[i] * b[i];
    return s;
}
static void rodrigues(const double* a, double th, double* R) {
    double c = cos(th), s = sin(th), t = 1 - c;
    R[0] = c + t * a[0] * a[0];        R[1] = t * a[0] * a[1] - s * a[2]; R[2] = t * a[0] * a[2] + s * a[1];
    R[3] = t * a[1] * a[0] + s * a[2]; R[4] = c + t * a[1] * a[1];        R[5] = t * a[1] * a[2] - s * a[0];
    R[6] = t * a[2] * a[0] - s * a[1]; R[7] = t * a[2] * a[1] + s * a[0]; R[8] = c + t * a[2] * a[2];
}

/* ---------- per-link kinematics (btMultiBody::updateCacheMultiDof equivalents) ---------- */
typedef struct {
    double E[9];    /* child <- parent */
    double r[3];    /* parent origin -> child origin, child frame (Bullet m_cachedRVector) */
    double s[6];    /* motion subspace [ang; lin] in child COM frame (Bullet m_axes[0]) */
} linkkin;

static void link_kin(const orc_model* m, int i, double qi, linkkin* k) {
    double t[3];
    if (m->jtype[i] == ORC_REVOLUTE) {
        double Rq[9];
        rodrigues(m->axis[i], -qi, Rq);        /* btQuaternion(axis, -q) * zeroRotParentToThis */
        m3m(Rq, m->E0[i], k->E);
        m3v(k->E, m->e[i], t);
        for (int a = 0; a < 3; a++) k->r[a] = t[a] + m->d[i][a];
        for (int a = 0; a < 3; a++) k->s[a] = m->axis[i][a];
        cross3(m->axis[i], m->d[i], k->s + 3); /* m_bottomVec = axis x dVector */
    } else {
        memcpy(k->E, m->E0[i], sizeof(k->E));
        m3v(k->E, m->e[i], t);
        double qq = (m->jtype[i] == ORC_PRISMATIC) ? qi : 0.0;
        for (int a = 0; a < 3; a++) k->r[a] = t[a] + m->d[i][a] + qq * m->axis[i][a];
        for (int a = 0; a < 6; a++) k->s[a] = 0;
        if (m->jtype[i] == ORC_PRISMATIC)
            for (int a = 0; a < 3; a++) k->s[3 + a] = m->axis[i][a];
    }
}
/* motion transform child<-parent: btSpatialTransformationMatrix::transform */
static void xmot(const linkkin* k, const double* vp, double* vc) {
    double w[3], v[3], c[3];
    m3v(k->E, vp, w);
    m3v(k->E, vp + 3, v);
    cross3(k->r, w, c);
    for (int a = 0; a < 3; a++) { vc[a] = w[a]; vc[3 + a] = v[a] - c[a]; }
}
/* force transform parent<-child: btSpatialTransformationMatrix::transformInverse */
static void xforce_inv(const linkkin* k, const double* fc, double* fp) {
    double c[3], t[3];
    cross3(k->r, fc + 3, c);
    for (int a = 0; a < 3; a++) t[a] = fc[a] + c[a];
    m3tv(k->E, t, fp);
    m3tv(k->E, fc + 3, fp + 3);
}
/* dense 6x6 motion transform matrix child<-parent */
static void xmat(const linkkin* k, double X[36]) {
    double rx[9] = {0, -k->r[2], k->r[1], k->r[2], 0, -k->r[0], -k->r[1], k->r[0], 0};
    double rxE[9];
    m3m(rx, k->E, rxE);
    memset(X, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            X[6 * i + j] = k->E[3 * i + j];
            X[6 * (i + 3) + (j + 3)] = k->E[3 * i + j];
            X[6 * (i + 3) + j] = -rxE[3 * i + j];
        }
}
static void m6v(const double* A, const double* x, double* y) {
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int j = 0; j < 6; j++) s += A[6 * i + j] * x[j];
        y[i] = s;
    }
}
/* Ip += X^T Ic X */
static void congruence_add(const double* X, const double* Ic, double* Ip) {
    double T[36];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += Ic[6 * i + k] * X[6 * k + j];
            T[6 * i + j] = s;
        }
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += X[6 * k + i] * T[6 * k + j];
            Ip[6 * i + j] += s;
        }
}

void orc_fk(const orc_model* m, const double* q, double* Rw, double* pw) {
    for (int i = 0; i < m->nl; i++) {
        linkkin k;
        link_kin(m, i, q[i], &k);
        const double* Rp = (m->parent[i] < 0) ? m->base_R : Rw + 9 * m->parent[i];
        const double* pp = (m->parent[i] < 0) ? m->base_p : pw + 3 * m->parent[i];
        double t[3];
        m3mt(Rp, k.E, Rw + 9 * i);          /* world<-child = world<-parent * E^T */
        m3v(Rw + 9 * i, k.r, t);
        for (int a = 0; a < 3; a++) pw[3 * i + a] = pp[a] + t[a];
    }
}

/* ---------- articulated-body workspace ---------- */
typedef struct {
    linkkin k[ORC_MAXL];
    double h[ORC_MAXL][6];
    double D[ORC_MAXL];
    int hasdof[ORC_MAXL];
} abawork;

/* btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof, fixed base, one dof per joint.
 * tau_ext may be NULL.  Joint damping torque (-damping*qd) is added here, as
 * PhysicsServerCommandProcessor does before stepping.  Fills the workspace reused by unit responses. */
static void aba_core(const orc_model* m, const double* q, const double* qd, const double* tau_ext,
                     double* qdd, abawork* w) {
    int nl = m->nl;
    double v[ORC_MAXL][6], c[ORC_MAXL][6], pA[ORC_MAXL][6], Y[ORC_MAXL];
    static const double zero6[6] = {0, 0, 0, 0, 0, 0};
    double (*IA)[36] = (double (*)[36])malloc(sizeof(double[36]) * nl);
    double Rw[ORC_MAXL][9];

    for (int i = 0; i < nl; i++) {
        linkkin* k = &w->k[i];
        link_kin(m, i, q[i], k);
        w->hasdof[i] = (m->jtype[i] != ORC_FIXED);
        int p = m->parent[i];
        const double* vp = p < 0 ? zero6 : v[p];
        const double* Rp = p < 0 ? m->base_R : Rw[p];
        m3mt(Rp, k->E, Rw[i]);
        xmot(k, vp, v[i]);
        double vj[6];
        for (int a = 0; a < 6; a++) { vj[a] = k->s[a] * qd[i]; v[i][a] += vj[a]; }
        /* spatVel.cross(spatJointVel, spatCoriolisAcc) */
        double t1[3], t2[3];
        cross3(v[i], vj, c[i]);
        cross3(v[i], vj + 3, t1);
        cross3(v[i] + 3, vj, t2);
        for (int a = 0; a < 3; a++) c[i][3 + a] = t1[a] + t2[a];
        /* isolated inertia and zero-acceleration force */
        memset(IA[i], 0, sizeof(double[36]));
        for (int a = 0; a < 3; a++) { IA[i][6 * a + a] = m->inertia[i][a]; IA[i][6 * (a + 3) + a + 3] = m->mass[i]; }
        double g_l[3], Iw[3], gy[3], cv[3];
        m3tv(Rw[i], m->gravity, g_l);       /* gravity in link frame */
        for (int a = 0; a < 3; a++) Iw[a] = m->inertia[i][a] * v[i][a];
        cross3(v[i], Iw, gy);               /* omega x I omega (m_useGyroTerm) */
        cross3(v[i], v[i] + 3, cv);         /* omega x v */
        double wn = sqrt(dot3(v[i], v[i])), vn = sqrt(dot3(v[i] + 3, v[i] + 3));
        for (int a = 0; a < 3; a++) {
            pA[i][a] = gy[a] + Iw[a] * (m->ang_damp + m->ang_damp * wn);
            pA[i][3 + a] = m->mass[i] * cv[a] - m->mass[i] * g_l[a]
                           + m->mass[i] * v[i][3 + a] * (m->lin_damp + m->lin_damp * vn);
        }
    }
    for (int i = nl - 1; i >= 0; i--) {
        linkkin* k = &w->k[i];
        int p = m->parent[i];
        double Ia[36], pa[6], Ic[6];
        memcpy(Ia, IA[i], sizeof(Ia));
        m6v(IA[i], c[i], Ic);
        for (int a = 0; a < 6; a++) pa[a] = pA[i][a] + Ic[a];
        if (w->hasdof[i]) {
            m6v(IA[i], k->s, w->h[i]);
            w->D[i] = dot6(k->s, w->h[i]);
            double tau = -m->damping[i] * qd[i] + (tau_ext ? tau_ext[i] : 0.0);
            Y[i] = tau - dot6(k->s, pA[i]) - dot6(c[i], w->h[i]);
            double invD = 1.0 / w->D[i];
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 6; b++) Ia[6 * a + b] -= w->h[i][a] * w->h[i][b] * invD;
            for (int a = 0; a < 6; a++) pa[a] += w->h[i][a] * (Y[i] * invD);
        } else {
            for (int a = 0; a < 6; a++) w->h[i][a] = 0;
            w->D[i] = 1.0;
            Y[i] = 0;
        }
        if (p >= 0) {
            double X[36], fp[6];
            xmat(k, X);
            congruence_add(X, Ia, IA[p]);
            xforce_inv(k, pa, fp);
            for (int a = 0; a < 6; a++) pA[p][a] += fp[a];
        }
    }
    double acc[ORC_MAXL][6];
    for (int i = 0; i < nl; i++) {
        int p = m->parent[i];
        const double* ap = p < 0 ? zero6 : acc[p];
        xmot(&w->k[i], ap, acc[i]);
        double a_j = 0;
        if (w->hasdof[i]) a_j = (Y[i] - dot6(acc[i], w->h[i])) / w->D[i];
        qdd[i] = a_j;
        for (int a = 0; a < 6; a++) acc[i][a] += c[i][a] + w->k[i].s[a] * a_j;
    }
    free(IA);
}

/* btMultiBody::calcAccelerationDeltasMultiDof for a unit joint-space impulse at link j */
static void unit_response(const orc_model* m, const abawork* w, int j, double* out) {
    int nl = m->nl;
    double zf[ORC_MAXL][6], Y[ORC_MAXL], acc[ORC_MAXL][6];
    static const double zero6[6] = {0, 0, 0, 0, 0, 0};
    memset(zf, 0, sizeof(zf));
    for (int i = nl - 1; i >= 0; i--) {
        int p = m->parent[i];
        double f[6];
        memcpy(f, zf[i], sizeof(f));
        if (w->hasdof[i]) {
            Y[i] = (i == j ? 1.0 : 0.0) - dot6(w->k[i].s, zf[i]);
            double t = Y[i] / w->D[i];
            for (int a = 0; a < 6; a++) f[a] += w->h[i][a] * t;
        } else Y[i] = 0;
        if (p >= 0) {
            double fp[6];
            xforce_inv(&w->k[i], f, fp);
            for (int a = 0; a < 6; a++) zf[p][a] += fp[a];
        }
    }
    for (int i = 0; i < nl; i++) {
        int p = m->parent[i];
        const double* ap = p < 0 ? zero6 : acc[p];
        xmot(&w->k[i], ap, acc[i]);
        double a_j = 0;
        if (w->hasdof[i]) a_j = (Y[i] - dot6(acc[i], w->h[i])) / w->D[i];
        out[i] = a_j;
        for (int a = 0; a < 6; a++) acc[i][a] += w->k[i].s[a] * a_j;
    }
}

void orc_aba(const orc_model* m, const double* q, const double* qd, const double* tau_ext, double* qdd) {
    abawork w;
    aba_core(m, q, qd, tau_ext, qdd, &w);
}

void orc_minv(const orc_model* m, const double* q, double* Minv) {
    abawork w;
    double qd[ORC_MAXL] = {0}, qdd[ORC_MAXL];
    aba_core(m, q, qd, NULL, qdd, &w);
    for (int j = 0; j < m->nl; j++) {
        double col[ORC_MAXL];
        if (w.hasdof[j]) unit_response(m, &w, j, col);
        else memset(col, 0, sizeof(col));
        for (int i = 0; i < m->nl; i++) Minv[i * m->nl + j] = col[i];
    }
}

/* ---------- independent formulation for cross-checks: recursive Newton-Euler with the gravity-as-
 * base-acceleration trick (Featherstone RBDA table 5.1).  Includes Bullet's link velocity drag and
 * joint damping when with_bias != 0. ---------- */
static void rnea(const orc_model* m, const double* q, const double* qd, const double* qdd, int with_bias,
                 double* tau) {
    int nl = m->nl;
    linkkin k[ORC_MAXL];
    double v[ORC_MAXL][6], a[ORC_MAXL][6], f[ORC_MAXL][6];
    double a0[6] = {0, 0, 0, 0, 0, 0}, zero6[6] = {0, 0, 0, 0, 0, 0};
    if (with_bias) { /* base accelerates with -g, expressed in base frame */
        double g_b[3];
        m3tv(m->base_R, m->gravity, g_b);
        for (int x = 0; x < 3; x++) a0[3 + x] = -g_b[x];
    }
    for (int i = 0; i < nl; i++) {
        link_kin(m, i, q[i], &k[i]);
        int p = m->parent[i];
        xmot(&k[i], p < 0 ? zero6 : v[p], v[i]);
        xmot(&k[i], p < 0 ? a0 : a[p], a[i]);
        double vj[6], cc[6], t1[3], t2[3];
        for (int x = 0; x < 6; x++) { vj[x] = k[i].s[x] * (with_bias ? qd[i] : 0.0); v[i][x] += vj[x]; }
        cross3(v[i], vj, cc);
        cross3(v[i], vj + 3, t1);
        cross3(v[i] + 3, vj, t2);
        for (int x = 0; x < 3; x++) cc[3 + x] = t1[x] + t2[x];
        for (int x = 0; x < 6; x++) a[i][x] += cc[x] + k[i].s[x] * qdd[i];
        /* f = I a + v x* I v (+ drag) */
        double Iw[3], gy[3], cv[3];
        for (int x = 0; x < 3; x++) Iw[x] = m->inertia[i][x] * v[i][x];
        cross3(v[i], Iw, gy);
        cross3(v[i], v[i] + 3, cv);
        double wn = sqrt(dot3(v[i], v[i])), vn = sqrt(dot3(v[i] + 3, v[i] + 3));
        for (int x = 0; x < 3; x++) {
            f[i][x] = m->inertia[i][x] * a[i][x] + gy[x];
            f[i][3 + x] = m->mass[i] * a[i][3 + x] + m->mass[i] * cv[x];
            if (with_bias) {
                f[i][x] += Iw[x] * (m->ang_damp + m->ang_damp * wn);
                f[i][3 + x] += m->mass[i] * v[i][3 + x] * (m->lin_damp + m->lin_damp * vn);
            }
        }
    }
    for (int i = nl - 1; i >= 0; i--) {
        tau[i] = (m->jtype[i] != ORC_FIXED) ? dot6(k[i].s, f[i]) : 0.0;
        if (with_bias) tau[i] += m->damping[i] * qd[i];
        int p = m->parent[i];
        if (p >= 0) {
            double fp[6];
            xforce_inv(&k[i], f[i], fp);
            for (int x = 0; x < 6; x++) f[p][x] += fp[x];
        }
    }
}
void orc_rnea_bias(const orc_model* m, const double* q, const double* qd, double* bias) {
    double z[ORC_MAXL] = {0};
    rnea(m, q, qd, z, 1, bias);
}
void orc_crba(const orc_model* m, const double* q, double* M) {
    int nl = m->nl;
    double z[ORC_MAXL] = {0};
    for (int j = 0; j < nl; j++) {
        double e[ORC_MAXL] = {0}, col[ORC_MAXL];
        e[j] = 1.0;
        rnea(m, q, z, e, 0, col);
        for (int i = 0; i < nl; i++) M[i * nl + j] = (m->jtype[j] != ORC_FIXED) ? col[i] : 0.0;
    }
}

/* ---------- contact rows against the two collidable fixed bodies of the reference (environment.py:252-255: the obstacle
 * sphere and the target cube are loaded with useFixedBase, so stepSimulation generates contacts between them and the
 * manipulator's links).  Restated from btMultiBodyConstraintSolver::setupMultiBodyContactConstraint /
 * resolveSingleConstraintRowGeneric: one NORMAL row per (shape, body) pair closer than the contact breaking threshold
 * (gContactBreakingThreshold = 0.02), lambda in [0, inf), no split impulse:
 *     separated  (d > 0): rhs = (-J qs - d / dt) / (J M^-1 J^T)        — acts only if the gap would close within the step
 *     penetrating (d <= 0): rhs = (-J qs - d erp / dt) / (J M^-1 J^T)  — erp = 0.2 pushes the penetration out
 * swept after the non-contact rows of every iteration, in contact order.  NOT restated: friction rows (lateral friction of
 * the two URDFs is not known here), the persistent manifold's multi-point caching (one point per pair), contacts of box /
 * hull shapes against the cube. ---------- */
typedef struct {
    int n;
    int link[ORC_MAXC];
    double nrm[ORC_MAXC][3];   /* on the fixed body, pointing at the link */
    double pA[ORC_MAXC][3];    /* contact point on the link, world */
    double dist[ORC_MAXC];
} orc_contact_set;

static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* argmin over t in [0,1] of the distance from a + t (b - a) to the origin-centred box h (same pieces as segment_box) */
static double segment_box_argmin(const double* a, const double* b, const double* h, double* t_best);

static void add_contact(orc_contact_set* cs, int link, const double* x, double r_shape, const double* from, double r_body,
                        double thr) {
    /* core point x of the link's shape against the closest point `from` of the fixed body (sphere centre / cube surface) */
    double v[3] = {x[0] - from[0], x[1] - from[1], x[2] - from[2]};
    double L = sqrt(dot3(v, v));
    if (L < 1e-9 || cs->n >= ORC_MAXC) return;
    double d = L - r_shape - r_body;
    if (!(d < thr)) return;
    int c = cs->n++;
    cs->link[c] = link;
    cs->dist[c] = d;
    for (int a = 0; a < 3; a++) {
        cs->nrm[c][a] = v[a] / L;
        cs->pA[c][a] = x[a] - r_shape * v[a] / L;
    }
}

static void find_contacts(const orc_model* m, const double* Rw, const double* pw, const double* obstacle,
                          const double* target, double thr, orc_contact_set* cs) {
    cs->n = 0;
    for (int s = 0; s < m->ns; s++) {
        int l = m->s_link[s], type = m->s_type[s];
        if (type == ORC_SHAPE_HULL) continue;
        double Rs[9], ps[3], t[3];
        m3m(Rw + 9 * l, m->s_R[s], Rs);
        m3v(Rw + 9 * l, m->s_p[s], t);
        for (int a = 0; a < 3; a++) ps[a] = pw[3 * l + a] + t[a];
        double ax[3] = {Rs[2], Rs[5], Rs[8]};
        double r_s = type == ORC_SHAPE_BOX ? 0.0 : m->s_dim[s][0], hz = m->s_dim[s][1];
        /* --- obstacle sphere --- */
        {
            double x[3];
            if (type == ORC_SHAPE_SPHERE) { for (int a = 0; a < 3; a++) x[a] = ps[a]; }
            else if (type == ORC_SHAPE_CAPSULE) {
                double rel[3] = {obstacle[0] - ps[0], obstacle[1] - ps[1], obstacle[2] - ps[2]};
                double tz = clampd(dot3(rel, ax), -hz, hz);
                for (int a = 0; a < 3; a++) x[a] = ps[a] + tz * ax[a];
            } else { /* box: closest point of the box to the sphere centre */
                double rel[3] = {obstacle[0] - ps[0], obstacle[1] - ps[1], obstacle[2] - ps[2]}, loc[3], w[3];
                m3tv(Rs, rel, loc);
                for (int a = 0; a < 3; a++) loc[a] = clampd(loc[a], -m->s_dim[s][a], m->s_dim[s][a]);
                m3v(Rs, loc, w);
                for (int a = 0; a < 3; a++) x[a] = ps[a] + w[a];
            }
            add_contact(cs, l, x, r_s, obstacle, m->obstacle_radius, thr);
        }
        /* --- target cube (axis-aligned) --- */
        if (type != ORC_SHAPE_BOX) {
            double x[3], y[3];
            if (type == ORC_SHAPE_SPHERE) { for (int a = 0; a < 3; a++) x[a] = ps[a]; }
            else {
                double a1[3], b1[3], tb;
                for (int a = 0; a < 3; a++) { a1[a] = ps[a] - target[a] - hz * ax[a]; b1[a] = ps[a] - target[a] + hz * ax[a]; }
                segment_box_argmin(a1, b1, m->target_half, &tb);
                for (int a = 0; a < 3; a++) x[a] = target[a] + a1[a] + tb * (b1[a] - a1[a]);
            }
            for (int a = 0; a < 3; a++) y[a] = target[a] + clampd(x[a] - target[a], -m->target_half[a], m->target_half[a]);
            add_contact(cs, l, x, r_s, y, 0.0, thr);
        }
    }
}

/* row of the point Jacobian: d(n . pA)/dq for every dof (link index), 0 off the contact link's chain */
static void contact_jacobian(const orc_model* m, const double* Rw, const double* pw, int link, const double* n, const double* pA,
                             double* J) {
    for (int i = 0; i < m->nl; i++) J[i] = 0.0;
    for (int k = link; k >= 0; k = m->parent[k]) {
        if (m->jtype[k] == ORC_FIXED) continue;
        double aw[3], dw[3];
        m3v(Rw + 9 * k, m->axis[k], aw);
        if (m->jtype[k] == ORC_PRISMATIC) { J[k] = dot3(n, aw); continue; }
        m3v(Rw + 9 * k, m->d[k], dw);                       /* pivot -> COM of link k, world */
        double rp[3] = {pA[0] - pw[3 * k] + dw[0], pA[1] - pw[3 * k + 1] + dw[1], pA[2] - pw[3 * k + 2] + dw[2]}, cr[3];
        cross3(aw, rp, cr);
        J[k] = dot3(n, cr);
    }
}

/* ---------- one stepSimulation: ABA -> rows -> PGS -> integrate ---------- */
static int substep_impl(const orc_model* m, const orc_motors* mot, double* q, double* qd, const double* obstacle,
                        const double* target, double contact_thr, int* n_contacts) {
    int nl = m->nl;
    abawork w;
    double qdd[ORC_MAXL], qs[ORC_MAXL];
    aba_core(m, q, qd, NULL, qdd, &w);
    for (int i = 0; i < nl; i++) {   /* applyDeltaVeeMultiDof(output, dt) with clamp */
        double x = qd[i] + m->dt * qdd[i];
        if (x > m->max_vel) x = m->max_vel;
        if (x < -m->max_vel) x = -m->max_vel;
        qs[i] = w.hasdof[i] ? x : 0.0;
    }
    /* unit responses (columns of M^-1) for every movable joint */
    double col[ORC_MAXL][ORC_MAXL];
    for (int j = 0; j < nl; j++)
        if (w.hasdof[j]) unit_response(m, &w, j, col[j]);
    /* rows: joint limits first (created at import), then motors (created after load) */
    int nrows = 0, rlink[2 * ORC_MAXL];
    double rsign[2 * ORC_MAXL], rrhs[2 * ORC_MAXL], rlo[2 * ORC_MAXL], rhi[2 * ORC_MAXL], rjdi[2 * ORC_MAXL],
        rapp[2 * ORC_MAXL];
    for (int i = 0; i < nl; i++) {
        if (!w.hasdof[i] || !m->has_limit[i]) continue;
        for (int row = 0; row < 2; row++) {
            double pen = row == 0 ? q[i] - m->lower[i] : m->upper[i] - q[i];
            if (pen > 0) continue;                       /* btMultiBodyJointLimitConstraint: skip */
            double sign = row ? -1.0 : 1.0;
            double jdi = 1.0 / col[i][i];
            double rel_vel = sign * qs[i];
            double pos_err = -pen * m->erp / m->dt;
            rlink[nrows] = i; rsign[nrows] = sign; rjdi[nrows] = jdi;
            rrhs[nrows] = (pos_err - rel_vel) * jdi;
            rlo[nrows] = 0.0; rhi[nrows] = m->limit_max_impulse; rapp[nrows] = 0.0;
            nrows++;
        }
    }
    for (int i = 0; i < nl; i++) {
        if (!w.hasdof[i]) continue;
        double jdi = 1.0 / col[i][i];
        /* btMultiBodyJointMotor::createConstraintRows, erp = 1 */
        double rhs = mot->kp[i] * (mot->tpos[i] - q[i]) / m->dt + qs[i] + mot->kd[i] * (mot->tvel[i] - qs[i]);
        rlink[nrows] = i; rsign[nrows] = 1.0; rjdi[nrows] = jdi;
        rrhs[nrows] = (rhs - qs[i]) * jdi;
        rlo[nrows] = -mot->max_imp[i]; rhi[nrows] = mot->max_imp[i]; rapp[nrows] = 0.0;
        nrows++;
    }
    /* contact rows (collision detection runs on the pre-step pose, like stepSimulation's first phase) */
    orc_contact_set cs;
    cs.n = 0;
    double cJ[ORC_MAXC][ORC_MAXL], cU[ORC_MAXC][ORC_MAXL], crhs[ORC_MAXC], cjdi[ORC_MAXC], capp[ORC_MAXC];
    if (obstacle != NULL) {
        double Rw[ORC_MAXL * 9], pw[ORC_MAXL * 3];
        orc_fk(m, q, Rw, pw);
        find_contacts(m, Rw, pw, obstacle, target, contact_thr, &cs);
        for (int c = 0; c < cs.n; c++) {
            contact_jacobian(m, Rw, pw, cs.link[c], cs.nrm[c], cs.pA[c], cJ[c]);
            double jmj = 0, rel = 0;
            for (int i = 0; i < nl; i++) {
                double u = 0;
                if (w.hasdof[i])
                    for (int k = 0; k < nl; k++)
                        if (w.hasdof[k]) u += col[k][i] * cJ[c][k];        /* (M^-1 J^T)_i, M^-1 symmetric */
                cU[c][i] = u;
                jmj += cJ[c][i] * u;
                rel += cJ[c][i] * qs[i];
            }
            cjdi[c] = jmj > 1e-12 ? 1.0 / jmj : 0.0;
            double d = cs.dist[c];
            double pos_err = d > 0 ? 0.0 : -d * m->erp / m->dt;
            double vel_err = -rel - (d > 0 ? d / m->dt : 0.0);
            crhs[c] = (pos_err + vel_err) * cjdi[c];
            capp[c] = 0.0;
        }
    }
    if (n_contacts) *n_contacts = cs.n;
    double dv[ORC_MAXL];
    memset(dv, 0, sizeof(dv));
    int it = 0;
    for (it = 0; it < m->iters; it++) {
        double resid = 0;
        for (int jj = 0; jj < nrows; jj++) {
            int r = (it & 1) ? jj : nrows - 1 - jj;
            int l = rlink[r];
            double delta = rrhs[r] - (rsign[r] * dv[l]) * rjdi[r];
            double sum = rapp[r] + delta;
            if (sum < rlo[r]) { delta = rlo[r] - rapp[r]; rapp[r] = rlo[r]; }
            else if (sum > rhi[r]) { delta = rhi[r] - rapp[r]; rapp[r] = rhi[r]; }
            else rapp[r] = sum;
            for (int i = 0; i < nl; i++)
                if (w.hasdof[i]) dv[i] += delta * rsign[r] * col[l][i];
            double dvel = delta / rjdi[r];
            if (dvel * dvel > resid) resid = dvel * dvel;
        }
        for (int c = 0; c < cs.n; c++) {                 /* normal contact rows, after the non-contact rows */
            if (cjdi[c] == 0.0) continue;
            double jdv = 0;
            for (int i = 0; i < nl; i++) jdv += cJ[c][i] * dv[i];
            double delta = crhs[c] - jdv * cjdi[c];
            double sum = capp[c] + delta;
            if (sum < 0.0) { delta = -capp[c]; capp[c] = 0.0; }
            else capp[c] = sum;
            for (int i = 0; i < nl; i++) dv[i] += delta * cU[c][i];
            double dvel = delta / cjdi[c];
            if (dvel * dvel > resid) resid = dvel * dvel;
        }
        if (resid <= m->resid_thresh || it >= m->iters - 1) { it++; break; }
    }
    for (int i = 0; i < nl; i++) {
        if (!w.hasdof[i]) { qd[i] = 0; continue; }
        double x = qs[i] + dv[i];
        if (x > m->max_vel) x = m->max_vel;
        if (x < -m->max_vel) x = -m->max_vel;
        qd[i] = x;
        q[i] += m->dt * x;                               /* stepPositionsMultiDof */
    }
    return it;
}

int orc_substep(const orc_model* m, const orc_motors* mot, double* q, double* qd) {
    return substep_impl(m, mot, q, qd, NULL, NULL, 0.0, NULL);
}

int orc_substep_contacts(const orc_model* m, const orc_motors* mot, double* q, double* qd, const double* obstacle,
                         const double* target, double contact_thr, int* n_contacts) {
    return substep_impl(m, mot, q, qd, obstacle, target, contact_thr, n_contacts);
}

/* ---------- closest distances (getClosestPoints restated for convex primitives) ---------- */
static double point_box_signed(const double* p, const double* h) {
    double o[3], mx = -1e300, s = 0;
    for (int a = 0; a < 3; a++) {
        o[a] = fabs(p[a]) - h[a];
        if (o[a] > mx) mx = o[a];
        if (o[a] > 0) s += o[a] * o[a];
    }
    return s > 0 ? sqrt(s) : mx;
}
static double point_segment(const double* p, const double* a, const double* b) {
    double ab[3], ap[3];
    for (int x = 0; x < 3; x++) { ab[x] = b[x] - a[x]; ap[x] = p[x] - a[x]; }
    double L2 = dot3(ab, ab), t = L2 > 0 ? dot3(ap, ab) / L2 : 0.0;
    if (t < 0) t = 0;
    if (t > 1) t = 1;
    double dd = 0;
    for (int x = 0; x < 3; x++) { double e = ap[x] - t * ab[x]; dd += e * e; }
    return sqrt(dd);
}
/* squared distance from point p0 + t*dir to the origin-centred box h */
static double seg_f(const double* p0, const double* dir, const double* h, double t) {
    double s = 0;
    for (int a = 0; a < 3; a++) {
        double o = fabs(p0[a] + t * dir[a]) - h[a];
        if (o > 0) s += o * o;
    }
    return s;
}
/* exact segment/box distance: f(t) is convex piecewise quadratic; examine every piece */
static double segment_box(const double* a, const double* b, const double* h) {
    double dir[3], bp[8];
    int nb = 0;
    for (int x = 0; x < 3; x++) dir[x] = b[x] - a[x];
    bp[nb++] = 0.0;
    bp[nb++] = 1.0;
    for (int x = 0; x < 3; x++) {
        if (fabs(dir[x]) < 1e-300) continue;
        double t1 = (h[x] - a[x]) / dir[x], t2 = (-h[x] - a[x]) / dir[x];
        if (t1 > 0 && t1 < 1) bp[nb++] = t1;
        if (t2 > 0 && t2 < 1) bp[nb++] = t2;
    }
    for (int i = 1; i < nb; i++) { /* insertion sort */
        double v = bp[i];
        int j = i - 1;
        while (j >= 0 && bp[j] > v) { bp[j + 1] = bp[j]; j--; }
        bp[j + 1] = v;
    }
    double best = 1e300;
    for (int i = 0; i < nb; i++) {
        double f = seg_f(a, dir, h, bp[i]);
        if (f < best) best = f;
        if (i + 1 < nb) { /* stationary point of the quadratic piece (bp[i], bp[i+1]) */
            double tm = 0.5 * (bp[i] + bp[i + 1]), A = 0, B = 0;
            for (int x = 0; x < 3; x++) {
                double xm = a[x] + tm * dir[x];
                if (fabs(xm) > h[x]) {
                    double sg = xm > 0 ? 1.0 : -1.0;
                    /* o = sg*(a + t dir) - h ; sum o^2 -> A t^2 + B t + C */
                    A += dir[x] * dir[x];
                    B += 2 * (sg * a[x] - h[x]) * sg * dir[x];
                }
            }
            if (A > 0) {
                double ts = -B / (2 * A);
                if (ts > bp[i] && ts < bp[i + 1]) {
                    double f2 = seg_f(a, dir, h, ts);
                    if (f2 < best) best = f2;
                }
            }
        }
    }
    return sqrt(best);
}

static double segment_box_argmin(const double* a, const double* b, const double* h, double* t_best) {
    double dir[3], bp[8];
    int nb = 0;
    for (int x = 0; x < 3; x++) dir[x] = b[x] - a[x];
    bp[nb++] = 0.0;
    bp[nb++] = 1.0;
    for (int x = 0; x < 3; x++) {
        if (fabs(dir[x]) < 1e-300) continue;
        double t1 = (h[x] - a[x]) / dir[x], t2 = (-h[x] - a[x]) / dir[x];
        if (t1 > 0 && t1 < 1) bp[nb++] = t1;
        if (t2 > 0 && t2 < 1) bp[nb++] = t2;
    }
    for (int i = 1; i < nb; i++) {
        double v = bp[i];
        int j = i - 1;
        while (j >= 0 && bp[j] > v) { bp[j + 1] = bp[j]; j--; }
        bp[j + 1] = v;
    }
    double best = 1e300, tb = 0.0;
    for (int i = 0; i < nb; i++) {
        double f = seg_f(a, dir, h, bp[i]);
        if (f < best) { best = f; tb = bp[i]; }
        if (i + 1 < nb) {
            double tm = 0.5 * (bp[i] + bp[i + 1]), A = 0, B = 0;
            for (int x = 0; x < 3; x++) {
                double xm = a[x] + tm * dir[x];
                if (fabs(xm) > h[x]) {
                    double sg = xm > 0 ? 1.0 : -1.0;
                    A += dir[x] * dir[x];
                    B += 2 * (sg * a[x] - h[x]) * sg * dir[x];
                }
            }
            if (A > 0) {
                double ts = -B / (2 * A);
                if (ts > bp[i] && ts < bp[i + 1]) {
                    double f2 = seg_f(a, dir, h, ts);
                    if (f2 < best) { best = f2; tb = ts; }
                }
            }
        }
    }
    *t_best = tb;
    return sqrt(best);
}

/* ---------- GJK distance between convex cores (btGjkPairDetector::getClosestPoints with
 * btVoronoiSimplexSolver restated; no EPA: a consumer on the path only compares the distance with a
 * threshold, so overlapping cores report 0 and the margins / radii make the result negative) ---------- */
typedef struct {
    int type;                    /* ORC_SHAPE_BOX or ORC_SHAPE_HULL */
    const double* dim;           /* box half extents */
    const double (*verts)[3];    /* hull vertices, shape frame */
    int nv;
    const double* R;             /* world <- shape */
    const double* p;
} orc_convex;

static void support_shape(const orc_convex* A, const double* dir_w, double* out_w) {
    double dl[3], loc[3], t[3];
    m3tv(A->R, dir_w, dl);
    if (A->type == ORC_SHAPE_BOX) {
        for (int a = 0; a < 3; a++) loc[a] = dl[a] >= 0 ? A->dim[a] : -A->dim[a];
    } else {
        double best = -1e300;
        int bi = 0;
        for (int i = 0; i < A->nv; i++) {
            double d = dot3(dl, A->verts[i]);
            if (d > best) { best = d; bi = i; }
        }
        for (int a = 0; a < 3; a++) loc[a] = A->verts[bi][a];
    }
    m3v(A->R, loc, t);
    for (int a = 0; a < 3; a++) out_w[a] = A->p[a] + t[a];
}

/* closest point of triangle (a, b, c) to the origin; mask = which of the three vertices support it
 * (Voronoi regions, Ericson "Real-Time Collision Detection" 5.1.5 == btVoronoiSimplexSolver::closestPtPointTriangle) */
static void tri_closest(const double* a, const double* b, const double* c, double* out, int* mask) {
    double ab[3], ac[3];
    for (int x = 0; x < 3; x++) { ab[x] = b[x] - a[x]; ac[x] = c[x] - a[x]; }
    double d1 = -dot3(ab, a), d2 = -dot3(ac, a);
    if (d1 <= 0 && d2 <= 0) { for (int x = 0; x < 3; x++) out[x] = a[x]; *mask = 1; return; }
    double d3 = -dot3(ab, b), d4 = -dot3(ac, b);
    if (d3 >= 0 && d4 <= d3) { for (int x = 0; x < 3; x++) out[x] = b[x]; *mask = 2; return; }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        double v = d1 / (d1 - d3);
        for (int x = 0; x < 3; x++) out[x] = a[x] + v * ab[x];
        *mask = 3; return;
    }
    double d5 = -dot3(ab, c), d6 = -dot3(ac, c);
    if (d6 >= 0 && d5 <= d6) { for (int x = 0; x < 3; x++) out[x] = c[x]; *mask = 4; return; }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        double w = d2 / (d2 - d6);
        for (int x = 0; x < 3; x++) out[x] = a[x] + w * ac[x];
        *mask = 5; return;
    }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int x = 0; x < 3; x++) out[x] = b[x] + w * (c[x] - b[x]);
        *mask = 6; return;
    }
    double den = 1.0 / (va + vb + vc), v = vb * den, w = vc * den;
    for (int x = 0; x < 3; x++) out[x] = a[x] + ab[x] * v + ac[x] * w;
    *mask = 7;
}

/* closest point of the simplex P[0..n) to the origin; P is reduced to the supporting sub-simplex.
 * Returns 1 when the origin lies inside a tetrahedron (the cores overlap). */
static int simplex_closest(double P[4][3], int* n, double* v) {
    if (*n == 1) { for (int x = 0; x < 3; x++) v[x] = P[0][x]; return 0; }
    if (*n == 2) {
        double ab[3];
        for (int x = 0; x < 3; x++) ab[x] = P[1][x] - P[0][x];
        double t = -dot3(P[0], ab), L2 = dot3(ab, ab);
        if (t <= 0 || L2 <= 0) { for (int x = 0; x < 3; x++) v[x] = P[0][x]; *n = 1; return 0; }
        if (t >= L2) { for (int x = 0; x < 3; x++) { v[x] = P[1][x]; P[0][x] = P[1][x]; } *n = 1; return 0; }
        for (int x = 0; x < 3; x++) v[x] = P[0][x] + (t / L2) * ab[x];
        return 0;
    }
    int mask = 0;
    if (*n == 3) {
        tri_closest(P[0], P[1], P[2], v, &mask);
    } else {
        /* tetrahedron: the closest point lies on a face whose plane separates the origin from the fourth vertex */
        static const int face[4][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 3, 2, 0}};
        double best = 1e300;
        int any = 0;
        for (int f = 0; f < 4; f++) {
            const double *a = P[face[f][0]], *b = P[face[f][1]], *c = P[face[f][2]], *d = P[face[f][3]];
            double ab[3], ac[3], nrm[3], ad[3];
            for (int x = 0; x < 3; x++) { ab[x] = b[x] - a[x]; ac[x] = c[x] - a[x]; ad[x] = d[x] - a[x]; }
            cross3(ab, ac, nrm);
            double sp = -dot3(a, nrm), sd = dot3(ad, nrm);      /* origin side, fourth-vertex side */
            if (sp * sd < 0 || sd * sd <= 1e-30 * dot3(nrm, nrm) * dot3(ad, ad)) {   /* outside, or a flat tetrahedron */
                double c3[3];
                int m3;
                tri_closest(a, b, c, c3, &m3);
                double dd = dot3(c3, c3);
                if (dd < best) {
                    best = dd;
                    any = 1;
                    for (int x = 0; x < 3; x++) v[x] = c3[x];
                    mask = 0;
                    for (int k = 0; k < 3; k++) if ((m3 >> k) & 1) mask |= 1 << face[f][k];
                }
            }
        }
        if (!any) { v[0] = v[1] = v[2] = 0; return 1; }
    }
    int k = 0;
    for (int i = 0; i < *n; i++)
        if ((mask >> i) & 1) { if (k != i) for (int x = 0; x < 3; x++) P[k][x] = P[i][x]; k++; }
    *n = k;
    return 0;
}

/* distance between the cores of two posed convex shapes; 0 when they overlap */
static double gjk_pair(const orc_convex* A, const orc_convex* B, int* iters_out) {
    double P[4][3], v[3], w[3], sa[3], sb[3], nd[3];
    int n = 0, it;
    for (int x = 0; x < 3; x++) v[x] = A->p[x] - B->p[x];
    if (dot3(v, v) < 1e-24) { v[0] = 1; v[1] = v[2] = 0; }
    for (it = 0; it < 64; it++) {
        for (int x = 0; x < 3; x++) nd[x] = -v[x];
        support_shape(A, nd, sa);
        support_shape(B, v, sb);
        for (int x = 0; x < 3; x++) w[x] = sa[x] - sb[x];      /* support of A - B along -v */
        if (it > 0) {
            double vv = dot3(v, v);
            if (vv - dot3(v, w) <= 1e-13 * vv) break;          /* no vertex of A - B is closer: v is the closest point */
            int dup = 0;
            for (int i = 0; i < n; i++)
                if (P[i][0] == w[0] && P[i][1] == w[1] && P[i][2] == w[2]) dup = 1;
            if (dup) break;
        }
        for (int x = 0; x < 3; x++) P[n][x] = w[x];
        n++;
        if (simplex_closest(P, &n, v) || dot3(v, v) < 1e-24) { if (iters_out) *iters_out = it + 1; return 0.0; }
    }
    if (iters_out) *iters_out = it;
    return sqrt(dot3(v, v));
}

/* shape A against an axis-aligned box (centre bc, half extents bh; bh = 0: a point) */
static double gjk_distance(const orc_convex* A, const double* bc, const double* bh, int* iters_out) {
    static const double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    orc_convex B = {ORC_SHAPE_BOX, bh, 0, 0, eye, bc};
    return gjk_pair(A, &B, iters_out);
}

/* core + radius of model shape s posed by (Rs, ps): sphere = point, capsule = segment (a box with two zero
 * half extents), box, hull (radius = margin) */
static double shape_core(const orc_model* m, int s, const double* Rs, const double* ps, double* half, orc_convex* out) {
    int t = m->s_type[s];
    double radius = 0;
    half[0] = half[1] = half[2] = 0;
    out->type = ORC_SHAPE_BOX; out->dim = half; out->verts = 0; out->nv = 0; out->R = Rs; out->p = ps;
    if (t == ORC_SHAPE_SPHERE) radius = m->s_dim[s][0];
    else if (t == ORC_SHAPE_CAPSULE) { half[2] = m->s_dim[s][1]; radius = m->s_dim[s][0]; }
    else if (t == ORC_SHAPE_BOX) { for (int a = 0; a < 3; a++) half[a] = m->s_dim[s][a]; }
    else { out->type = ORC_SHAPE_HULL; out->verts = m->verts + m->s_v0[s]; out->nv = m->s_vn[s]; radius = m->s_dim[s][0]; }
    return radius;
}

/* environment.py:394-412 + collision_detector.py:63-98: closest distance between every pair of links (min over
 * their shape pairs, saturated at 10); adjacent links and the diagonal are not queried by the reference: 10 */
void orc_self_distances(const orc_model* m, const double* q, double* out /*[nl][nl]*/) {
    double Rw[ORC_MAXL * 9], pw[ORC_MAXL * 3];
    double Rs[ORC_MAXS][9], ps[ORC_MAXS][3];
    int nl = m->nl;
    orc_fk(m, q, Rw, pw);
    for (int s = 0; s < m->ns; s++) {
        int l = m->s_link[s];
        double t[3];
        m3m(Rw + 9 * l, m->s_R[s], Rs[s]);
        m3v(Rw + 9 * l, m->s_p[s], t);
        for (int a = 0; a < 3; a++) ps[s][a] = pw[3 * l + a] + t[a];
    }
    for (int i = 0; i < nl * nl; i++) out[i] = 10.0;
    for (int sa = 0; sa < m->ns; sa++)
        for (int sb = sa + 1; sb < m->ns; sb++) {
            int i = m->s_link[sa], j = m->s_link[sb];
            if (i == j || i == j + 1 || j == i + 1) continue;
            double ha[3], hb[3];
            orc_convex A, B;
            double ra = shape_core(m, sa, Rs[sa], ps[sa], ha, &A), rb = shape_core(m, sb, Rs[sb], ps[sb], hb, &B);
            double d = gjk_pair(&A, &B, 0) - ra - rb;
            if (d < out[i * nl + j]) out[i * nl + j] = out[j * nl + i] = d;
        }
}

double orc_gjk_hull_box(const double* verts, int nv, const double* R, const double* p, const double* bc,
                        const double* bh, int* iters_out) {
    orc_convex A = {ORC_SHAPE_HULL, 0, (const double (*)[3])verts, nv, R, p};
    return gjk_distance(&A, bc, bh, iters_out);
}

void orc_distances(const orc_model* m, const double* q, const double* obstacle, const double* target,
                   double* link_obst, double* ee_target, double* ee_pos) {
    double Rw[ORC_MAXL * 9], pw[ORC_MAXL * 3];
    orc_fk(m, q, Rw, pw);
    for (int i = 0; i < m->nl; i++) link_obst[i] = 10.0;   /* no collision shape -> saturate */
    double ee = 10.0;
    for (int s = 0; s < m->ns; s++) {
        int l = m->s_link[s];
        double Rs[9], ps[3], t[3];
        m3m(Rw + 9 * l, m->s_R[s], Rs);                      /* world <- shape */
        m3v(Rw + 9 * l, m->s_p[s], t);
        for (int a = 0; a < 3; a++) ps[a] = pw[3 * l + a] + t[a];
        /* obstacle centre in shape frame */
        double rel[3], o_s[3], dist;
        for (int a = 0; a < 3; a++) rel[a] = obstacle[a] - ps[a];
        m3tv(Rs, rel, o_s);
        if (m->s_type[s] == ORC_SHAPE_SPHERE) dist = sqrt(dot3(o_s, o_s)) - m->s_dim[s][0];
        else if (m->s_type[s] == ORC_SHAPE_CAPSULE) {
            double a0[3] = {0, 0, -m->s_dim[s][1]}, b0[3] = {0, 0, m->s_dim[s][1]};
            dist = point_segment(o_s, a0, b0) - m->s_dim[s][0];
        } else if (m->s_type[s] == ORC_SHAPE_BOX) dist = point_box_signed(o_s, m->s_dim[s]);
        else { /* convex hull of a mesh: GJK between the hull and the sphere centre, minus the hull margin */
            static const double zero3[3] = {0, 0, 0};
            orc_convex A = {ORC_SHAPE_HULL, 0, m->verts + m->s_v0[s], m->s_vn[s], Rs, ps};
            dist = gjk_distance(&A, obstacle, zero3, 0) - m->s_dim[s][0];
        }
        dist -= m->obstacle_radius;
        if (dist < link_obst[l]) link_obst[l] = dist;
        if (l == m->ee_link) { /* vs axis-aligned target cube, in the cube frame */
            double c_t[3], dt_;
            for (int a = 0; a < 3; a++) c_t[a] = ps[a] - target[a];
            if (m->s_type[s] == ORC_SHAPE_SPHERE) dt_ = point_box_signed(c_t, m->target_half) - m->s_dim[s][0];
            else if (m->s_type[s] == ORC_SHAPE_CAPSULE) {
                double ax[3] = {Rs[2], Rs[5], Rs[8]}, a1[3], b1[3];
                for (int a = 0; a < 3; a++) {
                    a1[a] = c_t[a] - m->s_dim[s][1] * ax[a];
                    b1[a] = c_t[a] + m->s_dim[s][1] * ax[a];
                }
                double sd = segment_box(a1, b1, m->target_half);
                dt_ = sd - m->s_dim[s][0];
            } else if (m->s_type[s] == ORC_SHAPE_BOX) { /* box vs cube: GJK on the two boxes */
                orc_convex A = {ORC_SHAPE_BOX, m->s_dim[s], 0, 0, Rs, ps};
                dt_ = gjk_distance(&A, target, m->target_half, 0);
            } else {
                orc_convex A = {ORC_SHAPE_HULL, 0, m->verts + m->s_v0[s], m->s_vn[s], Rs, ps};
                dt_ = gjk_distance(&A, target, m->target_half, 0) - m->s_dim[s][0];
            }
            if (dt_ < ee) ee = dt_;
        }
    }
    *ee_target = ee;
    for (int a = 0; a < 3; a++) ee_pos[a] = pw[3 * m->ee_link + a];
}

/* environment.py:431-451 (state), :345-371 (reward), :311-343 (terminal) */
void orc_observe(const orc_model* m, const double* q, const double* qd, const double* obstacle,
                 const double* target, double* obs, double* reward, int* done) {
    double lo[ORC_MAXL], ee, eep[3];
    orc_distances(m, q, obstacle, target, lo, &ee, eep);
    int n = m->n_obs_joints;
    for (int i = 0; i < n; i++) { obs[i] = q[i]; obs[n + i] = qd[i]; }
    for (int a = 0; a < 3; a++) { obs[2 * n + a] = eep[a]; obs[2 * n + 3 + a] = target[a]; obs[2 * n + 6 + a] = obstacle[a]; }
    int hit = 0;
    for (int i = 0; i < m->nl; i++) if (lo[i] < 0.0) hit = 1;
    int goal = ee < 0.05;
    if (goal) *reward = 250.0;
    else if (hit) *reward = -1000.0;
    else *reward = -(ee - 0.05);
    *done = (hit || goal) ? 1 : 0;
}

void orc_batch_step(const orc_model* m, const orc_motors* mt, const int* act_joint, int n_act, int n,
                    double* q, double* qd, const double* actions, double max_force, const double* obstacle,
                    const double* target, double* obs, double* reward, int* done, int* iters_out, int nthreads) {
    int nl = m->nl, S = 9 + 2 * m->n_obs_joints;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (int e = 0; e < n; e++) {
        orc_motors mot = *mt;
        for (int k = 0; k < n_act; k++) {   /* VELOCITY_CONTROL: kp 0, kd 1, maxImpulse = force*dt */
            int j = act_joint[k];
            mot.kp[j] = 0; mot.kd[j] = 1; mot.tpos[j] = 0; mot.tvel[j] = actions[e * n_act + k];
            mot.max_imp[j] = max_force * m->dt;
        }
        int it = orc_substep(m, &mot, q + e * nl, qd + e * nl);
        if (iters_out) iters_out[e] = it;
        orc_observe(m, q + e * nl, qd + e * nl, obstacle + 3 * e, target + 3 * e, obs + e * S, reward + e, done + e);
    }
    (void)nthreads;
}

void orc_batch_reset(const orc_model* m, const orc_motors* mt, int n_init, int n, double* q, double* qd,
                     const double* init_targets, int nsub, int nthreads) {
    int nl = m->nl;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (int e = 0; e < n; e++) {
        orc_motors mot = *mt;
        for (int j = 0; j < n_init; j++) { /* POSITION_CONTROL defaults: kp .1 kd 1 force 1e5 */
            mot.kp[j] = 0.1; mot.kd[j] = 1.0; mot.tpos[j] = init_targets[e * n_init + j]; mot.tvel[j] = 0;
            mot.max_imp[j] = 100000.0 * m->dt;
        }
        for (int s = 0; s < nsub; s++) orc_substep(m, &mot, q + e * nl, qd + e * nl);
    }
    (void)nthreads;
}

/* orc_batch_step with the contact rows (contact_thr > 0) */
void orc_batch_step_contacts(const orc_model* m, const orc_motors* mt, const int* act_joint, int n_act, int n, double* q,
                             double* qd, const double* actions, double max_force, const double* obstacle, const double* target,
                             double contact_thr, double* obs, double* reward, int* done, int* iters_out, int* ncontacts_out,
                             int nthreads) {
    int nl = m->nl, S = 9 + 2 * m->n_obs_joints;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (int e = 0; e < n; e++) {
        orc_motors mot = *mt;
        for (int k = 0; k < n_act; k++) {
            int j = act_joint[k];
            mot.kp[j] = 0; mot.kd[j] = 1; mot.tpos[j] = 0; mot.tvel[j] = actions[e * n_act + k];
            mot.max_imp[j] = max_force * m->dt;
        }
        int nc = 0;
        int it = substep_impl(m, &mot, q + e * nl, qd + e * nl, obstacle + 3 * e, target + 3 * e, contact_thr, &nc);
        if (iters_out) iters_out[e] = it;
        if (ncontacts_out) ncontacts_out[e] = nc;
        orc_observe(m, q + e * nl, qd + e * nl, obstacle + 3 * e, target + 3 * e, obs + e * S, reward + e, done + e);
    }
    (void)nthreads;
}

void orc_batch_reset_contacts(const orc_model* m, const orc_motors* mt, int n_init, int n, double* q, double* qd,
                              const double* init_targets, int nsub, const double* obstacle, const double* target,
                              double contact_thr, int nthreads) {
    int nl = m->nl;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (int e = 0; e < n; e++) {
        orc_motors mot = *mt;
        for (int j = 0; j < n_init; j++) {
            mot.kp[j] = 0.1; mot.kd[j] = 1.0; mot.tpos[j] = init_targets[e * n_init + j]; mot.tvel[j] = 0;
            mot.max_imp[j] = 100000.0 * m->dt;
        }
        for (int s = 0; s < nsub; s++)
            substep_impl(m, &mot, q + e * nl, qd + e * nl, obstacle + 3 * e, target + 3 * e, contact_thr, NULL);
    }
    (void)nthreads;
}

int orc_sizeof_model(void) { return (int)sizeof(orc_model); }
int orc_sizeof_motors(void) { return (int)sizeof(orc_motors); }
