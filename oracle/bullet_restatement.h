/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under robotic_manipulator_rloa_b200/ may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: PyBullet (pybullet, unpinned in /root/reference/pyproject.toml:24) and its
 * pybullet_data assets are absent from this image and the reference's tests mock every p.* call
 * (tests/robotic_manipulator_rloa/environment/test_environment.py:79-97), so no golden physics
 * vector exists.  This file restates, in double precision, the published Bullet3 algorithm that
 * runs under the reference's call sites:
 *   environment.py:453-485  Environment.step      (setJointMotorControl2 x n, stepSimulation)
 *   environment.py:264-309  Environment.reset     (POSITION_CONTROL drive, 50 x stepSimulation)
 *   environment.py:311-392,414-451  reward / terminal / state logic
 *   collision_detector.py:33-61     closest-distance reduction, saturate at 10.0
 * Bullet algorithm restated (from Bullet3 src/BulletDynamics/Featherstone):
 *   btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof  -> orc_aba()
 *   btMultiBody::calcAccelerationDeltasMultiDof                        -> orc_unit_response()
 *   btMultiBodyJointMotor / btMultiBodyJointLimitConstraint rows       -> orc_substep()
 *   btMultiBodyConstraintSolver::solveSingleIteration (PGS)            -> orc_substep()
 *   btMultiBody::stepPositionsMultiDof                                 -> orc_substep()
 * and, for getClosestPoints on mesh (convex hull) shapes and box end-effector shapes
 * (src/BulletCollision/NarrowPhaseCollision):
 *   btGjkPairDetector::getClosestPoints + btVoronoiSimplexSolver       -> orc_gjk_distance()
 */
#ifndef ORC_BULLET_RESTATEMENT_H
#define ORC_BULLET_RESTATEMENT_H

#define ORC_MAXL 32   /* links (== pybullet getNumJoints) */
#define ORC_MAXS 64   /* collision shapes */
#define ORC_MAXC 4    /* contact rows per arm: the first four pairs in shape order (== kMaxContacts of the CUDA simulator) */
#define ORC_MAXV 16384 /* convex-hull vertices, all hull shapes together */

enum { ORC_FIXED = 0, ORC_REVOLUTE = 1, ORC_PRISMATIC = 2 };
enum { ORC_SHAPE_SPHERE = 1, ORC_SHAPE_CAPSULE = 2, ORC_SHAPE_BOX = 3, ORC_SHAPE_HULL = 4 };

typedef struct {
    int nl;
    int parent[ORC_MAXL];          /* -1 = the fixed base */
    int jtype[ORC_MAXL];
    double E0[ORC_MAXL][9];        /* child COM frame <- parent COM frame at q = 0, row-major */
    double e[ORC_MAXL][3];         /* parent COM -> joint pivot, in parent COM frame */
    double d[ORC_MAXL][3];         /* joint pivot -> child COM, in child COM frame */
    double axis[ORC_MAXL][3];      /* unit joint axis in child COM frame */
    double mass[ORC_MAXL];
    double inertia[ORC_MAXL][3];   /* principal moments (COM frame is principal) */
    double damping[ORC_MAXL];      /* <dynamics damping> */
    double lower[ORC_MAXL], upper[ORC_MAXL];
    int has_limit[ORC_MAXL];
    double base_R[9], base_p[3];   /* world <- base COM frame */
    double lin_damp, ang_damp;     /* btMultiBody m_linearDamping / m_angularDamping (0.04) */
    double gravity[3];
    double dt;                     /* 1/240 */
    int iters;                     /* 50 */
    double resid_thresh;           /* 1e-7 on the squared row velocity change */
    double erp;                    /* 0.2 */
    double max_vel;                /* 100 */
    double limit_max_impulse;      /* 100 */
    /* collision primitives, poses relative to the owning link's COM frame */
    int ns;
    int s_link[ORC_MAXS];
    int s_type[ORC_MAXS];
    double s_R[ORC_MAXS][9];       /* link COM frame <- shape frame */
    double s_p[ORC_MAXS][3];
    double s_dim[ORC_MAXS][3];     /* sphere: r,-,- ; capsule (local z): r, half_len,- ; box: half extents ;
                                      hull: collision margin (pybullet's URDF importer sets 0.001),-,- */
    double obstacle_radius;
    double target_half[3];
    int ee_link;
    int n_obs_joints;              /* len(involved_joints): get_state reads joints 0..n-1 */
    /* convex-hull shapes (mesh collision geometry: Bullet keeps the convex hull of the mesh vertices,
     * SURVEY.md A.5): vertices in the shape frame, shape s owns verts[s_v0[s] .. s_v0[s] + s_vn[s]) */
    int s_v0[ORC_MAXS], s_vn[ORC_MAXS];
    int nv;
    double verts[ORC_MAXV][3];
} orc_model;

/* per-joint motor settings (persist across steps, like btMultiBodyJointMotor) */
typedef struct {
    double kp[ORC_MAXL], kd[ORC_MAXL], tpos[ORC_MAXL], tvel[ORC_MAXL], max_imp[ORC_MAXL];
} orc_motors;

#ifdef __cplusplus
extern "C" {
#endif

void orc_fk(const orc_model* m, const double* q, double* Rw /*[nl][9]*/, double* pw /*[nl][3]*/);
void orc_aba(const orc_model* m, const double* q, const double* qd, const double* tau_ext, double* qdd);
void orc_minv(const orc_model* m, const double* q, double* Minv /*[nl*nl]*/);
void orc_crba(const orc_model* m, const double* q, double* M /*[nl*nl]*/);
void orc_rnea_bias(const orc_model* m, const double* q, const double* qd, double* bias);
int  orc_substep(const orc_model* m, const orc_motors* mot, double* q, double* qd);
/* the same with NORMAL contact rows against the obstacle sphere and the target cube (friction not restated) */
int  orc_substep_contacts(const orc_model* m, const orc_motors* mot, double* q, double* qd, const double* obstacle,
                          const double* target, double contact_thr, int* n_contacts);
void orc_distances(const orc_model* m, const double* q, const double* obstacle, const double* target,
                   double* link_obst /*[nl]*/, double* ee_target /*[1]*/, double* ee_pos /*[3]*/);
void orc_observe(const orc_model* m, const double* q, const double* qd, const double* obstacle,
                 const double* target, double* obs, double* reward, int* done);
/* distance between the convex hull of nv points (pose world <- shape: R row-major, p) and an axis-aligned box
 * (centre bc, half extents bh; bh = 0 -> a point); 0 when they overlap */
double orc_gjk_hull_box(const double* verts /*[nv][3]*/, int nv, const double* R, const double* p,
                        const double* bc, const double* bh, int* iters_out);
/* link <-> link closest distances, [nl][nl] (10 on the diagonal, for adjacent links and for links without shapes) */
void orc_self_distances(const orc_model* m, const double* q, double* out);
/* batched helpers (OpenMP over envs); q, qd are [n][nl] */
void orc_batch_step(const orc_model* m, const orc_motors* mot_template, const int* act_joint, int n_act,
                    int n, double* q, double* qd, const double* actions /*[n][n_act]*/, double max_force,
                    const double* obstacle /*[n][3]*/, const double* target /*[n][3]*/,
                    double* obs, double* reward, int* done, int* iters_out, int nthreads);
void orc_batch_reset(const orc_model* m, const orc_motors* mot_template, int n_init, int n, double* q,
                     double* qd, const double* init_targets /*[n][n_init]*/, int nsub, int nthreads);
int orc_sizeof_model(void);
int orc_sizeof_motors(void);

#ifdef __cplusplus
}
#endif
#endif
