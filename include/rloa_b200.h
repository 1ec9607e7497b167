/*
 * rloa_b200.h — C ABI of librloa_b200.so, the B200 (sm_100a) native library behind the
 * robotic_manipulator_rloa hot path.
 *
 * The reference (JavierMtz5/robotic_manipulator_rloa) is pure Python and has no FFI of its own: the
 * native work on its hot path is done by two third-party wheels, PyBullet and PyTorch.  Every entry
 * point below therefore cites the reference call site whose native work it replaces
 * (paths relative to /root/reference/robotic_manipulator_rloa/).
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++ or torch types;
 *   - every function returns 0 on success, a negative rloa_status on failure, and leaves a
 *     thread-local message readable through rloa_last_error();
 *   - all `const float*` / `float*` data arguments are DEVICE pointers unless the name ends in
 *     `_host`; buffers are caller-owned, row-major contiguous, fp32 unless stated;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are
 *     asynchronous on that stream, contain no hidden synchronisation and can be captured into a
 *     CUDA graph (exceptions: *_create / *_destroy / *_host getters, which synchronise);
 *   - a handle is bound to the CUDA device that was current when it was created and is not
 *     thread-safe: one host thread (one rank) per GPU drives it.
 */
#ifndef RLOA_B200_H
#define RLOA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLOA_MAX_LINKS 32      /* links == PyBullet getNumJoints() */
#define RLOA_MAX_SHAPES 32     /* collision shapes (sphere / capsule / box / convex hull) */
#define RLOA_MAX_HULL_VERTS 16384 /* vertices of all convex-hull shapes of a model together */
#define RLOA_MAX_DOF 16        /* movable joints (PGS rows = dof + active limit rows <= 32) */

typedef enum {
    RLOA_OK = 0,
    RLOA_ERR_INVALID = -1,     /* bad argument / unsupported model */
    RLOA_ERR_CUDA = -2,        /* a CUDA runtime call failed */
    RLOA_ERR_NO_DEVICE = -3    /* no sm_100 device visible */
} rloa_status;

enum { RLOA_JOINT_FIXED = 0, RLOA_JOINT_REVOLUTE = 1, RLOA_JOINT_PRISMATIC = 2 };
enum { RLOA_SHAPE_SPHERE = 1, RLOA_SHAPE_CAPSULE = 2, RLOA_SHAPE_BOX = 3, RLOA_SHAPE_HULL = 4 };

const char* rloa_last_error(void);
int rloa_version(void);
/* number of kernels this library has launched in the calling process (bench.py gpu_launches) */
uint64_t rloa_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Robot model.  Replaces p.loadURDF / p.loadSDF / p.getNumJoints (environment/environment.py:224-238)
 * and the obstacle / target bodies (environment.py:252-255).  All arrays are HOST pointers, double
 * precision, link frames at the centre of mass with principal axes (Bullet's btMultiBody layout).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t nl;                 /* links / joints */
    const int32_t* parent;      /* [nl] -1 = fixed base; links numbered depth first (parent < child) */
    const int32_t* jtype;       /* [nl] RLOA_JOINT_* */
    const double* E0;           /* [nl][9] child COM frame <- parent COM frame at q = 0 */
    const double* e;            /* [nl][3] parent COM -> joint pivot, parent COM frame */
    const double* d;            /* [nl][3] joint pivot -> child COM, child COM frame */
    const double* axis;         /* [nl][3] unit joint axis, child COM frame */
    const double* mass;         /* [nl] */
    const double* inertia;      /* [nl][3] principal moments */
    const double* damping;      /* [nl] joint damping */
    const double* lower;        /* [nl] */
    const double* upper;        /* [nl] */
    const int32_t* has_limit;   /* [nl] */
    double base_R[9];           /* world <- base COM frame */
    double base_p[3];
    double lin_damp, ang_damp;  /* Bullet per-link velocity drag (0.04) */
    double gravity[3];
    double dt;                  /* 1/240 */
    int32_t iters;              /* PGS iterations (50) */
    double resid_thresh;        /* 1e-7, on the squared row velocity change */
    double erp;                 /* 0.2 */
    double max_vel;             /* 100 */
    double limit_max_impulse;   /* 100 */
    int32_t ns;                 /* collision primitives */
    const int32_t* s_link;      /* [ns] owning link */
    const int32_t* s_type;      /* [ns] RLOA_SHAPE_* */
    const double* s_R;          /* [ns][9] link COM frame <- shape frame */
    const double* s_p;          /* [ns][3] */
    const double* s_dim;        /* [ns][3] sphere r,-,- | capsule r,half_len,- (local z) | box half extents |
                                   hull: collision margin,-,- (pybullet's importer: 0.001) */
    double obstacle_radius;     /* sphere_small.urdf x 2.5 -> 0.075 */
    double target_half[3];      /* cube_small.urdf -> 0.025 each */
    int32_t ee_link;            /* endeffector_index */
    int32_t n_obs_joints;       /* len(involved_joints); get_state reads joints 0..n-1 (environment.py:442-444) */
    /* RLOA_SHAPE_HULL: mesh / cylinder collision geometry, which p.loadURDF / p.loadSDF keep as the convex hull of
     * the vertices (environment.py:228-233).  Shape s owns verts[s_vert_first[s] .. + s_vert_count[s]), given in
     * the shape frame; closest distances come from GJK over the vertex cloud (p.getClosestPoints,
     * utils/collision_detector.py:47-52).  n_verts = 0: the three pointers are ignored. */
    int32_t n_verts;
    const int32_t* s_vert_first;  /* [ns] */
    const int32_t* s_vert_count;  /* [ns] */
    const double* verts;          /* [n_verts][3] */
} rloa_model_desc;

typedef struct rloa_model rloa_model;
int rloa_model_create(const rloa_model_desc* desc, rloa_model** out);
void rloa_model_destroy(rloa_model* m);

/* ------------------------------------------------------------------------------------------------
 * Batched simulator: n_envs independent arms, one THREAD per arm (struct-of-arrays state, coalesced across the
 * warp); one step = dynamics, M^-1 columns (warp = 32 arms x one dof) and solve kernels — DESIGN.md section 3.
 * Replaces the PyBullet world owned by Environment (environment.py:207-210) and
 * p.setJointMotorControl2 / p.stepSimulation / p.getJointState / p.getLinkState /
 * p.getClosestPoints on Environment.step / reset / get_state / get_reward / is_terminal_state
 * (environment.py:264-309, 311-371, 431-485; utils/collision_detector.py:33-61).
 * ---------------------------------------------------------------------------------------------- */
typedef struct rloa_sim rloa_sim;
int rloa_sim_create(const rloa_model* m, int32_t n_envs, rloa_sim** out);
void rloa_sim_destroy(rloa_sim* s);
int rloa_sim_num_envs(const rloa_sim* s);
int rloa_sim_obs_size(const rloa_sim* s);      /* 9 + 2 * n_obs_joints */

/* per-env target / obstacle positions, [n_envs][3] (environment.py:212-213 broadcast per env) */
int rloa_sim_set_task(rloa_sim* s, const float* target, const float* obstacle, void* stream);

/* joint coordinates, [n_envs][nl] (fixed joints carry 0); for parity tests and checkpoints */
int rloa_sim_set_state(rloa_sim* s, const float* q, const float* qd, void* stream);
int rloa_sim_get_state(const rloa_sim* s, float* q, float* qd, void* stream);

/* motor table, [n_envs][nl] each: what setJointMotorControl2 leaves behind in Bullet
 * (kd is 1 in every mode the reference uses).  NULL pointers are skipped. */
int rloa_sim_set_motors(rloa_sim* s, const float* kp, const float* target_pos, const float* target_vel,
                        const float* max_impulse, void* stream);
/* back to the load-time state: q = qd = 0, default motors (velocity target 0, max impulse 1) */
int rloa_sim_clear(rloa_sim* s, void* stream);

typedef struct {
    int32_t n_act;                          /* len(involved_joints) */
    int32_t act_joint[RLOA_MAX_LINKS];      /* involved_joints */
    int32_t n_fixed;                        /* len(fixed_joints) */
    int32_t fixed_joint[RLOA_MAX_LINKS];    /* fixed_joints */
    float max_force;                        /* environment.py:469 force=max_force */
    float target_threshold;                 /* 0.05  (environment.py:311, 364) */
    float obstacle_threshold;               /* 0.0   (environment.py:311, 368) */
    float contact_threshold;                /* contact rows against the obstacle sphere / target cube, which the reference
                                               loads as collidable fixed bodies (environment.py:252-255): Bullet's contact
                                               breaking threshold 0.02; 0 switches the rows off (free dynamics) */
} rloa_step_config;

/* Environment.step for every env (environment.py:453-485):
 *   actions [n_envs][n_act] -> obs [n_envs][S], reward [n_envs], done [n_envs] (uint8).
 * active (uint8 [n_envs]) may be NULL; envs with active == 0 are left untouched.
 * An env with pending reset sub-steps (rloa_sim_begin_reset) runs one of them instead of an action step
 * and reports valid = 0, reward = 0, done = 0; valid (uint8 [n_envs]) may be NULL. */
int rloa_sim_step(rloa_sim* s, const rloa_step_config* cfg, const float* actions, const uint8_t* active,
                  float* obs, float* reward, uint8_t* done, uint8_t* valid, void* stream);

/* Contact breaking threshold used by rloa_sim_reset's sub-steps and by rloa_sim_prepare's collision phase (Environment.reset
 * steps the same world as step, contact rows included; rloa_sim_step takes its threshold from the config and repeats the
 * collision phase when a prepare ran with another one).  0 = off. */
int rloa_sim_set_contacts(rloa_sim* s, float contact_threshold);
/* Number of contact rows the latest collision phase found per env (Bullet's contact manifold points of the arm against the two
 * bodies of environment.py:252-255, at most 4).  The collision phase runs on the pose a stepSimulation starts from: inside
 * rloa_sim_prepare (so after a prepare the counts are those of the NEXT step) or at the start of rloa_sim_step.
 * counts: int32 [n_envs] DEVICE. */
int rloa_sim_contact_counts(rloa_sim* s, int32_t* counts, void* stream);

/* Software pipelining of the step.  A stepSimulation has an action-independent half (forward kinematics,
 * articulated-body factorisation, free accelerations, M^-1: two of the three kernels) that only needs the
 * joint state the PREVIOUS step left behind.  rloa_sim_prepare runs that half for the next step on a stream
 * owned by the simulator, forked from `stream` at the call and joined by the next rloa_sim_step (or by
 * rloa_sim_join, e.g. before the end of a CUDA-graph capture), which then launches the solve only.  Same
 * arithmetic, same results; anything that changes (q, qd) in between (set_state, reset, clear) drops the
 * prepared half. */
int rloa_sim_prepare(rloa_sim* s, void* stream);
int rloa_sim_join(rloa_sim* s, void* stream);

/* Lock-step asynchronous Environment.reset for the vectorised loop: arms the POSITION_CONTROL motors of
 * the masked envs (joints 0..n_init-1 -> init_targets) and schedules n_substeps (50) reset sub-steps,
 * executed one per rloa_sim_step launch, so envs that keep acting never wait for a resetting one
 * (environment.py:295-301).  The obs written by the launch that finishes the last sub-step is the new
 * episode's first state. */
int rloa_sim_begin_reset(rloa_sim* s, const uint8_t* mask, const float* init_targets, int32_t n_init,
                         int32_t n_substeps, void* stream);

/* rloa_sim_begin_reset with the start pose drawn on the device (environment.py:284-293):
 * target_j = pos[j] + var[j] * U(-1, 1) for j < n_init, Philox4x32-10 keyed by (seed, *tick, env, j).
 * pos / var are DEVICE arrays [n_init] (NULL = zeros); tick is a DEVICE counter (may be NULL). */
int rloa_sim_begin_reset_random(rloa_sim* s, const uint8_t* mask, const float* pos, const float* var,
                                int32_t n_init, int32_t n_substeps, uint64_t seed, const uint64_t* tick,
                                void* stream);

/* Environment.reset for the envs with mask != 0 (environment.py:264-309):
 * POSITION_CONTROL targets init_targets [n_envs][n_init] on joints 0..n_init-1, then n_substeps
 * (50) simulation steps; writes the new state of those envs into obs [n_envs][S].
 * mask may be NULL (= all envs). */
int rloa_sim_reset(rloa_sim* s, const uint8_t* mask, const float* init_targets, int32_t n_init,
                   int32_t n_substeps, float* obs, void* stream);

/* get_state without stepping (environment.py:431-451), plus the distances consumed by
 * get_reward / is_terminal_state: link_obstacle [n_envs][nl] (10.0 when a link has no collision
 * shape, collision_detector.py:56-57) and ee_target [n_envs].  Any output may be NULL. */
int rloa_sim_observe(const rloa_sim* s, float* obs, float* link_obstacle, float* ee_target, void* stream);

/* get_manipulator_collisions_with_itself (environment.py:394-412 -> CollisionDetector.compute_collisions_in_manipulator,
 * utils/collision_detector.py:63-98): link_link [n_envs][nl][nl], entry [i][j] = closest distance between link i and
 * link j (p.getClosestPoints(body, body, 10, linkIndexA=i, linkIndexB=j), minimum over the shape pairs).  Entries the
 * reference does not query — the diagonal and adjacent indices |i - j| = 1 — and pairs with a shapeless link hold 10.0.
 * Not on the step path (Environment.step never passes consider_autocollision=True). */
int rloa_sim_self_distances(const rloa_sim* s, float* link_link, void* stream);

/* diagnostics: PGS iterations used by the last substep of every env, [n_envs] int32 */
int rloa_sim_last_iterations(const rloa_sim* s, int32_t* iters, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Episode bookkeeping for the vectorised rollout loop
 * (naf_components/naf_algorithm.py:243-277, rl_framework.py:336-354).
 * score += reward; frame += 1; finished = done | (frame >= frames);
 * finished envs get their (score, frame, reward) appended to the episode log ring and are flagged
 * in reset_mask; their counters restart at 0.  `active` (may be NULL) selects the envs that produced a
 * transition this step; `transitions` (device int64, may be NULL) accumulates their number; `tick`
 * (device uint64, may be NULL) is incremented once per call — the device-resident loop counter that
 * rloa_naf_act / rloa_replay_sample read through their *_offset arguments, so that a CUDA graph of
 * the whole loop iteration draws fresh random numbers on every replay.
 * ---------------------------------------------------------------------------------------------- */
int rloa_episode_update(int32_t n_envs, int32_t frames, const float* reward, const uint8_t* done,
                        const uint8_t* active, float* score, int32_t* frame, uint8_t* reset_mask,
                        float* log_score, int32_t* log_frame, float* log_last_reward, int32_t* log_env,
                        int32_t log_capacity, int32_t* log_count, int64_t* transitions, uint64_t* tick,
                        void* stream);

/* rloa_episode_update followed by rloa_sim_begin_reset_random(mask = the envs that just finished) in one
 * launch: the steady-state tail of the vectorised loop (naf_algorithm.py:263-277 + environment.py:284-301).
 * tick_next (device uint64): NULL or == tick -> *tick is incremented in place, as in rloa_episode_update.
 * A different pointer -> *tick_next = *tick + 1 and *tick is left alone (a ping-pong pair of counters), so the
 * readers of *tick on the caller's stream (rloa_replay_sample, ...) may run concurrently: while a rloa_sim_prepare
 * is in flight the launch then goes behind it on the simulator's side stream and completes no later than
 * rloa_sim_join / the next rloa_sim_step; its outputs (score, frame, logs, reset_mask, *tick_next) must not be read
 * before that. */
int rloa_episode_update_reset(rloa_sim* s, int32_t frames, const float* reward, const uint8_t* done,
                              const uint8_t* active, float* score, int32_t* frame, uint8_t* reset_mask,
                              float* log_score, int32_t* log_frame, float* log_last_reward, int32_t* log_env,
                              int32_t log_capacity, int32_t* log_count, int64_t* transitions, uint64_t* tick,
                              uint64_t* tick_next, const float* pos, const float* var, int32_t n_init,
                              int32_t n_substeps, uint64_t seed, void* stream);

/* ------------------------------------------------------------------------------------------------
 * NAF network (naf_components/naf_neural_network.py:8-123).  Parameters live in caller-owned
 * device tensors (the torch nn.Module owns them, so state_dict()/load_state_dict() keep the
 * checkpoint layout of naf_algorithm.py:280-289); the struct below just carries the pointers.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t state_size, action_size, hidden;     /* S, A, H (H = 256 in the reference, rl_framework.py:452) */
    float *w1, *b1;                              /* input_layer   [H][S], [H] */
    float *bn1_w, *bn1_b, *bn1_mean, *bn1_var;   /* bn1 affine + running stats [H] */
    int64_t* bn1_batches;                        /* bn1.num_batches_tracked */
    float *w2, *b2;                              /* hidden_layer  [H][H], [H] */
    float *bn2_w, *bn2_b, *bn2_mean, *bn2_var;
    int64_t* bn2_batches;
    float *w_mu, *b_mu;                          /* action_values [A][H], [A] */
    float *w_v, *b_v;                            /* value         [1][H], [1] */
    float *w_l, *b_l;                            /* matrix_entries [A(A+1)/2][H], [..] */
} rloa_naf_params;

typedef struct rloa_naf_ws rloa_naf_ws;          /* activations / gradients / optimiser workspace */
int rloa_naf_ws_create(int32_t state_size, int32_t action_size, int32_t hidden, int32_t max_batch,
                       rloa_naf_ws** out);
void rloa_naf_ws_destroy(rloa_naf_ws* ws);
/* 0 = fp32 CUDA-core trunk (reference-exact numerics), 1 = tcgen05 tensor-core trunk
 * (bf16 operands, fp32 TMEM accumulation; looser bound) */
int rloa_naf_ws_set_trunk(rloa_naf_ws* ws, int32_t mode);
/* Diagnostics of the fused tensor-core learn kernel (csrc/naf_learn_cluster.cu, trunk mode 1, batch <= 1024).
 * buffer != NULL: the next learn calls dump their intermediates as fp32 — six sections of 1024 x 256 floats:
 * z1 | z2 | dzh (row stride 64) | dz2 | da1 | dz1 — so each stage can be checked against autograd.
 * stamps != NULL: 16 x 32 int64 clock64() phase stamps, one row per CTA (tools/learn_cluster_profile.py).
 * NULL switches either off. */
int rloa_naf_ws_set_debug(rloa_naf_ws* ws, float* buffer, int64_t* stamps);

/* The trunk's hidden layer on its own (naf_neural_network.py:77-78):
 *   z2 [B][H] = relu(z1 * scale + shift) @ w2^T + b2,   scale / shift [H] = folded BatchNorm1 coefficients.
 * Runs on the path selected by rloa_naf_ws_set_trunk (fp32 CUDA cores, or tcgen05 with bf16 operands and
 * fp32 TMEM accumulation); exposed so the tensor-core kernel can be checked in isolation. */
int rloa_naf_hidden_layer(rloa_naf_ws* ws, const float* z1, const float* scale, const float* shift,
                          const float* w2, const float* b2, float* z2, int32_t batch, void* stream);

/* Self-test of the tcgen05 operand layouts the tensor-core learn path relies on (csrc/umma_probe.cu): one 128-row bf16
 * tile in the K-major SWIZZLE_128B layout read both K-major and, untransposed, MN-major.  a, b, c are fp32 row-major
 * device arrays:  mode 0: c[128][256] = a[128][256] b[256][256]^T;  1: c[128][256] = a[128][256] b[256][256];
 * 2: c[256][256] = a[128][256]^T b[128][256];  3: c[256][64] = a[128][256]^T b[128][64]  (operands rounded to bf16). */
int rloa_umma_probe(int32_t mode, const float* a, const float* b, float* c, void* stream);

/* NAF.forward (naf_neural_network.py:56-123) without the sampling tail:
 * mu [B][A], pdiag [B][A] (= diag of P = L o L^T, i.e. exp(2 tanh z_kk)), V [B], and when
 * action != NULL, Q [B] = -1/2 sum_k P_kk (u_k - mu_k)^2 + V.  train_mode != 0 uses batch
 * statistics and updates the running stats / num_batches_tracked like nn.BatchNorm1d.
 * trunc_action != 0 applies the reference's .long() cast to the actions (replay_buffer.py:60). */
int rloa_naf_forward(rloa_naf_ws* ws, const rloa_naf_params* p, const float* states, const float* action,
                     int32_t batch, int32_t train_mode, int32_t trunc_action, float* mu, float* pdiag,
                     float* q, float* v, void* stream);

/* NAFAgent.act for a batch of states (naf_algorithm.py:158-178 + naf_neural_network.py:119-121):
 * eval-mode forward, action = clamp(mu + exp(-tanh z_kk) * eps, -1, 1), eps ~ N(0,1) from
 * Philox4x32-10 keyed by (seed, step + *step_offset, row); step_offset is a DEVICE counter and may be
 * NULL.  noise_scale = 0 gives the mean action. */
int rloa_naf_act(rloa_naf_ws* ws, const rloa_naf_params* p, const float* states, int32_t batch,
                 uint64_t seed, uint64_t step, const uint64_t* step_offset, float noise_scale, float* actions,
                 void* stream);

typedef struct {
    float gamma, tau, lr;
    float beta1, beta2, eps;       /* Adam defaults 0.9, 0.999, 1e-8 (naf_algorithm.py:83) */
    float clip_norm;               /* 1.0 (naf_algorithm.py:209) */
    int32_t trunc_action;          /* 1 = reference .long() cast (replay_buffer.py:60) */
    int32_t use_done_mask;         /* 0 = reference (naf_algorithm.py:199 ignores done) */
    float grad_scale;              /* multiplies the gradient before clipping (1/world_size under DP) */
} rloa_naf_hyper;

typedef struct {
    float *m, *v;                  /* Adam moments, flat [n_params] in rloa_naf_param_order */
    int64_t* step;                 /* device scalar */
} rloa_adam_state;

int rloa_naf_num_params(int32_t state_size, int32_t action_size, int32_t hidden);

/* NAFAgent.learn split at the gradient all-reduce (naf_algorithm.py:180-213):
 *   _grads : target forward on next_states (train-mode BN), TD target r + gamma V', main forward
 *            on (states, actions), MSE loss, backward -> flat gradient grad [n_params], loss [1];
 *   _apply : clip_grad_norm_(1) + Adam + soft update of the target parameters (naf_algorithm.py:209-226).
 * Between the two the caller may all-reduce `grad` across ranks. */
int rloa_naf_learn_grads(rloa_naf_ws* ws, const rloa_naf_params* main_net, const rloa_naf_params* target_net,
                         const float* states, const float* actions, const float* rewards,
                         const float* next_states, const float* dones, int32_t batch,
                         const rloa_naf_hyper* hp, float* grad, float* loss, void* stream);
int rloa_naf_learn_apply(rloa_naf_ws* ws, const rloa_naf_params* main_net, const rloa_naf_params* target_net,
                         const rloa_adam_state* adam, const rloa_naf_hyper* hp, float* grad,
                         float* grad_norm, void* stream);
/* NAFAgent.learn in one call for a single rank (naf_algorithm.py:180-226): _grads followed by _apply, with the
 * split-K reduction, the gradient norm and clip + Adam + soft update fused into one kernel.  Same results
 * as the two calls, bit for bit; `grad` still receives the flat gradient. */
int rloa_naf_learn_step(rloa_naf_ws* ws, const rloa_naf_params* main_net, const rloa_naf_params* target_net,
                        const rloa_adam_state* adam, const float* states, const float* actions,
                        const float* rewards, const float* next_states, const float* dones, int32_t batch,
                        const rloa_naf_hyper* hp, float* grad, float* loss, float* grad_norm, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gradient exchange over NVLink peer memory (N > 1 ranks on one node, one process per GPU).  Replaces the
 * NCCL all-reduce between _grads and _apply: every rank owns an exchange block exported with CUDA IPC and
 * mapped by its peers; rloa_naf_learn_apply_xchg publishes the local gradient, sums all ranks' gradients
 * in rank order straight from the mapped peer blocks inside the optimiser's norm pass, then runs the same
 * clip + Adam + soft-update kernel.  No reference counterpart (the reference is single-process).
 *   rloa_xchg_create  -> rloa_xchg_handle (64-byte IPC handle, HOST) -> all-gather the handles by any
 *   means -> rloa_xchg_connect(rank, world, handles [world][64] HOST).  world == 1 needs no handles.
 * rloa_xchg_status returns 1 after a device-side wait for a peer timed out (~2 s), 0 otherwise.
 * ---------------------------------------------------------------------------------------------- */
typedef struct rloa_xchg rloa_xchg;
int rloa_xchg_create(int32_t n_floats, rloa_xchg** out);
int rloa_xchg_handle(const rloa_xchg* x, uint8_t* handle_out_host);
int rloa_xchg_connect(rloa_xchg* x, int32_t rank, int32_t world, const uint8_t* handles_host);
int rloa_xchg_status(const rloa_xchg* x);
void rloa_xchg_destroy(rloa_xchg* x);
int rloa_naf_learn_apply_xchg(rloa_naf_ws* ws, const rloa_naf_params* main_net, const rloa_naf_params* target_net,
                              const rloa_adam_state* adam, const rloa_naf_hyper* hp, rloa_xchg* xchg,
                              float* grad, float* grad_norm, void* stream);
/* rloa_naf_learn_grads + rloa_naf_learn_apply_xchg in one call: the last split-K sum of the backward pass is
 * folded into the publish kernel (one launch less on every rank's critical path); same results. */
int rloa_naf_learn_step_xchg(rloa_naf_ws* ws, const rloa_naf_params* main_net, const rloa_naf_params* target_net,
                             const rloa_adam_state* adam, rloa_xchg* xchg, const float* states, const float* actions,
                             const float* rewards, const float* next_states, const float* dones, int32_t batch,
                             const rloa_naf_hyper* hp, float* grad, float* loss, float* grad_norm, void* stream);

/* NAFAgent.soft_update alone (naf_algorithm.py:217-226) */
int rloa_naf_soft_update(const rloa_naf_params* main_net, const rloa_naf_params* target_net, float tau,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Replay ring in HBM (utils/replay_buffer.py:14-75).  Storage is caller-owned:
 * states/next_states [capacity][S], actions [capacity][A], rewards/dones [capacity].
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t capacity, state_size, action_size;
    float *states, *actions, *rewards, *next_states, *dones;
    int64_t* cursor;               /* device scalar: total transitions ever appended */
    int32_t* scratch;              /* device, ZERO-INITIALISED, >= (max n per append + 63) / 64 + 1 ints:
                                      [0] completion ticket (returns to 0), [1..] per-block valid counts */
} rloa_replay;

/* ReplayBuffer.add for n transitions (deque(maxlen) overwrite order, replay_buffer.py:32-45);
 * rows with valid == 0 are skipped (valid may be NULL). */
int rloa_replay_append(const rloa_replay* rb, int32_t n, const float* states, const float* actions,
                       const float* rewards, const float* next_states, const uint8_t* dones,
                       const uint8_t* valid, void* stream);
/* rloa_replay_append in two halves: _rows copies the (valid) rows into the slots the cursor points at without moving it,
 * _commit moves the cursor by the number of valid rows.  Between the two the ring's readers still see the old window
 * (rloa_naf_learn_step_pending reads the new rows from the caller's arrays meanwhile). */
int rloa_replay_append_rows(const rloa_replay* rb, int32_t n, const float* states, const float* actions,
                            const float* rewards, const float* next_states, const uint8_t* dones,
                            const uint8_t* valid, void* stream);
int rloa_replay_commit(const rloa_replay* rb, int32_t n, const uint8_t* valid, void* stream);
/* ReplayBuffer.sample (replay_buffer.py:47-67): `batch` distinct slots drawn uniformly from the
 * live window with a keyed Feistel permutation (seed, draw + *draw_offset); draw_offset is a DEVICE
 * counter and may be NULL; gathers the five fields. */
int rloa_replay_sample(const rloa_replay* rb, int32_t batch, uint64_t seed, uint64_t draw,
                       const uint64_t* draw_offset, float* states, float* actions, float* rewards,
                       float* next_states, float* dones, int32_t* indices, void* stream);

/* ReplayBuffer.sample + NAFAgent.learn in one call (naf_algorithm.py:152-154: `experiences = self.memory.sample();
 * self.learn(experiences)`), available when rloa_naf_learn_fused_supported(ws, batch) != 0 (trunk mode 1, batch <= 1024): the
 * fused learn kernel draws the same slots as rloa_replay_sample(seed, draw + *draw_offset) and reads its rows straight from
 * the ring, so no sampled copy of the batch is written.  xchg may be NULL (single rank). */
int rloa_naf_learn_fused_supported(const rloa_naf_ws* ws, int32_t batch);
int rloa_naf_learn_step_replay(rloa_naf_ws* ws, const rloa_naf_params* main_net, const rloa_naf_params* target_net,
                               const rloa_adam_state* adam, rloa_xchg* xchg, const rloa_replay* rb, uint64_t seed, uint64_t draw,
                               const uint64_t* draw_offset, int32_t batch, const rloa_naf_hyper* hyper, float* grad, float* loss,
                               float* grad_norm, void* stream);
/* The same with the step's transitions still PENDING (naf_algorithm.py:149-154: `self.memory.add(...)` immediately followed by
 * sample + learn): the caller copies the n_pending rows into the ring with rloa_replay_append_rows on ANOTHER stream, in
 * parallel with this call, and moves the cursor with rloa_replay_commit only after this call has completed.  The kernel
 * samples the ring as it will be after the commit - the same slots rloa_replay_append + rloa_naf_learn_step_replay draw - and
 * reads a drawn slot of the pending range from the caller's arrays (its k-th valid row) instead of the ring, so the copy is
 * off the update's critical path and the result is bit-identical.  dones / valid may be NULL (no done mask / all rows valid);
 * valid must be 16-byte aligned; n_pending <= 16384.  Layout of the pending arrays as for rloa_replay_append. */
int rloa_naf_learn_step_pending(rloa_naf_ws* ws, const rloa_naf_params* main_net, const rloa_naf_params* target_net,
                                const rloa_adam_state* adam, rloa_xchg* xchg, const rloa_replay* rb, uint64_t seed, uint64_t draw,
                                const uint64_t* draw_offset, int32_t batch, const rloa_naf_hyper* hyper, int32_t n_pending,
                                const float* states, const float* actions, const float* rewards, const float* next_states,
                                const uint8_t* dones, const uint8_t* valid, float* grad, float* loss, float* grad_norm,
                                void* stream);
/* Optional: writes the tensor-core weight images the next rloa_naf_learn_* call of this workspace needs on the workspace's
 * side stream NOW (the caller promises not to modify the parameters before that call), so that the call itself starts with
 * the images ready.  A no-op on the other paths. */
int rloa_naf_learn_prepack(rloa_naf_ws* ws, const rloa_naf_params* main_net, const rloa_naf_params* target_net, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RLOA_B200_H */
