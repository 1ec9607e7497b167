"""
ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes front end of oracle/bullet_restatement.c (CPU, fp64).

PARITY UNPINNED against PyBullet itself (no PyBullet / pybullet_data in this image; the reference's
tests mock every p.* call).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, '_build', 'liborc_bullet.so')
MAXL, MAXS, MAXV = 32, 64, 16384


class OrcModel(C.Structure):
    _fields_ = [
        ('nl', C.c_int), ('parent', C.c_int * MAXL), ('jtype', C.c_int * MAXL),
        ('E0', (C.c_double * 9) * MAXL), ('e', (C.c_double * 3) * MAXL), ('d', (C.c_double * 3) * MAXL),
        ('axis', (C.c_double * 3) * MAXL), ('mass', C.c_double * MAXL), ('inertia', (C.c_double * 3) * MAXL),
        ('damping', C.c_double * MAXL), ('lower', C.c_double * MAXL), ('upper', C.c_double * MAXL),
        ('has_limit', C.c_int * MAXL), ('base_R', C.c_double * 9), ('base_p', C.c_double * 3),
        ('lin_damp', C.c_double), ('ang_damp', C.c_double), ('gravity', C.c_double * 3), ('dt', C.c_double),
        ('iters', C.c_int), ('resid_thresh', C.c_double), ('erp', C.c_double), ('max_vel', C.c_double),
        ('limit_max_impulse', C.c_double),
        ('ns', C.c_int), ('s_link', C.c_int * MAXS), ('s_type', C.c_int * MAXS),
        ('s_R', (C.c_double * 9) * MAXS), ('s_p', (C.c_double * 3) * MAXS), ('s_dim', (C.c_double * 3) * MAXS),
        ('obstacle_radius', C.c_double), ('target_half', C.c_double * 3), ('ee_link', C.c_int),
        ('n_obs_joints', C.c_int),
        ('s_v0', C.c_int * MAXS), ('s_vn', C.c_int * MAXS), ('nv', C.c_int), ('verts', (C.c_double * 3) * MAXV),
    ]


class OrcMotors(C.Structure):
    _fields_ = [('kp', C.c_double * MAXL), ('kd', C.c_double * MAXL), ('tpos', C.c_double * MAXL),
                ('tvel', C.c_double * MAXL), ('max_imp', C.c_double * MAXL)]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, 'bullet_restatement.c')
    if force or not os.path.isfile(LIB_PATH) or (os.path.isfile(src) and
                                                 os.path.getmtime(src) > os.path.getmtime(LIB_PATH)):
        subprocess.check_call(['make', '-C', HERE], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        assert _lib.orc_sizeof_model() == C.sizeof(OrcModel), 'orc_model layout mismatch'
        assert _lib.orc_sizeof_motors() == C.sizeof(OrcMotors), 'orc_motors layout mismatch'
        _lib.orc_substep.restype = C.c_int
        _lib.orc_gjk_hull_box.restype = C.c_double
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def gjk_hull_box(verts, R, p, box_centre, box_half):
    """Distance between hull(verts) posed by (R, p) and an axis-aligned box (half = 0: a point); (distance, iterations)."""
    verts = np.ascontiguousarray(verts, np.float64).reshape(-1, 3)
    it = C.c_int(0)
    d = lib().orc_gjk_hull_box(_dp(verts), C.c_int(verts.shape[0]), _dp(np.ascontiguousarray(R, np.float64)),
                               _dp(np.ascontiguousarray(p, np.float64)),
                               _dp(np.ascontiguousarray(box_centre, np.float64)),
                               _dp(np.ascontiguousarray(box_half, np.float64)), C.byref(it))
    return float(d), it.value


class BulletOracle:
    """CPU fp64 restatement of the PyBullet step for one RobotModel (arrays only, no product code)."""

    def __init__(self, model, ee_link: int, n_obs_joints: int, obstacle_radius: float = 0.075,
                 target_half=(0.025, 0.025, 0.025)):
        m = OrcModel()
        nl = model.nl
        m.nl = nl
        for i in range(nl):
            m.parent[i] = int(model.parent[i]); m.jtype[i] = int(model.jtype[i])
            for k in range(9):
                m.E0[i][k] = float(model.E0[i][k])
            for k in range(3):
                m.e[i][k] = float(model.e[i][k]); m.d[i][k] = float(model.d[i][k])
                m.axis[i][k] = float(model.axis[i][k]); m.inertia[i][k] = float(model.inertia[i][k])
            m.mass[i] = float(model.mass[i]); m.damping[i] = float(model.damping[i])
            m.lower[i] = float(model.lower[i]); m.upper[i] = float(model.upper[i])
            m.has_limit[i] = int(model.has_limit[i])
        for k in range(9):
            m.base_R[k] = float(model.base_R[k])
        for k in range(3):
            m.base_p[k] = float(model.base_p[k]); m.gravity[k] = float(model.gravity[k])
            m.target_half[k] = float(target_half[k])
        m.lin_damp, m.ang_damp, m.dt, m.iters = model.lin_damp, model.ang_damp, model.dt, model.iters
        m.resid_thresh, m.erp, m.max_vel = model.resid_thresh, model.erp, model.max_vel
        m.limit_max_impulse = model.limit_max_impulse
        m.ns = model.ns
        for s in range(model.ns):
            m.s_link[s] = int(model.s_link[s]); m.s_type[s] = int(model.s_type[s])
            for k in range(9):
                m.s_R[s][k] = float(model.s_R[s][k])
            for k in range(3):
                m.s_p[s][k] = float(model.s_p[s][k]); m.s_dim[s][k] = float(model.s_dim[s][k])
        verts = np.asarray(getattr(model, 'verts', np.zeros((0, 3))), np.float64).reshape(-1, 3)
        assert verts.shape[0] <= MAXV, 'too many hull vertices for the oracle'
        m.nv = verts.shape[0]
        if m.nv:
            C.memmove(m.verts, np.ascontiguousarray(verts).ctypes.data, verts.nbytes)
            for s in range(model.ns):
                m.s_v0[s] = int(model.s_v0[s]); m.s_vn[s] = int(model.s_vn[s])
        m.obstacle_radius = obstacle_radius
        m.ee_link = ee_link
        m.n_obs_joints = n_obs_joints
        self.m, self.nl, self.S = m, nl, 9 + 2 * n_obs_joints
        self.movable = np.asarray(model.jtype) != 0
        self.motors = OrcMotors()
        for i in range(nl):                  # load-time default motor: velocity target 0, maxImpulse 1
            self.motors.kp[i], self.motors.kd[i], self.motors.tpos[i] = 0.0, 1.0, 0.0
            self.motors.tvel[i], self.motors.max_imp[i] = 0.0, 1.0

    # ---- motor bookkeeping mirroring setJointMotorControl2 ---------------------------------------
    def set_position_control(self, joint, target=0.0):
        mo = self.motors
        mo.kp[joint], mo.kd[joint], mo.tpos[joint], mo.tvel[joint] = 0.1, 1.0, float(target), 0.0
        mo.max_imp[joint] = 100000.0 * self.m.dt

    def set_velocity_control(self, joint, vel, force):
        mo = self.motors
        mo.kp[joint], mo.kd[joint], mo.tpos[joint], mo.tvel[joint] = 0.0, 1.0, 0.0, float(vel)
        mo.max_imp[joint] = float(force) * self.m.dt

    # ---- single-env primitives ------------------------------------------------------------------
    def fk(self, q):
        q = np.ascontiguousarray(q, np.float64)
        Rw, pw = np.zeros((self.nl, 9)), np.zeros((self.nl, 3))
        lib().orc_fk(C.byref(self.m), _dp(q), _dp(Rw), _dp(pw))
        return Rw.reshape(self.nl, 3, 3), pw

    def aba(self, q, qd, tau=None):
        q, qd = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(qd, np.float64)
        out = np.zeros(self.nl)
        tp = _dp(np.ascontiguousarray(tau, np.float64)) if tau is not None else None
        lib().orc_aba(C.byref(self.m), _dp(q), _dp(qd), tp, _dp(out))
        return out

    def minv(self, q):
        q = np.ascontiguousarray(q, np.float64)
        out = np.zeros((self.nl, self.nl))
        lib().orc_minv(C.byref(self.m), _dp(q), _dp(out))
        return out

    def crba(self, q):
        q = np.ascontiguousarray(q, np.float64)
        out = np.zeros((self.nl, self.nl))
        lib().orc_crba(C.byref(self.m), _dp(q), _dp(out))
        return out

    def bias(self, q, qd):
        q, qd = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(qd, np.float64)
        out = np.zeros(self.nl)
        lib().orc_rnea_bias(C.byref(self.m), _dp(q), _dp(qd), _dp(out))
        return out

    def substep(self, q, qd):
        """In-place one stepSimulation with the current motor table; returns PGS iterations used."""
        return lib().orc_substep(C.byref(self.m), C.byref(self.motors), _dp(q), _dp(qd))

    def distances(self, q, obstacle, target):
        q = np.ascontiguousarray(q, np.float64)
        lo, ee, eep = np.zeros(self.nl), np.zeros(1), np.zeros(3)
        lib().orc_distances(C.byref(self.m), _dp(q), _dp(np.ascontiguousarray(obstacle, np.float64)),
                            _dp(np.ascontiguousarray(target, np.float64)), _dp(lo), _dp(ee), _dp(eep))
        return lo, float(ee[0]), eep

    def self_distances(self, q):
        """[nl][nl] link-link closest distances (environment.py:394-412); 10 where the reference does not query."""
        out = np.zeros((self.nl, self.nl))
        lib().orc_self_distances(C.byref(self.m), _dp(np.ascontiguousarray(q, np.float64)), _dp(out))
        return out

    def observe(self, q, qd, obstacle, target):
        obs, rew, done = np.zeros(self.S), np.zeros(1), np.zeros(1, np.int32)
        lib().orc_observe(C.byref(self.m), _dp(np.ascontiguousarray(q, np.float64)),
                          _dp(np.ascontiguousarray(qd, np.float64)),
                          _dp(np.ascontiguousarray(obstacle, np.float64)),
                          _dp(np.ascontiguousarray(target, np.float64)), _dp(obs), _dp(rew), _ip(done))
        return obs, float(rew[0]), int(done[0])

    # ---- batched (Environment.step / reset over n envs) -----------------------------------------
    def batch_step(self, q, qd, actions, act_joints, max_force, obstacle, target, nthreads=1, contact_threshold=0.0):
        """contact_threshold > 0 (Bullet: 0.02) adds the normal contact rows against obstacle and target; the number of
        contact rows per env is left in self.last_contacts."""
        n = q.shape[0]
        assert q.dtype == np.float64 and qd.dtype == np.float64 and q.flags.c_contiguous
        actions = np.ascontiguousarray(actions, np.float64)
        aj = np.ascontiguousarray(act_joints, np.int32)
        obstacle = np.ascontiguousarray(np.broadcast_to(obstacle, (n, 3)), np.float64)
        target = np.ascontiguousarray(np.broadcast_to(target, (n, 3)), np.float64)
        obs, rew = np.zeros((n, self.S)), np.zeros(n)
        done, iters = np.zeros(n, np.int32), np.zeros(n, np.int32)
        if contact_threshold > 0.0:
            nc = np.zeros(n, np.int32)
            lib().orc_batch_step_contacts(C.byref(self.m), C.byref(self.motors), _ip(aj), C.c_int(len(aj)), C.c_int(n),
                                          _dp(q), _dp(qd), _dp(actions), C.c_double(max_force), _dp(obstacle), _dp(target),
                                          C.c_double(contact_threshold), _dp(obs), _dp(rew), _ip(done), _ip(iters), _ip(nc),
                                          C.c_int(nthreads))
            self.last_contacts = nc
            return obs, rew, done, iters
        lib().orc_batch_step(C.byref(self.m), C.byref(self.motors), _ip(aj), C.c_int(len(aj)), C.c_int(n),
                             _dp(q), _dp(qd), _dp(actions), C.c_double(max_force), _dp(obstacle), _dp(target),
                             _dp(obs), _dp(rew), _ip(done), _ip(iters), C.c_int(nthreads))
        return obs, rew, done, iters

    def batch_reset(self, q, qd, init_targets, nsub=50, nthreads=1, obstacle=None, target=None, contact_threshold=0.0):
        n = q.shape[0]
        init_targets = np.ascontiguousarray(init_targets, np.float64)
        if contact_threshold > 0.0:
            obstacle = np.ascontiguousarray(np.broadcast_to(obstacle, (n, 3)), np.float64)
            target = np.ascontiguousarray(np.broadcast_to(target, (n, 3)), np.float64)
            lib().orc_batch_reset_contacts(C.byref(self.m), C.byref(self.motors), C.c_int(init_targets.shape[1]), C.c_int(n),
                                           _dp(q), _dp(qd), _dp(init_targets), C.c_int(nsub), _dp(obstacle), _dp(target),
                                           C.c_double(contact_threshold), C.c_int(nthreads))
            return
        lib().orc_batch_reset(C.byref(self.m), C.byref(self.motors), C.c_int(init_targets.shape[1]), C.c_int(n),
                              _dp(q), _dp(qd), _dp(init_targets), C.c_int(nsub), C.c_int(nthreads))
