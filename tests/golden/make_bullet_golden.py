"""
Records golden vectors of the reference's PyBullet path — the piece this repository cannot produce itself: the image it
is built in has neither the `pybullet` wheel nor `pybullet_data` ("parity unpinned").  Run it on any
machine that has them, next to a checkout of the reference and of this repository:

    python -m pip install pybullet==3.2.6 numpy torch matplotlib     # any 3.2.x; the reference leaves it unpinned (pyproject.toml:24)
    export PYTHONPATH=/path/to/robotic_manipulator_rloa:$PYTHONPATH  # the UNMODIFIED reference checkout (or pip install it)
    python tests/golden/make_bullet_golden.py [all|kuka|xarm6|panda] [n_records=256]      # ~10 s per robot on one core
    git add tests/golden/bullet_*.npz                                # ~150 KB each; the tests below stop skipping


It drives the UNMODIFIED reference `Environment` (PyBullet DIRECT): for every record the joints are put into a seeded
state with `p.resetJointState`, one `Environment.step(action)` runs, and joint states, observation, reward, done, the
per-link obstacle distances and the end-effector / target distance are stored — together with the model THIS repository's
loader builds from the very same asset file, so the checks need no asset at test time.  It also records what Bullet itself
says about the model and the world (getDynamicsInfo, getJointInfo, getCollisionShapeData, getPhysicsEngineParameters):
masses, local inertia diagonals, inertial frames, joint damping, linear / angular damping, collision margins, time step,
solver iterations, ERP — the constants SURVEY.md Appendices A and C recall from memory, checked by
tests/test_bullet_golden.py::test_model_constants_against_recorded_pybullet.  The output,
tests/golden/bullet_<robot>.npz, is picked up by tests/test_bullet_golden.py (oracle on CPU, CUDA simulator with -m gpu),
which skip while the file is absent.  NOT runnable in the build image; written against the PyBullet API the reference
itself uses (environment.py:207-255, 431-485; utils/collision_detector.py:33-61).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))                       # tests/ (helpers)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))      # repository root

import pybullet as p                                            # noqa: E402
import pybullet_data                                            # noqa: E402
from robotic_manipulator_rloa.environment.environment import Environment, EnvironmentConfiguration   # noqa: E402
from robotic_manipulator_rloa.utils.collision_detector import CollisionDetector, CollisionObject      # noqa: E402

from helpers import KUKA, PANDA, XARM6, model_to_entries        # noqa: E402
from robotic_manipulator_rloa_b200.environment.robot_model import load_manipulator                    # noqa: E402

def record(robot: str, n_rec: int) -> None:
    cfg = {'kuka': KUKA, 'xarm6': XARM6, 'panda': PANDA}[robot]
    path = os.path.join(pybullet_data.getDataPath(), cfg['file'])

    env = Environment(path, EnvironmentConfiguration(
        endeffector_index=cfg['ee'], fixed_joints=cfg['fixed'], involved_joints=cfg['involved'],
        target_position=cfg['target'], obstacle_position=cfg['obstacle'], initial_joint_positions=cfg['start'],
        initial_positions_variation_range=[0.0] * len(cfg['start']), max_force=200., visualize=False))
    uid, nl = env.manipulator_uid, env.num_joints
    model = load_manipulator(path)                                  # this repository's front end on the same file
    assert model.nl == nl, f'joint count: loader {model.nl}, PyBullet {nl}'

    rng = np.random.default_rng(0)
    lo = np.array([p.getJointInfo(uid, j)[8] for j in range(nl)])
    hi = np.array([p.getJointInfo(uid, j)[9] for j in range(nl)])
    movable = np.array([p.getJointInfo(uid, j)[2] != p.JOINT_FIXED for j in range(nl)])
    rec = {k: [] for k in ('q0', 'qd0', 'action', 'q1', 'qd1', 'obs', 'reward', 'done', 'link_obstacle', 'ee_target', 'com')}
    env.step(np.zeros(len(cfg['involved'])))                        # the motors every later step re-arms are in place
    for k in range(n_rec):
        q0, qd0 = np.zeros(nl), np.zeros(nl)
        for j in range(nl):
            if not movable[j]:
                continue
            a, b = (lo[j], hi[j]) if lo[j] < hi[j] else (-np.pi, np.pi)
            if j in cfg['fixed']:                                   # the reference holds these at 0
                q0[j], qd0[j] = rng.uniform(-0.02, 0.02), rng.uniform(-0.2, 0.2)
            else:
                q0[j], qd0[j] = 0.5 * (a + b) + 0.45 * (b - a) * rng.uniform(-1, 1), rng.uniform(-1, 1)
            p.resetJointState(uid, j, q0[j], qd0[j])
        action = rng.uniform(-1, 1, len(cfg['involved']))
        obs, reward, done = env.step(action)
        js = [p.getJointState(uid, j) for j in range(nl)]
        rec['q0'].append(q0); rec['qd0'].append(qd0); rec['action'].append(action)
        rec['q1'].append([s[0] for s in js]); rec['qd1'].append([s[1] for s in js])
        rec['obs'].append(np.asarray(obs, float)); rec['reward'].append(float(reward)); rec['done'].append(int(done))
        rec['link_obstacle'].append([CollisionDetector(CollisionObject(uid, l), [env.obstacle]).compute_distances()[0]
                                     for l in range(nl)])
        rec['ee_target'].append(CollisionDetector(CollisionObject(uid, cfg['ee']), [env.target]).compute_distances()[0])
        rec['com'].append([p.getLinkState(uid, l)[0] for l in range(nl)])

    # ---- 400-step trajectories (north_star: "400-step trajectory divergence reported") -----------------------------
    n_traj, T = 8, 400
    traj_q0 = np.zeros((n_traj, nl)); traj_qd0 = np.zeros((n_traj, nl))
    traj_actions = rng.uniform(-1, 1, (n_traj, T, len(cfg['involved'])))
    traj_q = np.zeros((n_traj, T, nl)); traj_qd = np.zeros((n_traj, T, nl))
    traj_done = np.zeros((n_traj, T), int)
    for k in range(n_traj):
        for j in range(nl):
            if movable[j] and j not in cfg['fixed']:
                a, b = (lo[j], hi[j]) if lo[j] < hi[j] else (-np.pi, np.pi)
                traj_q0[k, j] = 0.5 * (a + b) + 0.3 * (b - a) * rng.uniform(-1, 1)
            p.resetJointState(uid, j, traj_q0[k, j], 0.0)
        for t in range(T):
            _, _, done = env.step(traj_actions[k, t])          # the recording does not stop at done: physics only
            js = [p.getJointState(uid, j) for j in range(nl)]
            traj_q[k, t] = [s[0] for s in js]; traj_qd[k, t] = [s[1] for s in js]
            traj_done[k, t] = int(done)

    out = {k: np.asarray(v) for k, v in rec.items()}
    out.update(traj_q0=traj_q0, traj_qd0=traj_qd0, traj_actions=traj_actions, traj_q=traj_q, traj_qd=traj_qd,
               traj_done=traj_done)
    out.update(model_to_entries(model))
    # ---- what Bullet says about the model and the world (SURVEY.md Appendix A / C constants) ----
    dyn = [p.getDynamicsInfo(uid, l) for l in range(nl)]
    out['bt_mass'] = np.asarray([d[0] for d in dyn])
    out['bt_lateral_friction'] = np.asarray([d[1] for d in dyn])
    out['bt_local_inertia_diag'] = np.asarray([d[2] for d in dyn])
    out['bt_local_inertial_pos'] = np.asarray([d[3] for d in dyn])
    out['bt_local_inertial_orn'] = np.asarray([d[4] for d in dyn])
    out['bt_contact_damping'] = np.asarray([d[8] for d in dyn])
    out['bt_contact_stiffness'] = np.asarray([d[9] for d in dyn])
    out['bt_collision_margin'] = np.asarray([d[11] if len(d) > 11 else np.nan for d in dyn])
    ji = [p.getJointInfo(uid, j) for j in range(nl)]
    out['bt_joint_type'] = np.asarray([j[2] for j in ji])
    out['bt_joint_damping'] = np.asarray([j[6] for j in ji])
    out['bt_joint_friction'] = np.asarray([j[7] for j in ji])
    out['bt_joint_lower'] = np.asarray([j[8] for j in ji]); out['bt_joint_upper'] = np.asarray([j[9] for j in ji])
    out['bt_joint_max_force'] = np.asarray([j[10] for j in ji]); out['bt_joint_max_velocity'] = np.asarray([j[11] for j in ji])
    out['bt_joint_axis'] = np.asarray([j[13] for j in ji])
    out['bt_parent_frame_pos'] = np.asarray([j[14] for j in ji]); out['bt_parent_frame_orn'] = np.asarray([j[15] for j in ji])
    out['bt_parent_index'] = np.asarray([j[16] for j in ji])
    shapes = []
    for l in range(-1, nl):
        for sh in p.getCollisionShapeData(uid, l):          # (uid, link, geom type, dimensions, file, local pos, local orn)
            shapes.append([l, sh[2], *sh[3], *sh[5], *sh[6]])
    out['bt_collision_shapes'] = np.asarray(shapes, float).reshape(-1, 12) if shapes else np.zeros((0, 12))
    for body, tag in ((env.obstacle, 'obstacle'), (env.target, 'target')):
        sh = p.getCollisionShapeData(body, -1)[0]
        out[f'bt_{tag}_shape'] = np.asarray([sh[2], *sh[3]], float)
        out[f'bt_{tag}_margin'] = np.asarray(p.getDynamicsInfo(body, -1)[11] if len(p.getDynamicsInfo(body, -1)) > 11 else np.nan)
    pe = p.getPhysicsEngineParameters()
    for k in ('fixedTimeStep', 'numSubSteps', 'numSolverIterations', 'erp', 'contactERP', 'frictionERP', 'solverResidualThreshold',
              'contactBreakingThreshold', 'gravityAccelerationZ'):
        if k in pe:
            out['bt_world_' + k] = np.asarray(pe[k])
    out['pybullet_api_version'] = np.asarray(p.getAPIVersion())
    out['robot'] = np.asarray(robot)
    dst = os.path.join(HERE, f'bullet_{robot}.npz')
    np.savez_compressed(dst, **out)
    p.disconnect()
    print(f'wrote {dst}: {n_rec} records, {nl} joints, {int(np.sum(out["done"]))} terminal steps')


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    for name in (['kuka', 'xarm6', 'panda'] if which == 'all' else [which]):
        record(name, count)
