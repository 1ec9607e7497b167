"""Shared test helpers: model loading, oracle construction, seeded state generation."""
import numpy as np

from oracle.bullet_oracle import BulletOracle
from robotic_manipulator_rloa_b200.environment.robot_model import load_manipulator

KUKA = dict(file='kuka_iiwa/kuka_with_gripper2.sdf', ee=13, involved=[0, 1, 2, 3, 4, 5],
            fixed=[6, 7, 8, 9, 10, 11, 12, 13], target=[0.4, 0.85, 0.71], obstacle=[0.45, 0.55, 0.55],
            start=[0.9, 0.45, 0, 0, 0, 0])
PANDA = dict(file='franka_panda/panda.urdf', ee=11, involved=[0, 1, 2, 3, 4, 5, 6], fixed=[7, 8, 9, 10, 11],
             target=[0.4, 0.3, 0.5], obstacle=[0.3, 0.0, 0.6], start=[0, 0, 0, -1.5, 0, 1.5, 0])


# the reference's xarm6 demo (rl_framework.py:571-580): involved joints 1-6, joint 0 is the fixed world joint
XARM6 = dict(file='xarm/xarm6_with_gripper.urdf', ee=12, involved=[1, 2, 3, 4, 5, 6], fixed=[0, 7, 8, 9, 10, 11, 12, 13],
             target=[0.3, 0.47, 0.61], obstacle=[0.25, 0.27, 0.5], start=[0., 1., 0., -2.3, 0., 0., 0.])


def parity_report(case, dq_e, dqd_e, tol_q=1e-4, tol_qd=1e-4, **extra):
    """Counts against the FLAT north_star bounds (1e-4 rad, 1e-4 rad/s): how many envs exceed them and the worst case.
    Printed (so the driver's log shows them) and appended to gpurun_out/parity_counts.jsonl when that directory exists;
    the callers assert the counts against a stated budget instead of widening the bound."""
    import json
    import os
    dq_e, dqd_e = np.asarray(dq_e, float), np.asarray(dqd_e, float)
    rep = dict(case=case, n=int(dq_e.size), n_over_q=int((dq_e > tol_q).sum()), n_over_qd=int((dqd_e > tol_qd).sum()),
               worst_q=float(dq_e.max()), worst_qd=float(dqd_e.max()), p99_qd=float(np.quantile(dqd_e, 0.99)),
               median_qd=float(np.median(dqd_e)))
    rep.update({k: (float(v) if isinstance(v, (float, np.floating)) else int(v)) for k, v in extra.items()})
    print('PARITY', json.dumps(rep))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    if os.path.isdir(out):
        try:
            with open(os.path.join(out, 'parity_counts.jsonl'), 'a') as f:
                f.write(json.dumps(rep) + '\n')
        except OSError:
            pass
    return rep


def make_oracle(cfg):
    model = load_manipulator(cfg['file'])
    return model, BulletOracle(model, cfg['ee'], len(cfg['involved']))


def step_motors(oracle, cfg):
    """Motor table after Environment.step's setJointMotorControl2 calls (fixed joints; involved set per env)."""
    for j in cfg['fixed']:
        oracle.set_position_control(j, 0.0)


def random_states(model, n, seed, vel=2.0, frac_limit=0.95, near_limit=0.0, held=None, held_range=0.02):
    rng = np.random.default_rng(seed)
    nl = model.nl
    q = np.zeros((n, nl))
    qd = np.zeros((n, nl))
    for i in range(nl):
        if model.jtype[i] == 0:
            continue
        if model.has_limit[i]:
            lo, hi = model.lower[i], model.upper[i]
        else:
            lo, hi = -np.pi, np.pi
        mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo)
        q[:, i] = mid + half * frac_limit * rng.uniform(-1, 1, n)
        if near_limit > 0 and model.has_limit[i]:
            pick = rng.uniform(size=n) < near_limit
            side = rng.uniform(size=n) < 0.5
            over = rng.uniform(-2e-3, 2e-3, n)
            q[pick, i] = np.where(side[pick], lo + over[pick], hi + over[pick])
        qd[:, i] = rng.uniform(-vel, vel, n)
        if held is not None and i in held:     # joints the reference holds at 0 with POSITION_CONTROL
            q[:, i] = rng.uniform(-held_range, held_range, n)
            qd[:, i] = rng.uniform(-0.2, 0.2, n)
    return q, qd


def assert_params_close(got, want, name, rtol=2e-4, atol=2e-6, lr=1e-3, max_frac=1e-3):
    """Post-Adam parameters: Adam moves every element by ~lr * g / (|g| + eps) — for elements whose gradient is at
    rounding-noise level the direction is arbitrary, so a small fraction of elements may differ by up to a couple of
    lr-sized steps; everything else must agree tightly."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    if got.dtype.kind != 'f' or got.ndim == 0:
        assert np.array_equal(got, want), name
        return
    bad = np.abs(got - want) > atol + rtol * np.abs(want)
    assert bad.mean() <= max_frac, f'{name}: {bad.sum()} of {bad.size} elements differ'
    assert np.abs(got - want).max() <= 3.1 * lr, f'{name}: max diff {np.abs(got - want).max():.3e}'


# ---- mesh (convex hull) variants of the shipped models -----------------------------------------------------------
def _fibonacci_sphere(n):
    k = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * k / n)
    th = np.pi * (1 + 5 ** 0.5) * k
    return np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)


def hull_cloud(kind, dim, n=40):
    """Vertex cloud standing in for a primitive: box -> its 8 corners (the same shape), sphere / capsule -> n points
    on the surface (a polytope inscribed in the primitive)."""
    dim = np.asarray(dim, float)
    if kind == 3:
        return np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float) * dim
    pts = _fibonacci_sphere(n) * dim[0]
    if kind == 2:
        pts[:, 2] += np.where(pts[:, 2] >= 0, dim[1], -dim[1])
    return pts


def hullified(model, margin=0.001, n=40, only_links=None):
    """Copy of a RobotModel whose collision primitives became SHAPE_HULL vertex clouds (what a file with <mesh>
    collision elements loads as)."""
    import dataclasses
    s_type, s_dim = model.s_type.copy(), model.s_dim.copy()
    v0, vn, verts, at = [], [], [], 0
    for s in range(model.ns):
        if only_links is not None and int(model.s_link[s]) not in only_links:
            v0.append(at); vn.append(0)
            continue
        cloud = hull_cloud(int(model.s_type[s]), model.s_dim[s], n)
        s_type[s], s_dim[s] = 4, (margin, 0.0, 0.0)
        v0.append(at); vn.append(len(cloud)); verts.append(cloud); at += len(cloud)
    return dataclasses.replace(model, s_type=s_type, s_dim=s_dim, s_v0=np.asarray(v0, np.int32),
                               s_vn=np.asarray(vn, np.int32), verts=np.concatenate(verts, axis=0))


BOX_CORNERS = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float)
BOX_TRIS = [(0, 1, 3), (0, 3, 2), (4, 6, 7), (4, 7, 5), (0, 4, 5), (0, 5, 1), (2, 3, 7), (2, 7, 6), (0, 2, 6), (0, 6, 4),
            (1, 5, 7), (1, 7, 3)]


def write_box_mesh(path, half, fmt):
    """A box as binary STL / ASCII STL / OBJ (12 triangles)."""
    import struct
    v = BOX_CORNERS * np.asarray(half, float)
    if fmt == 'obj':
        with open(path, 'w') as f:
            f.write('# box\n')
            for p in v:
                f.write('v %.9g %.9g %.9g\n' % tuple(p))
            for t in BOX_TRIS:
                f.write('f %d %d %d\n' % (t[0] + 1, t[1] + 1, t[2] + 1))
    elif fmt == 'stl_ascii':
        with open(path, 'w') as f:
            f.write('solid box\n')
            for t in BOX_TRIS:
                f.write(' facet normal 0 0 0\n  outer loop\n')
                for i in t:
                    f.write('   vertex %.9g %.9g %.9g\n' % tuple(v[i]))
                f.write('  endloop\n endfacet\n')
            f.write('endsolid box\n')
    else:
        with open(path, 'wb') as f:
            f.write(b'solid binary box'.ljust(80, b' '))       # starts with "solid" like many exporters' files
            f.write(struct.pack('<I', len(BOX_TRIS)))
            for t in BOX_TRIS:
                f.write(struct.pack('<3f', 0, 0, 0))
                for i in t:
                    f.write(struct.pack('<3f', *v[i]))
                f.write(struct.pack('<H', 0))


def write_test_arm(dirname, geometry):
    """A 3-joint arm as URDF; geometry(link index) -> the inner XML of each link's <collision><geometry>."""
    links, joints = [], []
    for i in range(4):
        col = '' if i == 0 else f'<collision><origin xyz="0 0 0.1" rpy="0.1 0.2 0.3"/><geometry>{geometry(i)}</geometry></collision>'
        links.append(f'<link name="l{i}"><inertial><origin xyz="0 0 0.1"/><mass value="1.5"/>'
                     f'<inertia ixx="0.01" iyy="0.012" izz="0.005" ixy="0" ixz="0" iyz="0"/></inertial>{col}</link>')
    axes = ['0 0 1', '0 1 0', '1 0 0']
    for i in range(3):
        joints.append(f'<joint name="j{i}" type="revolute"><parent link="l{i}"/><child link="l{i + 1}"/>'
                      f'<origin xyz="0 0 0.2" rpy="0 0 0"/><axis xyz="{axes[i]}"/>'
                      f'<limit lower="-2.5" upper="2.5" effort="100" velocity="10"/></joint>')
    path = f'{dirname}/arm.urdf'
    with open(path, 'w') as f:
        f.write('<?xml version="1.0"?><robot name="arm">' + ''.join(links) + ''.join(joints) + '</robot>')
    return path


# ---- RobotModel <-> flat npz entries (tests/golden/make_bullet_golden.py stores the model PyBullet loaded) -------
_MODEL_ARRAYS = ('parent', 'jtype', 'E0', 'e', 'd', 'axis', 'mass', 'inertia', 'damping', 'lower', 'upper', 'has_limit',
                 'base_R', 'base_p', 's_link', 's_type', 's_R', 's_p', 's_dim', 'joint_axis_link', 's_v0', 's_vn', 'verts')
_MODEL_SCALARS = ('lin_damp', 'ang_damp', 'dt', 'iters', 'resid_thresh', 'erp', 'max_vel', 'limit_max_impulse')


def model_to_entries(model, prefix='model_'):
    out = {prefix + k: np.asarray(getattr(model, k)) for k in _MODEL_ARRAYS}
    out.update({prefix + k: np.asarray(getattr(model, k)) for k in _MODEL_SCALARS})
    out[prefix + 'gravity'] = np.asarray(model.gravity, float)
    out[prefix + 'joint_names'] = np.asarray(model.joint_names)
    out[prefix + 'link_names'] = np.asarray(model.link_names)
    return out


def model_from_entries(z, prefix='model_'):
    from robotic_manipulator_rloa_b200.environment.robot_model import RobotModel
    kw = {k: np.asarray(z[prefix + k]) for k in _MODEL_ARRAYS}
    for k in ('parent', 'jtype', 'has_limit', 's_link', 's_type', 's_v0', 's_vn'):
        kw[k] = kw[k].astype(np.int32)
    kw.update({k: (int if k == 'iters' else float)(z[prefix + k]) for k in _MODEL_SCALARS})
    return RobotModel(nl=int(kw['parent'].shape[0]), gravity=tuple(float(v) for v in z[prefix + 'gravity']),
                      joint_names=[str(s) for s in z[prefix + 'joint_names']],
                      link_names=[str(s) for s in z[prefix + 'link_names']], **kw)
