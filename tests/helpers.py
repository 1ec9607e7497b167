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
