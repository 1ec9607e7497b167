"""
GPU parity of the convex-hull (mesh) collision path: GJK in the CUDA simulator, through the C ABI, against the fp64
restatement in oracle/ on the same seeded states.  Hull vertices come from (a) the shipped models with every
primitive replaced by a vertex cloud and (b) URDF files with STL / OBJ collision meshes written by the test.
Tolerance: 2e-5 m on distances (fp32 GJK in world coordinates of order 1 m), flags exact away from boundaries.
"""
import numpy as np
import pytest
import torch

from helpers import KUKA, PANDA, hullified, make_oracle, random_states, step_motors, write_box_mesh, write_test_arm
from oracle.bullet_oracle import BulletOracle
from robotic_manipulator_rloa_b200.environment import mesh_io
from robotic_manipulator_rloa_b200.environment.robot_model import load_urdf

pytestmark = pytest.mark.gpu


def make_sim(model, cfg, n):
    from robotic_manipulator_rloa_b200.environment.simulator import BatchedSimulator
    sim = BatchedSimulator(model, n, cfg['ee'], cfg['involved'], cfg['fixed'], max_force=200.0, contacts=False)
    sim.set_task(cfg['target'], cfg['obstacle'])
    return sim


def spread_tasks(cfg, n, seed, orc=None, q=None):
    """Per-env obstacle / target positions; with an oracle and states, half of the envs get the obstacle next to a
    random link and the target next to the end-effector link, so that contact is well represented."""
    rng = np.random.default_rng(seed)
    ob = np.asarray(cfg['obstacle']) + rng.uniform(-0.25, 0.25, (n, 3))
    tg = np.asarray(cfg['target']) + rng.uniform(-0.25, 0.25, (n, 3))
    if orc is not None:
        for e in range(0, n, 2):
            _, pw = orc.fk(q[e])
            ob[e] = pw[rng.integers(1, orc.nl)] + rng.uniform(-0.12, 0.12, 3)
            tg[e] = pw[cfg['ee']] + rng.uniform(-0.08, 0.08, 3)
    return ob.astype(np.float32).astype(np.float64), tg.astype(np.float32).astype(np.float64)


@pytest.mark.parametrize('cfg', [KUKA, PANDA], ids=['kuka', 'panda'])
def test_hull_distances_match_oracle(cfg):
    model, _ = make_oracle(cfg)
    hm = hullified(model)
    orc = BulletOracle(hm, cfg['ee'], len(cfg['involved']))
    n = 1024
    q, qd = random_states(model, n, seed=21)
    ob, tg = spread_tasks(cfg, n, 22, orc, q)
    sim = make_sim(hm, cfg, n)
    sim.set_task(torch.as_tensor(tg, dtype=torch.float32), torch.as_tensor(ob, dtype=torch.float32))
    sim.set_state(q, qd)
    _, link, ee = sim.observe(want_distances=True)
    link, ee = link.cpu().numpy(), ee.cpu().numpy()
    q32 = q.astype(np.float32).astype(np.float64)
    worst_l = worst_e = 0.0
    hits = 0
    for e in range(n):
        lo, eet, _ = orc.distances(q32[e], ob[e], tg[e])
        worst_l = max(worst_l, np.abs(link[e] - lo).max())
        worst_e = max(worst_e, abs(ee[e] - eet))
        hits += int(lo.min() < 0)
    print(f'hull distances: max link err {worst_l:.2e}, max ee err {worst_e:.2e}, envs in collision {hits}/{n}')
    assert hits > n // 50                       # the sample does exercise contact
    assert worst_l <= 2e-5 and worst_e <= 2e-5


def test_box_end_effector_against_cube():
    """A box shape on the end-effector link (KUKA stand-in link 7) takes the box-vs-cube GJK path."""
    cfg = dict(KUKA, ee=7)
    model, orc = make_oracle(cfg)
    n = 1024
    q, qd = random_states(model, n, seed=31)
    ob, tg = spread_tasks(cfg, n, 32, orc, q)
    sim = make_sim(model, cfg, n)
    sim.set_task(torch.as_tensor(tg, dtype=torch.float32), torch.as_tensor(ob, dtype=torch.float32))
    sim.set_state(q, qd)
    _, link, ee = sim.observe(want_distances=True)
    q32 = q.astype(np.float32).astype(np.float64)
    want = np.array([orc.distances(q32[e], ob[e], tg[e])[1] for e in range(n)])
    assert (want < 9.0).all() and (want < 0.05).sum() > 0
    assert np.abs(ee.cpu().numpy() - want).max() <= 2e-5


def test_hull_step_flags_and_rewards_match_oracle():
    """Environment.step on a meshed model: the bounding-sphere broad phase and the GJK narrow phase inside the solve
    kernel give the oracle's done / reward away from the contact boundaries."""
    cfg = KUKA
    model, _ = make_oracle(cfg)
    hm = hullified(model)
    orc = BulletOracle(hm, cfg['ee'], len(cfg['involved']))
    n = 4096
    q, qd = random_states(model, n, seed=41, held=cfg['fixed'])
    ob, tg = spread_tasks(cfg, n, 42, orc, q)
    rng = np.random.default_rng(43)
    actions = rng.uniform(-1, 1, (n, len(cfg['involved']))).astype(np.float32)
    sim = make_sim(hm, cfg, n)
    sim.set_task(torch.as_tensor(tg, dtype=torch.float32), torch.as_tensor(ob, dtype=torch.float32))
    sim.set_state(q, qd)
    obs, rew, done = sim.step(torch.as_tensor(actions, device='cuda'))
    step_motors(orc, cfg)
    q32, qd32 = q.astype(np.float32).astype(np.float64), qd.astype(np.float32).astype(np.float64)
    obs_o, rew_o, done_o, _ = orc.batch_step(q32, qd32, actions.astype(np.float64), cfg['involved'], 200.0, ob, tg, nthreads=8)
    lo = np.array([orc.distances(q32[e], ob[e], tg[e])[0].min() for e in range(n)])
    eet = np.array([orc.distances(q32[e], ob[e], tg[e])[1] for e in range(n)])
    safe = (np.abs(lo) > 1e-4) & (np.abs(eet - 0.05) > 1e-4)
    dg, rg = done.cpu().numpy(), rew.cpu().numpy()
    print(f'meshed step: {int(done_o.sum())} terminal of {n}, {int((rew_o == 250).sum())} at the target, {int((~safe).sum())} near a boundary')
    assert done_o.sum() > n // 50 and (rew_o == 250).sum() > 0
    assert (dg[safe] == done_o[safe]).all()
    assert np.abs(rg[safe] - rew_o[safe]).max() <= 2e-4


def test_urdf_with_mesh_files_end_to_end(tmp_path):
    """box meshes (binary STL, ASCII STL, OBJ) vs the same arm with <box> primitives, both on the GPU: link distances
    differ by the hull margin; the mesh arm also matches the oracle."""
    half = {1: (0.04, 0.05, 0.1), 2: (0.03, 0.03, 0.09), 3: (0.02, 0.06, 0.05)}
    (tmp_path / 'prim').mkdir()
    (tmp_path / 'mesh' / 'meshes').mkdir(parents=True)
    fmts = {1: ('stl_binary', 'stl'), 2: ('stl_ascii', 'stl'), 3: ('obj', 'obj')}
    for i, (fmt, ext) in fmts.items():
        write_box_mesh(str(tmp_path / 'mesh' / 'meshes' / f'l{i}.{ext}'), half[i], fmt)
    prim = load_urdf(write_test_arm(str(tmp_path / 'prim'), lambda i: '<box size="%g %g %g"/>' % tuple(2 * np.asarray(half[i]))))
    mesh = load_urdf(write_test_arm(str(tmp_path / 'mesh'), lambda i: '<mesh filename="meshes/l%d.%s"/>' % (i, fmts[i][1])))
    cfg = dict(ee=2, involved=[0, 1, 2], fixed=[], target=[0.2, 0.1, 0.5], obstacle=[-0.1, 0.2, 0.4])
    n = 512
    rng = np.random.default_rng(51)
    q = rng.uniform(-2.4, 2.4, (n, 3))
    ob, tg = spread_tasks(cfg, n, 52, BulletOracle(mesh, cfg['ee'], 3), q)
    out = []
    for m in (prim, mesh):
        sim = make_sim(m, cfg, n)
        sim.set_task(torch.as_tensor(tg, dtype=torch.float32), torch.as_tensor(ob, dtype=torch.float32))
        sim.set_state(q, np.zeros_like(q))
        _, link, ee = sim.observe(want_distances=True)
        out.append((link.cpu().numpy(), ee.cpu().numpy()))
    (lp, ep), (lm, em) = out
    sep = lp > -0.075 + 1e-4                   # sphere centre outside the box: both paths report the true distance
    assert np.abs((lm + mesh_io.HULL_MARGIN - lp)[sep]).max() <= 2e-5
    assert (lm[~sep] < 0).all()
    far = ep > 1e-4
    assert np.abs((em + mesh_io.HULL_MARGIN - ep)[far]).max() <= 2e-5
    orc = BulletOracle(mesh, cfg['ee'], 3)
    q32 = q.astype(np.float32).astype(np.float64)
    for e in range(0, n, 4):
        lo, eet, _ = orc.distances(q32[e], ob[e], tg[e])
        assert np.abs(lm[e] - lo).max() <= 2e-5 and abs(em[e] - eet) <= 2e-5


# ---- link-link distances (get_manipulator_collisions_with_itself) ------------------------------------------------
@pytest.mark.parametrize('meshed', [False, True], ids=['primitives', 'hulls'])
def test_self_distances_match_oracle(meshed):
    cfg = KUKA
    model, orc = make_oracle(cfg)
    if meshed:
        model = hullified(model, n=24)
        orc = BulletOracle(model, cfg['ee'], len(cfg['involved']))
    n = 256
    q, qd = random_states(model, n, seed=61)
    sim = make_sim(model, cfg, n)
    sim.set_state(q, qd)
    got = sim.self_distances().cpu().numpy()
    q32 = q.astype(np.float32).astype(np.float64)
    worst = 0.0
    for e in range(n):
        want = orc.self_distances(q32[e])
        assert np.array_equal(got[e] == 10.0, want == 10.0)
        worst = max(worst, np.abs(got[e] - want).max())
    print(f'self distances ({"hulls" if meshed else "primitives"}): max err {worst:.2e}')
    assert worst <= 2e-5


def test_environment_autocollision_api():
    """get_manipulator_collisions_with_itself / consider_autocollision through the drop-in Environment (reference
    environment.py:311-371, 394-412)."""
    from robotic_manipulator_rloa_b200.environment.environment import Environment, EnvironmentConfiguration
    cfg = KUKA
    env = Environment(cfg['file'], EnvironmentConfiguration(
        endeffector_index=cfg['ee'], fixed_joints=cfg['fixed'], involved_joints=cfg['involved'],
        target_position=cfg['target'], obstacle_position=cfg['obstacle'], initial_joint_positions=cfg['start'],
        initial_positions_variation_range=[0] * 6, max_force=200., visualize=False))
    env.reset(verbose=False)
    d = env.get_manipulator_collisions_with_itself()
    nl = env.num_joints
    assert sorted(d) == sorted(f'joint_{i}' for i in range(nl))
    for i in range(nl):
        assert d[f'joint_{i}'].shape == (len([j for j in range(nl) if abs(j - i) > 1]),)
    _, orc = make_oracle(cfg)
    q, _ = env.sim.get_state()
    want = orc.self_distances(q[0].double().cpu().numpy())
    for i in range(nl):
        others = [j for j in range(nl) if abs(j - i) > 1]
        assert np.abs(d[f'joint_{i}'] - want[i, others]).max() <= 2e-5
    # CollisionDetector.compute_collisions_in_manipulator (reference utils/collision_detector.py:63-98): the same
    # distances per link, neighbours skipped, saturated at max_distance
    from robotic_manipulator_rloa_b200.utils.collision_detector import CollisionDetector, CollisionObject
    for link in (0, 5, nl - 1):
        det = CollisionDetector(CollisionObject(body=env.sim, link=link), obstacle_ids=[])
        got = det.compute_collisions_in_manipulator(list(range(nl)))
        others = [j for j in range(nl) if abs(j - link) > 1]
        assert got.shape == (len(others),) and np.abs(got - want[link, others]).max() <= 2e-5
        capped = det.compute_collisions_in_manipulator(list(range(nl)), max_distance=0.05)
        assert capped.max() <= 0.05 and np.array_equal(capped < 0.05, got < 0.05)
    hit = any((v < 0).any() for v in d.values())
    base_r, base_t = env.get_reward(), env.is_terminal_state()
    assert env.get_reward(consider_autocollision=True) == (-1000 if hit and base_r != 250 else base_r)
    assert env.is_terminal_state(consider_autocollision=True) == (1 if hit else base_t)
    env.close()


def test_cylinder_and_dense_hulls_against_oracle(tmp_path):
    """Vertex clouds with many coplanar / cocircular points (pybullet's 2 x 32-point cylinders) and a dense 300-point
    hull: the degenerate simplices GJK meets there must not cost accuracy."""
    arm = load_urdf(write_test_arm(str(tmp_path), lambda i: '<cylinder radius="%g" length="%g"/>' % (0.03 + 0.01 * i, 0.12 + 0.03 * i)))
    cfg = dict(ee=2, involved=[0, 1, 2], fixed=[], target=[0.2, 0.1, 0.5], obstacle=[-0.1, 0.2, 0.4])
    dense = hullified(make_oracle(KUKA)[0], n=300)
    for model, c, seed in ((arm, cfg, 71), (dense, KUKA, 72)):
        orc = BulletOracle(model, c['ee'], len(c['involved']))
        n = 512
        if model is arm:
            q = np.random.default_rng(seed).uniform(-2.4, 2.4, (n, 3))
        else:
            q, _ = random_states(model, n, seed=seed)
        ob, tg = spread_tasks(c, n, seed + 10, orc, q)
        sim = make_sim(model, c, n)
        sim.set_task(torch.as_tensor(tg, dtype=torch.float32), torch.as_tensor(ob, dtype=torch.float32))
        sim.set_state(q, np.zeros_like(q))
        _, link, ee = sim.observe(want_distances=True)
        link, ee = link.cpu().numpy(), ee.cpu().numpy()
        q32 = q.astype(np.float32).astype(np.float64)
        worst = 0.0
        for e in range(0, n, 2):
            lo, eet, _ = orc.distances(q32[e], ob[e], tg[e])
            worst = max(worst, np.abs(link[e] - lo).max(), abs(ee[e] - eet))
        print(f'{"cylinders" if model is arm else "300-point hulls"}: max distance err {worst:.2e}')
        assert worst <= 2e-5
