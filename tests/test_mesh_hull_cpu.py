"""
CPU tests of the mesh front end and of the oracle's GJK restatement: STL / OBJ readers,
<mesh> / <cylinder> collision elements as convex vertex clouds, and the hull / box distance queries against an
independent quadratic-programming solution and against the closed-form primitives.
"""
import os

import numpy as np
import pytest
from scipy.optimize import minimize
from scipy.spatial.transform import Rotation

from helpers import BOX_CORNERS, hullified, make_oracle, write_box_mesh, write_test_arm, KUKA, random_states
from oracle import bullet_oracle as bo
from robotic_manipulator_rloa_b200.environment import mesh_io
from robotic_manipulator_rloa_b200.environment.robot_model import ModelError, SHAPE_BOX, SHAPE_HULL, load_urdf


@pytest.mark.parametrize('fmt,ext', [('stl_binary', 'stl'), ('stl_ascii', 'stl'), ('obj', 'obj')])
def test_mesh_readers_return_the_vertices(tmp_path, fmt, ext):
    path = str(tmp_path / f'box.{ext}')
    half = (0.03, 0.05, 0.11)
    write_box_mesh(path, half, fmt)
    raw = mesh_io.load_mesh_vertices(path)
    assert raw.shape == ((8, 3) if fmt == 'obj' else (36, 3))
    cloud = mesh_io.convex_vertex_cloud(raw)
    want = np.unique(BOX_CORNERS * np.asarray(half), axis=0)
    assert cloud.shape == (8, 3) and np.allclose(np.unique(cloud, axis=0), want, atol=1e-7)


def test_hull_reduction_drops_interior_points_and_keeps_flat_clouds():
    rng = np.random.default_rng(0)
    inner = rng.uniform(-0.5, 0.5, (200, 3)) * 0.9
    cloud = mesh_io.convex_vertex_cloud(np.concatenate([BOX_CORNERS, inner]))
    assert cloud.shape == (8, 3)
    flat = np.concatenate([rng.uniform(-1, 1, (20, 2)), np.zeros((20, 1))], axis=1)     # Qhull refuses: kept as is
    assert mesh_io.convex_vertex_cloud(flat).shape == (20, 3)


def test_cylinder_cloud_is_two_rims():
    c = mesh_io.cylinder_vertex_cloud(0.04, 0.3)
    assert c.shape == (64, 3)
    assert np.allclose(np.hypot(c[:, 0], c[:, 1]), 0.04) and np.allclose(np.abs(c[:, 2]), 0.15)


def test_bad_mesh_files_raise(tmp_path):
    with pytest.raises(mesh_io.MeshError):
        mesh_io.load_mesh_vertices(str(tmp_path / 'missing.stl'))
    p = tmp_path / 'x.dae'
    p.write_text('<COLLADA/>')
    with pytest.raises(mesh_io.MeshError):
        mesh_io.load_mesh_vertices(str(p))
    with pytest.raises(ModelError):
        load_urdf(write_test_arm(str(tmp_path), lambda i: '<mesh filename="meshes/nope.stl"/>'))


def _load_arm_pair(tmp_path):
    half = {1: (0.04, 0.05, 0.1), 2: (0.03, 0.03, 0.09), 3: (0.02, 0.06, 0.05)}
    (tmp_path / 'prim').mkdir()
    (tmp_path / 'mesh' / 'meshes').mkdir(parents=True)
    fmts = {1: ('stl_binary', 'stl'), 2: ('stl_ascii', 'stl'), 3: ('obj', 'obj')}
    for i, (fmt, ext) in fmts.items():
        write_box_mesh(str(tmp_path / 'mesh' / 'meshes' / f'l{i}.{ext}'), (0.5, 0.5, 0.5), fmt)   # unit cube, scaled below
    prim = load_urdf(write_test_arm(str(tmp_path / 'prim'),
                                    lambda i: '<box size="%g %g %g"/>' % tuple(2 * np.asarray(half[i]))))
    mesh = load_urdf(write_test_arm(str(tmp_path / 'mesh'),
                                    lambda i: '<mesh filename="package://arm/meshes/l%d.%s" scale="%g %g %g"/>'
                                    % ((i, fmts[i][1]) + tuple(2 * np.asarray(half[i])))))
    return prim, mesh


def test_urdf_mesh_elements_load_as_hulls(tmp_path):
    prim, mesh = _load_arm_pair(tmp_path)
    assert list(prim.s_type) == [SHAPE_BOX] * 3 and list(mesh.s_type) == [SHAPE_HULL] * 3
    assert mesh.verts.shape == (24, 3) and list(mesh.s_vn) == [8, 8, 8] and list(mesh.s_v0) == [0, 8, 16]
    assert np.allclose(mesh.s_dim[:, 0], mesh_io.HULL_MARGIN)
    for s in range(3):                       # scale applied: the cloud spans the primitive's half extents
        assert np.allclose(np.abs(mesh.verts[8 * s:8 * s + 8]).max(axis=0), prim.s_dim[s])
    assert np.allclose(mesh.s_R, prim.s_R) and np.allclose(mesh.s_p, prim.s_p)


def test_urdf_cylinder_loads_as_hull(tmp_path):
    m = load_urdf(write_test_arm(str(tmp_path), lambda i: '<cylinder radius="0.05" length="0.2"/>'))
    assert list(m.s_type) == [SHAPE_HULL] * 3 and m.verts.shape == (192, 3)


def test_oracle_distances_box_mesh_equals_box_primitive(tmp_path):
    """The same arm with <box> primitives and with box meshes: link distances differ by the hull margin only, and the
    end-effector/target distance (box vs cube and hull vs cube, both GJK) by the margin."""
    prim, mesh = _load_arm_pair(tmp_path)
    op, om = bo.BulletOracle(prim, 2, 3), bo.BulletOracle(mesh, 2, 3)
    rng = np.random.default_rng(3)
    for _ in range(300):
        q = rng.uniform(-2.4, 2.4, 3)
        ob, tg = rng.uniform(-0.5, 0.5, 3) + [0, 0, 0.4], rng.uniform(-0.5, 0.5, 3) + [0, 0, 0.4]
        lp, ep, _ = op.distances(q, ob, tg)
        lm, em, _ = om.distances(q, ob, tg)
        inside = lp < -0.075 + 1e-9            # sphere centre inside the box: the hull path reports the overlap bound
        assert np.abs((lm + mesh_io.HULL_MARGIN - lp)[~inside]).max() <= 1e-9
        assert np.all(lm[inside] <= -0.075)
        if ep > 1e-6:
            assert abs(em + mesh_io.HULL_MARGIN - ep) <= 1e-9


def _qp_distance(W, bc, bh, rng):
    n = len(W)

    def f(x):
        d = x[:n] @ W - x[n:]
        return d @ d
    cons = [{'type': 'eq', 'fun': lambda x: x[:n].sum() - 1}]
    bnds = [(0, 1)] * n + [(bc[i] - bh[i], bc[i] + bh[i]) for i in range(3)]
    best = np.inf
    for _ in range(3):
        x0 = np.concatenate([rng.dirichlet(np.ones(n)), bc])
        best = min(best, minimize(f, x0, method='SLSQP', bounds=bnds, constraints=cons,
                                  options={'ftol': 1e-16, 'maxiter': 500}).fun)
    return np.sqrt(max(best, 0.0))


def test_oracle_gjk_against_quadratic_program():
    rng = np.random.default_rng(0)
    for trial in range(60):
        V = rng.normal(size=(rng.integers(4, 24), 3)) * rng.uniform(0.02, 0.3, size=3)
        R = Rotation.random(random_state=trial).as_matrix()
        p, bc = rng.normal(size=3) * 0.3, rng.normal(size=3) * 0.4
        bh = np.full(3, 0.025) if trial % 2 else np.zeros(3)
        d, it = bo.gjk_hull_box(V, R.reshape(9), p, bc, bh)
        assert it <= 20
        assert abs(d - _qp_distance((R @ V.T).T + p, bc, bh, rng)) <= 1e-6


def test_oracle_gjk_closed_forms():
    rng = np.random.default_rng(1)
    h = np.array([0.1, 0.05, 0.2])
    eye = np.eye(3).reshape(9)
    for _ in range(500):
        pt = rng.normal(size=3) * 0.3
        o = np.abs(pt) - h
        want = np.sqrt((np.maximum(o, 0) ** 2).sum())          # 0 inside
        d, _ = bo.gjk_hull_box(BOX_CORNERS * h, eye, np.zeros(3), pt, np.zeros(3))
        assert abs(d - want) <= 1e-12
    # two axis-aligned boxes: per-axis gaps
    for _ in range(500):
        c = rng.normal(size=3) * 0.4
        gap = np.maximum(np.abs(c) - h - 0.025, 0)
        d, _ = bo.gjk_hull_box(BOX_CORNERS * h, eye, np.zeros(3), c, np.full(3, 0.025))
        assert abs(d - np.sqrt((gap ** 2).sum())) <= 1e-12
    # rigid motion invariance
    V = rng.normal(size=(30, 3)) * 0.1
    for k in range(50):
        R = Rotation.random(random_state=k).as_matrix()
        p, pt = rng.normal(size=3), rng.normal(size=3) * 0.5
        d0, _ = bo.gjk_hull_box(V, eye, np.zeros(3), pt, np.zeros(3))
        d1, _ = bo.gjk_hull_box(V, R.reshape(9), p, R @ pt + p, np.zeros(3))
        assert abs(d0 - d1) <= 1e-10


def test_hullified_kuka_tracks_the_primitive_model():
    """An inscribed 40-point polytope per primitive: distances can only grow, by less than the sagitta of the cloud."""
    cfg = KUKA
    model, orc = make_oracle(cfg)
    hm = hullified(model, margin=0.0)
    oh = bo.BulletOracle(hm, cfg['ee'], len(cfg['involved']))
    q, _ = random_states(model, 40, seed=5)
    for e in range(40):
        lp, ep, _ = orc.distances(q[e], cfg['obstacle'], cfg['target'])
        lh, eh, _ = oh.distances(q[e], cfg['obstacle'], cfg['target'])
        sep = lp > -0.075                       # the overlap bound replaces the depth when the centre is inside
        assert np.all(lh[sep] >= lp[sep] - 1e-9) and np.all(lh[sep] <= lp[sep] + 0.02)
        assert eh >= ep - 1e-9 and eh <= ep + 0.02


# ---- link-link distances (get_manipulator_collisions_with_itself) ---------------------------
def _segment_segment(p1, q1, p2, q2):
    f = lambda x: np.sum((p1 + x[0] * (q1 - p1) - p2 - x[1] * (q2 - p2)) ** 2)
    starts = ([.5, .5], [0, 0], [1, 1], [0, 1], [1, 0])
    return np.sqrt(min(minimize(f, x0, bounds=[(0, 1), (0, 1)], method='L-BFGS-B',
                                options={'ftol': 1e-18, 'gtol': 1e-14}).fun for x0 in starts))


def test_oracle_self_distances_structure_and_capsule_pairs():
    cfg = KUKA
    model, orc = make_oracle(cfg)
    q, _ = random_states(model, 30, seed=3)
    nl = model.nl
    separated = 0
    for e in range(30):
        D = orc.self_distances(q[e])
        assert np.array_equal(D, D.T)
        for i in range(nl):
            assert D[i, i] == 10.0 and (i == 0 or D[i, i - 1] == 10.0)
        Rw, pw = orc.fk(q[e])

        def seg(s):
            l = model.s_link[s]
            R = Rw[l] @ model.s_R[s].reshape(3, 3)
            p = pw[l] + Rw[l] @ model.s_p[s]
            return p - model.s_dim[s][1] * R[:, 2], p + model.s_dim[s][1] * R[:, 2], model.s_dim[s][0]
        for a, b in [(0, 2), (0, 3), (1, 4), (2, 5), (0, 5), (1, 3)]:        # links 0-5 carry one capsule each
            a0, a1, ra = seg(a)
            b0, b1, rb = seg(b)
            core = _segment_segment(a0, a1, b0, b1)
            separated += core > 1e-3
            assert abs(D[a, b] - (core - ra - rb)) <= 1e-7
    assert separated > 100


def test_sdf_mesh_element_with_uri_and_scale(tmp_path):
    """The shipped KUKA SDF with its gripper-base <box> swapped for a unit-cube mesh + <scale>: same model otherwise,
    the box becomes an 8-vertex hull with the box's half extents, and the oracle distances move by the margin."""
    import os
    import shutil
    from robotic_manipulator_rloa_b200.environment.robot_model import DATA_PATH, load_sdf
    src = os.path.join(DATA_PATH, 'kuka_iiwa', 'kuka_with_gripper2.sdf')
    text = open(src).read()
    box = '<geometry><box><size>0.05 0.1 0.04</size></box></geometry>'
    assert text.count(box) == 1
    text = text.replace(box, '<geometry><mesh><uri>model://kuka/meshes/cube.obj</uri><scale>0.05 0.1 0.04</scale></mesh></geometry>')
    (tmp_path / 'meshes').mkdir()
    write_box_mesh(str(tmp_path / 'meshes' / 'cube.obj'), (0.5, 0.5, 0.5), 'obj')
    (tmp_path / 'kuka.sdf').write_text(text)
    prim, mesh = load_sdf(src), load_sdf(str(tmp_path / 'kuka.sdf'))
    s = int(np.flatnonzero(prim.s_type == SHAPE_BOX)[0])
    assert mesh.s_type[s] == SHAPE_HULL and mesh.s_vn[s] == 8 and mesh.verts.shape == (8, 3)
    assert np.allclose(np.abs(mesh.verts).max(axis=0), prim.s_dim[s])
    keep = np.arange(prim.ns) != s
    assert np.array_equal(mesh.s_type[keep], prim.s_type[keep]) and np.allclose(mesh.E0, prim.E0)
    op, om = bo.BulletOracle(prim, 13, 6), bo.BulletOracle(mesh, 13, 6)
    q, _ = random_states(prim, 50, seed=9)
    link = int(prim.s_link[s])
    for e in range(50):
        lp, ep, _ = op.distances(q[e], KUKA['obstacle'], KUKA['target'])
        lm, em, _ = om.distances(q[e], KUKA['obstacle'], KUKA['target'])
        other = np.arange(prim.nl) != link
        assert np.array_equal(lp[other], lm[other]) and ep == em
        if lp[link] > -0.075:
            assert abs(lm[link] + mesh_io.HULL_MARGIN - lp[link]) <= 1e-9


def test_oracle_gjk_degenerate_clouds():
    """Single vertex, segment, planar polygon, duplicated vertices, query point inside / on the hull."""
    eye, z3 = np.eye(3).reshape(9), np.zeros(3)
    # one vertex: plain point distances
    d, _ = bo.gjk_hull_box(np.array([[0.1, 0.2, 0.3]]), eye, z3, np.array([0.1, 0.2, 1.3]), z3)
    assert abs(d - 1.0) <= 1e-12
    # two vertices: point-segment
    seg = np.array([[0.0, 0, 0], [1.0, 0, 0]])
    for pt, want in (([0.5, 0.3, 0], 0.3), ([-0.4, 0.3, 0], 0.5), ([2.0, 0, 0], 1.0), ([0.25, 0, 0], 0.0)):
        d, _ = bo.gjk_hull_box(seg, eye, z3, np.array(pt, float), z3)
        assert abs(d - want) <= 1e-9
    # planar square (a flat "hull"): distance to the plate, from above and from the side
    sq = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], float)
    for pt, want in (([0.2, 0.3, 0.7], 0.7), ([2.0, 0, 0], 1.0), ([2.0, 2.0, 0], 2 ** 0.5), ([0.3, -0.2, 0], 0.0)):
        d, _ = bo.gjk_hull_box(sq, eye, z3, np.array(pt, float), z3)
        assert abs(d - want) <= 1e-9
    # duplicated / interior vertices do not change the answer
    rng = np.random.default_rng(4)
    cube = BOX_CORNERS * 0.5
    noisy = np.concatenate([cube, cube, rng.uniform(-0.4, 0.4, (50, 3))])
    for _ in range(100):
        pt = rng.normal(size=3)
        a, _ = bo.gjk_hull_box(cube, eye, z3, pt, z3)
        b, it = bo.gjk_hull_box(noisy, eye, z3, pt, z3)
        assert abs(a - b) <= 1e-12 and it <= 16
    # the same cube against a box that touches it, overlaps it, and sits inside it
    for c, want in (([1.0, 0, 0], 0.25), ([0.75, 0, 0], 0.0), ([0.6, 0.1, 0], 0.0), ([0.0, 0, 0], 0.0)):
        d, _ = bo.gjk_hull_box(cube, eye, z3, np.array(c, float), np.full(3, 0.25))
        assert abs(d - want) <= 1e-12


def test_loader_mesh_edge_cases(tmp_path):
    """Non-uniform and negative mesh scales, a mesh referenced by an absolute path, and a link with two collision
    elements (mesh + primitive): shape order, vertex ranges and bounding data stay consistent."""
    (tmp_path / 'm').mkdir()
    cube = str(tmp_path / 'm' / 'cube.stl')
    write_box_mesh(cube, (0.5, 0.5, 0.5), 'stl_binary')

    def geom(i):
        if i == 1:
            return f'<mesh filename="{cube}" scale="0.1 -0.2 0.3"/>'          # absolute path, mirrored in y
        if i == 2:
            return '<mesh filename="m/cube.stl" scale="0.05 0.05 0.05"/>'
        return '<sphere radius="0.04"/>'
    path = write_test_arm(str(tmp_path), geom)
    text = open(path).read().replace(
        '</link><link name="l3">', '<collision><origin xyz="0 0 0.2"/><geometry><box size="0.02 0.02 0.02"/></geometry>'
                                   '</collision></link><link name="l3">')
    open(path, 'w').write(text)
    m = load_urdf(path)
    assert list(m.s_link) == [0, 1, 1, 2] and list(m.s_type) == [SHAPE_HULL, SHAPE_HULL, SHAPE_BOX, 1]
    assert list(m.s_vn) == [8, 8, 0, 0] and list(m.s_v0) == [0, 8, 16, 16] and m.verts.shape == (16, 3)
    assert np.allclose(np.abs(m.verts[:8]).max(axis=0), [0.05, 0.1, 0.15])
    assert np.allclose(np.abs(m.verts[8:]).max(axis=0), [0.025, 0.025, 0.025])
    orc = bo.BulletOracle(m, 2, 3)
    lo, ee, _ = orc.distances(np.zeros(3), [0.5, 0.5, 0.5], [0.0, 0.0, 1.2])
    assert np.all(np.isfinite(lo)) and lo.max() < 10.0 and np.isfinite(ee)


def test_large_hulls_are_thinned_to_support_points():
    """Hulls above MAX_SHAPE_VERTS keep only support points: a subset of the original vertices (so the thinned hull is
    inside the original), within the reported deviation, and GJK distances move by no more than that deviation."""
    rng = np.random.default_rng(7)
    k = np.arange(20000) + 0.5
    phi, th = np.arccos(1 - 2 * k / 20000), np.pi * (1 + 5 ** 0.5) * k
    pts = np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1) * [0.05, 0.07, 0.16]
    thin, dev = mesh_io.thin_vertex_cloud(pts, mesh_io.MAX_SHAPE_VERTS)
    # an inscribed 512-vertex polytope of a smooth body with a 0.16 m half axis cannot be closer than ~0.8 mm (sagitta)
    assert thin.shape[0] <= mesh_io.MAX_SHAPE_VERTS and dev <= 1.5e-3
    full_set = {tuple(p) for p in pts}
    assert all(tuple(p) in full_set for p in thin)
    cloud = mesh_io.convex_vertex_cloud(pts)                 # the loader's entry point applies the same thinning
    assert cloud.shape[0] <= mesh_io.MAX_SHAPE_VERTS
    eye, z3 = np.eye(3).reshape(9), np.zeros(3)
    for _ in range(50):
        q = rng.normal(size=3)
        q *= rng.uniform(0.2, 0.5) / np.linalg.norm(q)
        a, _ = bo.gjk_hull_box(pts, eye, z3, q, z3)
        b, _ = bo.gjk_hull_box(thin, eye, z3, q, z3)
        assert -1e-12 <= b - a <= dev + 1e-9


def test_real_pybullet_data_is_preferred_when_installed(tmp_path, monkeypatch):
    """With pybullet_data importable, relative names resolve there first (the reference's search path,
    environment.py:210); RLOA_ASSETS=standin keeps the shipped models."""
    import sys
    import types
    from robotic_manipulator_rloa_b200.environment import robot_model as rm
    data = tmp_path / 'pbdata'                   # not the CWD: a relative name must not resolve "as given"
    (data / 'kuka_iiwa').mkdir(parents=True)
    fake_file = data / 'kuka_iiwa' / 'kuka_with_gripper2.sdf'
    fake_file.write_text(open(os.path.join(rm.DATA_PATH, 'kuka_iiwa', 'kuka_with_gripper2.sdf')).read())
    fake = types.ModuleType('pybullet_data')
    fake.getDataPath = lambda: str(data)
    monkeypatch.setitem(sys.modules, 'pybullet_data', fake)
    monkeypatch.delenv('RLOA_ASSETS', raising=False)
    assert rm.resolve_manipulator_file('kuka_iiwa/kuka_with_gripper2.sdf') == str(fake_file)
    assert rm.resolve_manipulator_file('franka_panda/panda.urdf').startswith(rm.DATA_PATH)     # not in the fake dir
    monkeypatch.setenv('RLOA_ASSETS', 'standin')
    assert rm.resolve_manipulator_file('kuka_iiwa/kuka_with_gripper2.sdf').startswith(rm.DATA_PATH)
    assert rm.load_manipulator('kuka_iiwa/kuka_with_gripper2.sdf').nl == 14


def test_unloadable_pybullet_data_file_falls_back_to_the_stand_in(tmp_path, monkeypatch):
    import sys
    import types
    from robotic_manipulator_rloa_b200.environment import robot_model as rm
    data = tmp_path / 'pbdata'
    (data / 'kuka_iiwa').mkdir(parents=True)
    (data / 'kuka_iiwa' / 'kuka_with_gripper2.sdf').write_text('<sdf version="1.6"><model name="m"><link name="a"/>'
                                                              '<link name="b"/></model></sdf>')       # two roots
    fake = types.ModuleType('pybullet_data')
    fake.getDataPath = lambda: str(data)
    monkeypatch.setitem(sys.modules, 'pybullet_data', fake)
    monkeypatch.delenv('RLOA_ASSETS', raising=False)
    assert rm.load_manipulator('kuka_iiwa/kuka_with_gripper2.sdf').nl == 14          # the stand-in
    with pytest.raises(ModelError):                                                   # an explicit path gets no fallback
        rm.load_manipulator(str(data / 'kuka_iiwa' / 'kuka_with_gripper2.sdf'))
