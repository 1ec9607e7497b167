"""
CPU pins of the physics oracle (oracle/bullet_restatement.c).  PyBullet itself is absent ("parity unpinned",
DESIGN.md section 2), so the restatement is held to what CAN be checked without it:
  * the closed-form regime of Bullet's motor model — exact to solver tolerance;
  * rigid-body identities that any correct articulated-body implementation satisfies: M^-1 from unit impulses
    through the ABA factors == inverse of the CRBA mass matrix; ABA accelerations == M^-1 (tau - bias);
  * the reward / done truth table and the state layout pinned by the reference's mocked tests
    (tests/robotic_manipulator_rloa/environment/test_environment.py:219-265, 514-520; environment.py:311-371, 431-451).
"""
import numpy as np
import pytest

from helpers import KUKA, PANDA, XARM6, make_oracle, random_states, step_motors


@pytest.mark.parametrize('cfg', [KUKA, PANDA], ids=['kuka', 'panda'])
def test_model_shape_matches_the_survey(cfg):
    model, _ = make_oracle(cfg)
    nl, ndof = model.nl, int((np.asarray(model.jtype) != 0).sum())
    assert (nl, ndof) == ((14, 12) if cfg is KUKA else (12, 9))          # SURVEY.md section 8 sizes
    assert all(p < i for i, p in enumerate(model.parent))                # depth-first numbering


@pytest.mark.parametrize('cfg', [KUKA, PANDA], ids=['kuka', 'panda'])
def test_minv_is_the_inverse_of_the_crba_mass_matrix(cfg):
    model, orc = make_oracle(cfg)
    mov = np.nonzero(np.asarray(model.jtype) != 0)[0]
    q, _ = random_states(model, 8, seed=2)
    for e in range(8):
        Mi = orc.minv(q[e])[np.ix_(mov, mov)]
        Mm = orc.crba(q[e])[np.ix_(mov, mov)]
        assert np.allclose(Mm, Mm.T, atol=1e-12) and np.all(np.linalg.eigvalsh(Mm) > 0)
        assert np.abs(Mi @ Mm - np.eye(len(mov))).max() <= 1e-8


def test_aba_equals_minv_times_generalised_force():
    """qdd = M^-1 (tau - C(q, qd)): the ABA (with Bullet's link velocity drag, gyroscopic terms, gravity and joint
    damping) against an independent recursive Newton-Euler bias and the unit-impulse M^-1."""
    model, orc = make_oracle(KUKA)
    mov = np.nonzero(np.asarray(model.jtype) != 0)[0]
    q, qd = random_states(model, 6, seed=4)
    rng = np.random.default_rng(0)
    for e in range(6):
        tau = np.zeros(model.nl); tau[mov] = rng.uniform(-5, 5, len(mov))
        qdd = orc.aba(q[e], qd[e], tau)
        rhs = tau - orc.bias(q[e], qd[e])                 # the RNEA bias already carries damping * qd and the drag
        want = orc.minv(q[e])[np.ix_(mov, mov)] @ rhs[mov]
        assert np.abs(qdd[mov] - want).max() <= 1e-7 * max(1.0, np.abs(want).max())


def test_unsaturated_motors_are_kinematic():
    """A.6: velocity-controlled joints reach the commanded velocity, q += a/240; position-held joints close 10 %
    of their error per step (kp 0.1) and get velocity -24 q."""
    cfg = KUKA
    model, orc = make_oracle(cfg)
    step_motors(orc, cfg)
    n = 64
    rng = np.random.default_rng(3)
    q = np.zeros((n, model.nl)); qd = np.zeros((n, model.nl))
    q[:, :6] = rng.uniform(-0.5, 0.5, (n, 6))
    q[:, 6] = 0.2
    a = rng.uniform(-0.02, 0.02, (n, 6))
    qd[:, :6] = a
    q0 = q.copy()
    orc.batch_step(q, qd, a, cfg['involved'], 200.0, cfg['obstacle'], cfg['target'])
    # PGS stops at a squared residual of 1e-7: a few times sqrt(1e-7) = 3e-4 rad/s from the exact solution
    assert np.abs(qd[:, :6] - a).max() <= 2e-3
    assert np.abs(q[:, :6] - (q0[:, :6] + qd[:, :6] / 240.0)).max() <= 1e-12       # semi-implicit Euler, exactly
    assert np.abs(q[:, 6] - 0.9 * 0.2).max() <= 1e-4
    assert np.abs(qd[:, 6] + 24.0 * 0.2).max() <= 2e-2


def test_reset_drives_towards_the_start_pose():
    """Environment.reset = 50 POSITION_CONTROL steps (environment.py:295-301): residual 0.9^50 of the error."""
    cfg = KUKA
    model, orc = make_oracle(cfg)
    n = 4
    q = np.zeros((n, model.nl)); qd = np.zeros((n, model.nl))
    init = np.tile(np.asarray(cfg['start'], dtype=np.float64), (n, 1))
    orc.batch_reset(q, qd, init, 50)
    resid = 0.9 ** 50
    assert np.abs(q[:, :6] - init * (1 - resid)).max() <= 2e-3
    assert np.abs(q[:, 6:]).max() <= 5e-3                                         # never-commanded joints stay put


def test_reward_and_done_truth_table():
    """get_reward / is_terminal_state (environment.py:311-371): 250 & done on target; -1000 & done on obstacle;
    -(d - 0.05) & not done otherwise; target wins the reward when both hold, done either way."""
    cfg = KUKA
    model, orc = make_oracle(cfg)
    q = np.zeros(model.nl); q[:6] = cfg['start']
    qd = np.zeros(model.nl)
    _, ee = orc.fk(q)
    ee = ee[cfg['ee']]
    far = ee + np.array([0.0, 0.0, 5.0])
    obs, rew, done = orc.observe(q, qd, far, far)                                  # nothing near
    lo, d_ee, eep = orc.distances(q, far, far)
    assert done == 0 and rew == pytest.approx(-(d_ee - 0.05), abs=1e-12) and rew < 0
    assert obs.shape == (21,)                                                      # 9 + 2 * 6 (environment.py:261)
    assert np.allclose(obs[:6], q[:6]) and np.allclose(obs[6:12], qd[:6])          # joints 0..n-1 (:442-444)
    assert np.allclose(obs[12:15], eep) and np.allclose(obs[15:18], far) and np.allclose(obs[18:21], far)
    _, rew, done = orc.observe(q, qd, far, ee)                                     # target on the end effector
    assert (rew, done) == (250.0, 1)
    _, rew, done = orc.observe(q, qd, ee, far)                                     # obstacle on the end effector
    assert (rew, done) == (-1000.0, 1)
    _, rew, done = orc.observe(q, qd, ee, ee)                                      # both: reward 250, done
    assert (rew, done) == (250.0, 1)
    lo, _, _ = orc.distances(q, far, far)
    assert (lo[np.isin(np.arange(model.nl), np.asarray(model.s_link))] < 10.0).all()
    no_shape = ~np.isin(np.arange(model.nl), np.asarray(model.s_link))
    assert (lo[no_shape] == 10.0).all()                                            # collision_detector.py:56-57


def test_joint_limits_hold():
    """A.4: a joint commanded through its limit stops there (unilateral row, erp 0.2 push-out)."""
    cfg = KUKA
    model, orc = make_oracle(cfg)
    step_motors(orc, cfg)
    q = np.zeros((1, model.nl)); qd = np.zeros((1, model.nl))
    q[0, 1] = model.upper[1] - 1e-3
    a = np.zeros((1, 6)); a[0, 1] = 1.0
    for _ in range(40):
        orc.batch_step(q, qd, a, cfg['involved'], 200.0, [5, 5, 5], [5, 5, 6])
    assert q[0, 1] <= model.upper[1] + 2e-3 and abs(qd[0, 1]) <= 5e-2


def test_saturated_step_satisfies_the_lcp_optimality_conditions():
    """Whatever sweep order the solver uses, the solution of one step must satisfy the complementarity conditions of the
    motor rows (no joint limit active here).  The joint-space impulse is recovered independently of the solver from
    lambda = M (qd' - qd_free), with M from the CRBA and the free velocity from the ABA; then per motor row
      |lambda| <  max impulse  ->  the joint moves at its target velocity  v* = kp (q_des - q) / dt + qd_des
      |lambda| == max impulse  ->  the impulse pushes towards the target (sign(lambda) = sign(v* - qd'))
    up to the solver's own stopping residual (1e-7 on the squared velocity change)."""
    cfg = KUKA
    model, orc = make_oracle(cfg)
    step_motors(orc, cfg)
    dt, nl = orc.m.dt, model.nl
    mov = np.nonzero(np.asarray(model.jtype) != 0)[0]
    n = 200
    q0, qd0 = random_states(model, n, seed=17, vel=2.0, frac_limit=0.8, held=cfg['fixed'])
    rng = np.random.default_rng(18)
    actions = rng.uniform(-1, 1, (n, 6))
    q, qd = q0.copy(), qd0.copy()
    _, _, _, iters = orc.batch_step(q, qd, actions, cfg['involved'], 200.0, cfg['obstacle'], cfg['target'], nthreads=4)
    sat_rows = free_rows = 0
    for e in range(n):
        if iters[e] >= 50:                                    # not converged within Bullet's cap: nothing to check
            continue
        qs = qd0[e] + dt * orc.aba(q0[e], qd0[e])             # free velocity (no clamp reached at these speeds)
        M = orc.crba(q0[e])[np.ix_(mov, mov)]
        lam = M @ (qd[e] - qs)[mov]
        for r, j in enumerate(mov):
            if j in cfg['involved']:
                vstar, mx = actions[e, cfg['involved'].index(j)], 200.0 * dt
            else:                                             # POSITION_CONTROL target 0: kp 0.1, force 1e5
                vstar, mx = 0.1 * (0.0 - q0[e, j]) / dt, 100000.0 * dt
            assert abs(lam[r]) <= mx * (1 + 1e-6) + 1e-9
            if abs(lam[r]) < mx * (1 - 1e-6):
                free_rows += 1
                assert abs(qd[e, j] - vstar) <= 3e-3
            else:
                sat_rows += 1
                assert np.sign(lam[r]) == np.sign(vstar - qd[e, j])
    assert sat_rows > 50 and free_rows > 500


def test_closed_form_distances_against_gjk():
    """The closed forms behind get_reward / is_terminal_state (capsule-vs-cube = exact segment/box distance, box-vs-sphere
    = signed point/box distance, capsule-vs-sphere = point/segment) against the GJK restatement, which is itself
    pinned to a quadratic-programming solution (tests/test_mesh_hull_cpu.py)."""
    from helpers import BOX_CORNERS
    from oracle.bullet_oracle import gjk_hull_box
    cfg = KUKA
    model, orc = make_oracle(cfg)
    rng = np.random.default_rng(23)
    q, _ = random_states(model, 150, seed=23)
    eye, z3 = np.eye(3).reshape(9), np.zeros(3)
    half = np.full(3, 0.025)
    checked_box = 0
    for e in range(150):
        Rw, pw = orc.fk(q[e])
        ee_pos = pw[cfg['ee']]
        target = ee_pos + rng.uniform(-0.15, 0.15, 3)
        obstacle = pw[rng.integers(0, model.nl)] + rng.uniform(-0.2, 0.2, 3)
        lo, ee, _ = orc.distances(q[e], obstacle, target)
        for s in range(model.ns):
            l, kind = int(model.s_link[s]), int(model.s_type[s])
            R = Rw[l] @ model.s_R[s].reshape(3, 3)
            p = pw[l] + Rw[l] @ model.s_p[s]
            if kind == 2:                                    # capsule: a segment with a radius
                ends = np.stack([p - model.s_dim[s][1] * R[:, 2], p + model.s_dim[s][1] * R[:, 2]])
                core, _ = gjk_hull_box(ends, eye, z3, obstacle, z3)
                want = core - model.s_dim[s][0] - 0.075
                if l == cfg['ee']:
                    core_t, _ = gjk_hull_box(ends, eye, z3, target, half)
                    assert abs(ee - (core_t - model.s_dim[s][0])) <= 1e-9
            elif kind == 1:
                want = np.linalg.norm(obstacle - p) - model.s_dim[s][0] - 0.075
            else:                                            # box: compare where the sphere centre is outside the box
                core, _ = gjk_hull_box(BOX_CORNERS * model.s_dim[s], R.reshape(9), p, obstacle, z3)
                if core <= 1e-9:
                    continue
                want = core - 0.075
                checked_box += 1
            only = [t for t in range(model.ns) if int(model.s_link[t]) == l]
            if len(only) == 1:                               # the per-link value is the minimum over the link's shapes
                assert abs(lo[l] - want) <= 1e-9, (l, kind)
    assert checked_box > 20


@pytest.mark.parametrize('cfg', [KUKA, PANDA], ids=['kuka', 'panda'])
def test_aba_against_the_euler_lagrange_equations(cfg):
    """Independent of the Newton-Euler recursions: with Bullet's velocity drag and the joint damping switched off, the
    ABA accelerations must solve  M qdd + Mdot qd - dT/dq + dV/dq = tau  where M(q) is the CRBA mass matrix,
    T = 1/2 qd^T M qd and V = -sum_i m_i g . p_i(q) comes from the forward kinematics; derivatives by central
    differences.  This pins the Coriolis / centrifugal and gravity terms (signs, frames, COM offsets)."""
    model, orc = make_oracle(cfg)                         # KUKA: SDF, base pose, gripper tree; Panda: prismatic fingers
    m = orc.m
    m.lin_damp, m.ang_damp = 0.0, 0.0
    for i in range(model.nl):
        m.damping[i] = 0.0
    mov = np.nonzero(np.asarray(model.jtype) != 0)[0]
    g = np.array([m.gravity[0], m.gravity[1], m.gravity[2]])
    mass = np.asarray(model.mass, float)

    def M_of(q):
        return orc.crba(q)[np.ix_(mov, mov)]

    def V_of(q):
        return -float(sum(mass[i] * (g @ orc.fk(q)[1][i]) for i in range(model.nl)))

    q_all, qd_all = random_states(model, 5, seed=29, vel=1.5, frac_limit=0.7)
    rng = np.random.default_rng(30)
    h = 1e-5
    for e in range(5):
        q, qd = q_all[e], qd_all[e]
        tau = np.zeros(model.nl); tau[mov] = rng.uniform(-3, 3, len(mov))
        v = qd[mov]
        dM = []
        dV = np.zeros(len(mov))
        for r, j in enumerate(mov):
            qp, qm = q.copy(), q.copy()
            qp[j] += h; qm[j] -= h
            dM.append((M_of(qp) - M_of(qm)) / (2 * h))
            dV[r] = (V_of(qp) - V_of(qm)) / (2 * h)
        Mdot_v = sum(dM[r] * v[r] for r in range(len(mov))) @ v
        dT = np.array([0.5 * v @ dM[r] @ v for r in range(len(mov))])
        want = np.linalg.solve(M_of(q), tau[mov] - Mdot_v + dT - dV)
        got = orc.aba(q, qd, tau)[mov]
        assert np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max()), (e, np.abs(got - want).max())


def _direct_urdf_com_positions(path, q_by_joint):
    """Link COM world positions straight from the URDF semantics (joint origin, then the joint motion about / along the
    axis given in the child frame, inertial origin) with scipy rotations — no COM-frame algebra, no loader code."""
    import xml.etree.ElementTree as ET
    from scipy.spatial.transform import Rotation as Rot
    root = ET.parse(path).getroot()

    def origin(node):
        o = node.find('origin') if node is not None else None
        xyz = np.array([float(v) for v in (o.get('xyz', '0 0 0') if o is not None else '0 0 0').split()])
        rpy = [float(v) for v in (o.get('rpy', '0 0 0') if o is not None else '0 0 0').split()]
        return Rot.from_euler('xyz', rpy).as_matrix(), xyz          # extrinsic x, y, z = URDF fixed-axis rpy

    com = {l.get('name'): origin(l.find('inertial'))[1] for l in root.findall('link')}
    children = {}
    for j in root.findall('joint'):
        children.setdefault(j.find('parent').get('link'), []).append(j)
    child_links = {j.find('child').get('link') for j in root.findall('joint')}
    base = [l.get('name') for l in root.findall('link') if l.get('name') not in child_links][0]
    out = {}
    frames = {}

    def walk(link, R, p):
        out[link] = p + R @ com[link]
        frames[link] = (R, p)
        for j in children.get(link, []):
            Rj, pj = origin(j)
            Rc, pc = R @ Rj, p + R @ pj
            ax = j.find('axis')
            axis = np.array([float(v) for v in ax.get('xyz').split()]) if ax is not None else np.array([1.0, 0, 0])
            axis = axis / np.linalg.norm(axis)
            qj = q_by_joint.get(j.get('name'), 0.0)
            if j.get('type') in ('revolute', 'continuous'):
                Rc = Rc @ Rot.from_rotvec(axis * qj).as_matrix()
            elif j.get('type') == 'prismatic':
                pc = pc + Rc @ (axis * qj)
            walk(j.find('child').get('link'), Rc, pc)

    walk(base, np.eye(3), np.zeros(3))
    shapes = {}                                   # link -> [(world rotation, world position) of every collision origin]
    for l in root.findall('link'):
        R, p = frames[l.get('name')]
        shapes[l.get('name')] = [(R @ origin(c)[0], p + R @ origin(c)[1]) for c in l.findall('collision')]
    out['__shapes__'] = shapes
    return out


@pytest.mark.parametrize('cfg', [PANDA, XARM6], ids=['panda', 'xarm6'])
def test_loader_and_fk_against_direct_urdf_kinematics(cfg):
    """The loader moves every link frame to its centre of mass with principal axes and the oracle's FK works in those
    frames (Bullet's btMultiBody layout); the COM world positions must equal a direct walk of the URDF tree."""
    from robotic_manipulator_rloa_b200.environment.robot_model import resolve_manipulator_file
    model, orc = make_oracle(cfg)
    path = resolve_manipulator_file(cfg['file'])
    q_all, _ = random_states(model, 6, seed=31, frac_limit=0.9)
    for e in range(6):
        want = _direct_urdf_com_positions(path, {model.joint_names[i]: q_all[e, i] for i in range(model.nl)})
        Rw, pw = orc.fk(q_all[e])
        for i in range(model.nl):
            assert np.abs(pw[i] - want[model.link_names[i]]).max() <= 1e-9, (model.link_names[i], pw[i], want[model.link_names[i]])
        # collision shapes: the loader re-expresses every <collision><origin> in the link's COM frame
        seen = {}
        for s in range(model.ns):
            l = int(model.s_link[s])
            k = seen.get(l, 0)
            seen[l] = k + 1
            R_want, p_want = want['__shapes__'][model.link_names[l]][k]
            assert np.abs(pw[l] + Rw[l] @ model.s_p[s] - p_want).max() <= 1e-9
            assert np.abs(Rw[l] @ model.s_R[s].reshape(3, 3) - R_want).max() <= 1e-9


@pytest.mark.parametrize('cfg', [KUKA, PANDA], ids=['kuka', 'panda'])
def test_mass_matrix_against_the_kinetic_energy_of_the_links(cfg):
    """1/2 qd^T M(q) qd (CRBA) must equal sum_i 1/2 m_i |v_i|^2 + 1/2 w_i^T R_i I_i R_i^T w_i with the link COM
    velocities and angular velocities taken by central differences of the forward kinematics along qd."""
    model, orc = make_oracle(cfg)
    mov = np.nonzero(np.asarray(model.jtype) != 0)[0]
    mass, I = np.asarray(model.mass, float), np.asarray(model.inertia, float)
    q_all, qd_all = random_states(model, 6, seed=37, vel=1.0, frac_limit=0.8)
    h = 1e-6
    for e in range(6):
        q, qd = q_all[e], qd_all[e]
        Rp, pp = orc.fk(q + h * qd)
        Rm, pm = orc.fk(q - h * qd)
        R0, _ = orc.fk(q)
        T = 0.0
        for i in range(model.nl):
            v = (pp[i] - pm[i]) / (2 * h)
            W = (Rp[i] - Rm[i]) / (2 * h) @ R0[i].T                       # [w]x = Rdot R^T
            w = np.array([W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]]) * 0.5
            wl = R0[i].T @ w
            T += 0.5 * mass[i] * (v @ v) + 0.5 * (wl * I[i]) @ wl
        M = orc.crba(q)[np.ix_(mov, mov)]
        assert abs(0.5 * qd[mov] @ M @ qd[mov] - T) <= 1e-7 * max(1.0, T)


def test_sdf_loader_and_fk_against_direct_sdf_kinematics():
    """The same pin for the SDF front end (the KUKA asset): SDF gives absolute link poses at q = 0 and joint axes in
    the child frame; the COM world positions from a direct walk must equal the loader + oracle FK."""
    import xml.etree.ElementTree as ET
    from scipy.spatial.transform import Rotation as Rot
    from robotic_manipulator_rloa_b200.environment.robot_model import resolve_manipulator_file
    cfg = KUKA
    model, orc = make_oracle(cfg)
    sdf = ET.parse(resolve_manipulator_file(cfg['file'])).getroot().find('.//model')

    def pose(node):
        v = [float(x) for x in (node.text if node is not None else '0 0 0 0 0 0').split()]
        return Rot.from_euler('xyz', v[3:]).as_matrix(), np.array(v[:3])

    Rm, pm = pose(sdf.find('pose'))
    zero, com = {}, {}
    for l in sdf.findall('link'):
        Rl, pl = pose(l.find('pose'))
        zero[l.get('name')] = (Rm @ Rl, pm + Rm @ pl)
        com[l.get('name')] = pose(l.find('inertial').find('pose'))[1] if l.find('inertial') is not None else np.zeros(3)
    joints = {j.findtext('child'): j for j in sdf.findall('joint') if j.findtext('parent') != 'world'}
    q_all, _ = random_states(model, 6, seed=41, frac_limit=0.9)
    for e in range(6):
        qj = {model.joint_names[i]: q_all[e, i] for i in range(model.nl)}
        world = {}

        def place(link):
            if link in world:
                return world[link]
            if link not in joints:
                world[link] = zero[link]
                return world[link]
            j = joints[link]
            Rp, pp = place(j.findtext('parent'))
            Rp0, pp0 = zero[j.findtext('parent')]
            Rc0, pc0 = zero[link]
            R = Rp @ (Rp0.T @ Rc0)
            p = pp + Rp @ (Rp0.T @ (pc0 - pp0))
            ax = j.find('axis')
            if ax is not None and j.get('type') != 'fixed':
                axis = np.array([float(x) for x in ax.findtext('xyz').split()])
                if ax.findtext('use_parent_model_frame', '0').strip() in ('1', 'true'):
                    axis = Rc0.T @ (Rm @ axis)
                axis = axis / np.linalg.norm(axis)
                if j.get('type') == 'prismatic':
                    p = p + R @ (axis * qj[j.get('name')])
                else:
                    R = R @ Rot.from_rotvec(axis * qj[j.get('name')]).as_matrix()
            world[link] = (R, p)
            return world[link]

        _, pw = orc.fk(q_all[e])
        for i in range(model.nl):
            R, p = place(model.link_names[i])
            assert np.abs(pw[i] - (p + R @ com[model.link_names[i]])).max() <= 1e-9, model.link_names[i]


def test_non_diagonal_inertia_tensors_survive_the_principal_axes_transform(tmp_path):
    """A URDF whose links have full inertia tensors and rotated inertial frames: the loader diagonalises them into the
    COM frame Bullet wants.  1/2 qd^T M qd from the loaded model must equal the kinetic energy computed directly from
    the URDF quantities (mass, COM offset, R_inertial I R_inertial^T) on a direct walk of the tree."""
    from scipy.spatial.transform import Rotation as Rot
    from oracle.bullet_oracle import BulletOracle
    from robotic_manipulator_rloa_b200.environment.robot_model import load_urdf
    rng = np.random.default_rng(43)
    links, joints, spec = ['<link name="l0"/>'], [], {}
    axes = ['0 0 1', '0 1 0', '1 0 0', '0.6 0 0.8']
    for i in range(1, 5):
        A = rng.normal(size=(3, 3))
        I = A @ A.T * 0.01 + np.eye(3) * 0.005                       # a full SPD tensor
        com, rpy = rng.uniform(-0.05, 0.05, 3), rng.uniform(-0.8, 0.8, 3)
        spec[f'l{i}'] = (1.0 + i * 0.3, com, Rot.from_euler('xyz', rpy).as_matrix(), I)
        links.append(f'<link name="l{i}"><inertial><origin xyz="{com[0]} {com[1]} {com[2]}" rpy="{rpy[0]} {rpy[1]} {rpy[2]}"/>'
                     f'<mass value="{1.0 + i * 0.3}"/><inertia ixx="{I[0, 0]}" ixy="{I[0, 1]}" ixz="{I[0, 2]}" iyy="{I[1, 1]}" '
                     f'iyz="{I[1, 2]}" izz="{I[2, 2]}"/></inertial></link>')
        o, r = rng.uniform(-0.2, 0.2, 3), rng.uniform(-0.5, 0.5, 3)
        spec[f'j{i}'] = (o, Rot.from_euler('xyz', r).as_matrix(), np.array([float(v) for v in axes[i - 1].split()]))
        joints.append(f'<joint name="j{i}" type="revolute"><parent link="l{i - 1}"/><child link="l{i}"/>'
                      f'<origin xyz="{o[0]} {o[1]} {o[2]}" rpy="{r[0]} {r[1]} {r[2]}"/><axis xyz="{axes[i - 1]}"/>'
                      f'<limit lower="-3" upper="3" effort="10" velocity="10"/></joint>')
    path = str(tmp_path / 'arm.urdf')
    open(path, 'w').write('<?xml version="1.0"?><robot name="a">' + ''.join(links) + ''.join(joints) + '</robot>')
    model = load_urdf(path)
    orc = BulletOracle(model, 3, 4)

    def frames(q):                                   # URDF link frames and COM positions by a direct chain walk
        R, p, out = np.eye(3), np.zeros(3), []
        for i in range(1, 5):
            o, Rj, ax = spec[f'j{i}']
            p = p + R @ o
            R = R @ Rj @ Rot.from_rotvec(ax / np.linalg.norm(ax) * q[i - 1]).as_matrix()
            out.append((R.copy(), p + R @ spec[f'l{i}'][1]))
        return out

    h = 1e-6
    for _ in range(5):
        q, qd = rng.uniform(-2, 2, 4), rng.uniform(-1, 1, 4)
        f0, fp, fm = frames(q), frames(q + h * qd), frames(q - h * qd)
        T = 0.0
        for i in range(4):
            m, _, Ri, I = spec[f'l{i + 1}']
            v = (fp[i][1] - fm[i][1]) / (2 * h)
            W = (fp[i][0] - fm[i][0]) / (2 * h) @ f0[i][0].T
            w = 0.5 * np.array([W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]])
            Iw = f0[i][0] @ Ri @ I @ Ri.T @ f0[i][0].T
            T += 0.5 * m * (v @ v) + 0.5 * w @ Iw @ w
        assert abs(0.5 * qd @ orc.crba(q) @ qd - T) <= 1e-7 * max(1.0, T)


def test_joint_limit_rows_satisfy_their_complementarity_conditions():
    """Joints pushed a little past a limit (the velocity-level limit row of Bullet: target velocity
    v_l = -penetration * erp / dt, impulse >= 0 towards the inside) together with the motor of the same joint.
    With the total joint impulse tau = M (qd' - qd_free) recovered independently of the solver:
      qd' > v_l  ->  the limit is inactive and tau is the motor's impulse (|tau| <= max)
      qd' = v_l  ->  the limit impulse tau - lambda_motor is non-negative (towards the inside), the motor being saturated
                     towards its own target
      qd' < v_l  ->  never (the limit impulse cap of 100 is out of reach here)."""
    cfg = KUKA
    model, orc = make_oracle(cfg)
    step_motors(orc, cfg)
    dt, erp = orc.m.dt, orc.m.erp
    mov = np.nonzero(np.asarray(model.jtype) != 0)[0]
    lower, upper = np.asarray(model.lower), np.asarray(model.upper)
    rng = np.random.default_rng(47)
    n = 300
    q0, qd0 = random_states(model, n, seed=47, vel=0.3, frac_limit=0.5, held=cfg['fixed'])
    over = np.zeros((n, model.nl), int)                      # -1: below the lower limit, +1: above the upper one
    for e in range(n):
        for j in rng.choice(cfg['involved'], size=2, replace=False):
            side = rng.choice([-1, 1])
            q0[e, j] = (lower[j] if side < 0 else upper[j]) + side * rng.uniform(1e-4, 3e-3)
            over[e, j] = side
    actions = rng.uniform(-1, 1, (n, 6))
    q, qd = q0.copy(), qd0.copy()
    _, _, _, iters = orc.batch_step(q, qd, actions, cfg['involved'], 200.0, cfg['obstacle'], cfg['target'], nthreads=4)
    active = inactive = 0
    tol = 5e-3
    for e in range(n):
        if iters[e] >= 50:
            continue
        qs = qd0[e] + dt * orc.aba(q0[e], qd0[e])
        tau = orc.crba(q0[e])[np.ix_(mov, mov)] @ (qd[e] - qs)[mov]
        for r, j in enumerate(mov):
            if over[e, j] == 0:
                continue
            s = -over[e, j]                                   # +1: the limit pushes towards +q (lower limit), -1 the other way
            pen = (q0[e, j] - lower[j]) if over[e, j] < 0 else (upper[j] - q0[e, j])      # negative
            v_l = -pen * erp / dt                             # >= 0, in the limit's own (inward) direction
            v_in = s * qd[e, j]                               # joint velocity in that direction
            assert v_in >= v_l - tol, (e, j, v_in, v_l)
            mx = 200.0 * dt
            vstar = actions[e, cfg['involved'].index(j)]
            if v_in > v_l + tol:
                inactive += 1
                assert abs(tau[r]) <= mx * (1 + 1e-6) + 1e-9
            else:
                active += 1
                lam_motor = mx * np.sign(vstar - qd[e, j]) if abs(vstar - qd[e, j]) > tol else None
                if lam_motor is not None:
                    assert s * (tau[r] - lam_motor) >= -1e-6, (e, j, tau[r], lam_motor)
    assert active > 30 and inactive > 10, (active, inactive)
