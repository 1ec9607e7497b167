"""
Robot model front end: URDF / SDF -> flat Bullet-style multibody description.

Replaces what ``p.loadURDF`` / ``p.loadSDF`` do for the reference
(/root/reference/robotic_manipulator_rloa/environment/environment.py:224-238): the file is parsed on
the host once, links are numbered depth-first exactly like PyBullet numbers joints, every link frame
is moved to its centre of mass (principal axes), and the result is a set of flat arrays that
``rloa_model_create`` (include/rloa_b200.h) copies to the device.

``pybullet_data`` is not available in this image, so the package ships its own stand-in assets under
``robotic_manipulator_rloa_b200/data`` (same relative names as pybullet_data) and
:func:`resolve_manipulator_file` searches that directory the way
``p.setAdditionalSearchPath(pybullet_data.getDataPath())`` (environment.py:210) does.
Collision geometry: ``<sphere>``, ``<capsule>``, ``<box>`` primitives, and — like pybullet's importer —
``<mesh>`` (binary / ASCII STL, OBJ) and ``<cylinder>`` elements as convex vertex clouds with a 0.001 margin
(:mod:`mesh_io`).
"""
from __future__ import annotations

import logging
import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import mesh_io

FIXED, REVOLUTE, PRISMATIC = 0, 1, 2
SHAPE_SPHERE, SHAPE_CAPSULE, SHAPE_BOX, SHAPE_HULL = 1, 2, 3, 4
MAX_LINKS = 32
MAX_SHAPES = 32
MAX_VERTS = 16384              # hull vertices per model (RLOA_MAX_HULL_VERTS)

logger = logging.getLogger(__name__)

DATA_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'data')


class ModelError(Exception):
    """Raised for files that cannot be turned into a simulator model."""


def rpy_to_R(r: float, p: float, y: float) -> np.ndarray:
    """URDF fixed-axis roll/pitch/yaw -> rotation matrix (R = Rz(y) Ry(p) Rx(r))."""
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _floats(text: Optional[str], n: int, default: float = 0.0) -> List[float]:
    if text is None:
        return [default] * n
    vals = [float(v) for v in text.split()]
    if len(vals) != n:
        raise ModelError(f'expected {n} numbers, got "{text}"')
    return vals


@dataclass
class _Shape:
    kind: int
    R: np.ndarray          # link frame <- shape frame
    p: np.ndarray
    dim: Tuple[float, float, float]
    verts: Optional[np.ndarray] = None     # SHAPE_HULL: [n][3] in the shape frame


@dataclass
class _Link:
    name: str
    mass: float = 0.0
    com: np.ndarray = field(default_factory=lambda: np.zeros(3))
    R_inertial: np.ndarray = field(default_factory=lambda: np.eye(3))
    inertia: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))
    shapes: List[_Shape] = field(default_factory=list)


@dataclass
class _Joint:
    name: str
    jtype: int
    parent: str
    child: str
    R: np.ndarray          # parent link frame <- joint (= child link) frame
    p: np.ndarray
    axis: np.ndarray       # in the child link frame
    lower: float = 0.0
    upper: float = -1.0
    damping: float = 0.0


@dataclass
class RobotModel:
    """Flat, Bullet-style description (link frames at the COM, principal axes)."""
    nl: int
    parent: np.ndarray
    jtype: np.ndarray
    E0: np.ndarray
    e: np.ndarray
    d: np.ndarray
    axis: np.ndarray
    mass: np.ndarray
    inertia: np.ndarray
    damping: np.ndarray
    lower: np.ndarray
    upper: np.ndarray
    has_limit: np.ndarray
    base_R: np.ndarray
    base_p: np.ndarray
    s_link: np.ndarray
    s_type: np.ndarray
    s_R: np.ndarray
    s_p: np.ndarray
    s_dim: np.ndarray
    joint_names: List[str]
    link_names: List[str]
    joint_axis_link: np.ndarray     # axis in the child link frame (what getJointInfo()[13] reports)
    # physics-engine defaults PyBullet applies
    lin_damp: float = 0.04
    ang_damp: float = 0.04
    gravity: Tuple[float, float, float] = (0.0, 0.0, -9.81)
    dt: float = 1.0 / 240.0
    iters: int = 50
    resid_thresh: float = 1e-7
    erp: float = 0.2
    max_vel: float = 100.0
    limit_max_impulse: float = 100.0
    # convex vertex clouds of the SHAPE_HULL shapes: shape s owns verts[s_v0[s] : s_v0[s] + s_vn[s]]
    s_v0: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    s_vn: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    verts: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))

    @property
    def ns(self) -> int:
        return int(self.s_link.shape[0])

    @property
    def ndof(self) -> int:
        return int((self.jtype != FIXED).sum())


def _principal(I: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Diagonalise a 3x3 inertia tensor; returns (principal moments, R with I = R diag R^T)."""
    off = abs(I[0, 1]) + abs(I[0, 2]) + abs(I[1, 2])
    if off < 1e-12 * max(1.0, np.trace(I)):
        return np.diag(I).copy(), np.eye(3)
    w, V = np.linalg.eigh(I)
    if np.linalg.det(V) < 0:
        V[:, 2] = -V[:, 2]
    return w, V


def _compile(links: dict, joints: List[_Joint], base_R: np.ndarray, base_p: np.ndarray) -> RobotModel:
    children = {}
    child_names = set()
    for j in joints:
        children.setdefault(j.parent, []).append(j)
        child_names.add(j.child)
    roots = [n for n in links if n not in child_names]
    if len(roots) != 1:
        raise ModelError(f'expected exactly one root link, found {roots}')
    root = roots[0]

    order: List[Tuple[_Joint, int]] = []

    def visit(link_name: str, parent_index: int) -> None:   # URDF2Bullet depth-first numbering
        for j in children.get(link_name, []):
            order.append((j, parent_index))
            visit(j.child, len(order) - 1)

    visit(root, -1)
    nl = len(order)
    if nl == 0 or nl > MAX_LINKS:
        raise ModelError(f'unsupported number of joints: {nl} (1..{MAX_LINKS})')

    # COM frames: link frame -> principal inertial frame
    def com_frame(l: _Link) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        if l.mass <= 0 and not np.any(l.inertia):
            return np.eye(3), np.zeros(3), np.zeros(3)
        I_link = l.R_inertial @ l.inertia @ l.R_inertial.T
        w, V = _principal(I_link)
        return V, l.com.copy(), w

    frames = {n: com_frame(l) for n, l in links.items()}

    m = dict(parent=np.zeros(nl, np.int32), jtype=np.zeros(nl, np.int32), E0=np.zeros((nl, 9)),
             e=np.zeros((nl, 3)), d=np.zeros((nl, 3)), axis=np.zeros((nl, 3)), mass=np.zeros(nl),
             inertia=np.zeros((nl, 3)), damping=np.zeros(nl), lower=np.zeros(nl), upper=np.zeros(nl),
             has_limit=np.zeros(nl, np.int32))
    axis_link = np.zeros((nl, 3))
    s_link, s_type, s_R, s_p, s_dim = [], [], [], [], []
    s_v0, s_vn, verts = [], [], []
    nverts = 0

    def add_shapes(link_name: str, index: int) -> None:
        nonlocal nverts
        R_lc, c, _ = frames[link_name]
        for sh in links[link_name].shapes:
            s_link.append(index)
            s_type.append(sh.kind)
            s_R.append((R_lc.T @ sh.R).reshape(9))
            s_p.append(R_lc.T @ (sh.p - c))
            s_dim.append(sh.dim)
            n = 0 if sh.verts is None else int(sh.verts.shape[0])
            s_v0.append(nverts)
            s_vn.append(n)
            if n:
                verts.append(np.asarray(sh.verts, float))
                nverts += n

    Rb_lc, cb, _ = frames[root]
    base_R_com = base_R @ Rb_lc
    base_p_com = base_p + base_R @ cb

    for i, (j, pidx) in enumerate(order):
        Rp_lc, cp, _ = frames[j.parent]
        Rc_lc, cc, Ic = frames[j.child]
        m['parent'][i] = pidx
        m['jtype'][i] = j.jtype
        m['E0'][i] = (Rc_lc.T @ j.R.T @ Rp_lc).reshape(9)
        m['e'][i] = Rp_lc.T @ (j.p - cp)
        m['d'][i] = Rc_lc.T @ cc
        ax = np.asarray(j.axis, float)
        n = np.linalg.norm(ax)
        ax = ax / n if n > 0 else np.array([0.0, 0.0, 1.0])
        axis_link[i] = ax
        m['axis'][i] = Rc_lc.T @ ax
        m['mass'][i] = links[j.child].mass
        m['inertia'][i] = Ic
        m['damping'][i] = j.damping
        m['lower'][i], m['upper'][i] = j.lower, j.upper
        m['has_limit'][i] = 1 if (j.jtype != FIXED and j.lower <= j.upper) else 0
        add_shapes(j.child, i)
        if j.jtype != FIXED and links[j.child].mass <= 0:
            raise ModelError(f'movable link {j.child} has no mass')

    ns = len(s_link)
    if ns > MAX_SHAPES:
        raise ModelError(f'too many collision shapes: {ns} > {MAX_SHAPES}')
    if nverts > MAX_VERTS:
        raise ModelError(f'too many convex-hull vertices: {nverts} > {MAX_VERTS}; simplify the collision meshes')
    return RobotModel(
        s_v0=np.asarray(s_v0, np.int32).reshape(ns), s_vn=np.asarray(s_vn, np.int32).reshape(ns),
        verts=np.concatenate(verts, axis=0) if verts else np.zeros((0, 3)),
        nl=nl, base_R=base_R_com.reshape(9), base_p=base_p_com,
        s_link=np.asarray(s_link, np.int32).reshape(ns), s_type=np.asarray(s_type, np.int32).reshape(ns),
        s_R=np.asarray(s_R, float).reshape(ns, 9), s_p=np.asarray(s_p, float).reshape(ns, 3),
        s_dim=np.asarray(s_dim, float).reshape(ns, 3),
        joint_names=[j.name for j, _ in order], link_names=[j.child for j, _ in order],
        joint_axis_link=axis_link, **m)


# ----------------------------------------------------------------------------------------- geometry
def _parse_geometry(geom: ET.Element, R: np.ndarray, p: np.ndarray, sdf: bool, model_dir: str) -> Optional[_Shape]:
    if geom is None:
        return None
    node = geom.find('sphere')
    if node is not None:
        r = float(node.findtext('radius')) if sdf else float(node.get('radius'))
        return _Shape(SHAPE_SPHERE, R, p, (r, 0.0, 0.0))
    node = geom.find('capsule')
    if node is not None:
        r = float(node.findtext('radius')) if sdf else float(node.get('radius'))
        ln = float(node.findtext('length')) if sdf else float(node.get('length'))
        return _Shape(SHAPE_CAPSULE, R, p, (r, 0.5 * ln, 0.0))
    node = geom.find('box')
    if node is not None:
        size = _floats(node.findtext('size') if sdf else node.get('size'), 3)
        return _Shape(SHAPE_BOX, R, p, (0.5 * size[0], 0.5 * size[1], 0.5 * size[2]))
    node = geom.find('cylinder')
    if node is not None:                               # imported as the hull of 2 x 32 rim points, like pybullet
        r = float(node.findtext('radius')) if sdf else float(node.get('radius'))
        ln = float(node.findtext('length')) if sdf else float(node.get('length'))
        return _Shape(SHAPE_HULL, R, p, (mesh_io.HULL_MARGIN, 0.0, 0.0), mesh_io.cylinder_vertex_cloud(r, ln))
    node = geom.find('mesh')
    if node is not None:
        uri = node.findtext('uri') if sdf else node.get('filename')
        scale = _floats((node.findtext('scale') if sdf else node.get('scale')) or '1 1 1', 3)
        if not uri:
            raise ModelError('<mesh> collision element without a file name')
        try:
            v = mesh_io.load_mesh_vertices(mesh_io.resolve_mesh_file(uri.strip(), model_dir, [d for d in (_pybullet_data_path(), DATA_PATH) if d]))
            v = mesh_io.convex_vertex_cloud(v * np.asarray(scale))
        except mesh_io.MeshError as err:
            raise ModelError(str(err))
        return _Shape(SHAPE_HULL, R, p, (mesh_io.HULL_MARGIN, 0.0, 0.0), v)
    return None


_JTYPES = {'revolute': REVOLUTE, 'continuous': REVOLUTE, 'prismatic': PRISMATIC, 'fixed': FIXED}


def load_urdf(path: str) -> RobotModel:
    try:
        root = ET.parse(path).getroot()
    except (ET.ParseError, OSError) as err:
        raise ModelError(str(err))
    if root.tag != 'robot':
        raise ModelError('not a URDF file (no <robot> root)')
    links, joints = {}, []
    for ln in root.findall('link'):
        l = _Link(ln.get('name'))
        inert = ln.find('inertial')
        if inert is not None:
            org = inert.find('origin')
            if org is not None:
                l.com = np.array(_floats(org.get('xyz'), 3))
                l.R_inertial = rpy_to_R(*_floats(org.get('rpy'), 3))
            l.mass = float(inert.find('mass').get('value')) if inert.find('mass') is not None else 0.0
            it = inert.find('inertia')
            if it is not None:
                g = lambda k: float(it.get(k, 0.0))
                l.inertia = np.array([[g('ixx'), g('ixy'), g('ixz')], [g('ixy'), g('iyy'), g('iyz')],
                                      [g('ixz'), g('iyz'), g('izz')]])
        for col in ln.findall('collision'):
            org = col.find('origin')
            R, p = np.eye(3), np.zeros(3)
            if org is not None:
                p = np.array(_floats(org.get('xyz'), 3))
                R = rpy_to_R(*_floats(org.get('rpy'), 3))
            sh = _parse_geometry(col.find('geometry'), R, p, False, os.path.dirname(os.path.abspath(path)))
            if sh is not None:
                l.shapes.append(sh)
        links[l.name] = l
    for jn in root.findall('joint'):
        t = jn.get('type')
        if t not in _JTYPES:
            raise ModelError(f'unsupported joint type {t}')
        org = jn.find('origin')
        R, p = np.eye(3), np.zeros(3)
        if org is not None:
            p = np.array(_floats(org.get('xyz'), 3))
            R = rpy_to_R(*_floats(org.get('rpy'), 3))
        ax = jn.find('axis')
        axis = np.array(_floats(ax.get('xyz'), 3)) if ax is not None else np.array([1.0, 0.0, 0.0])
        j = _Joint(jn.get('name'), _JTYPES[t], jn.find('parent').get('link'), jn.find('child').get('link'),
                   R, p, axis)
        lim = jn.find('limit')
        if lim is not None and t != 'continuous':
            j.lower, j.upper = float(lim.get('lower', 0.0)), float(lim.get('upper', 0.0))
        dyn = jn.find('dynamics')
        if dyn is not None:
            j.damping = float(dyn.get('damping', 0.0))
        joints.append(j)
    children = {j.child for j in joints}
    for name, l in links.items():
        if name not in children and l.mass > 0:
            # the reference calls p.loadURDF(file) without useFixedBase (environment.py:229): a root link with mass
            # would be a free-floating base there and fall under gravity; this simulator keeps every base fixed
            logger.warning(f'root link "{name}" has mass {l.mass}: PyBullet would treat it as a free-floating base '
                           f'(loadURDF without useFixedBase); the B200 simulator keeps the base fixed')
    return _compile(links, joints, np.eye(3), np.zeros(3))


def load_sdf(path: str) -> RobotModel:
    """First <model> of an SDF file (the reference keeps ``p.loadSDF(file)[0]``, environment.py:231)."""
    try:
        root = ET.parse(path).getroot()
    except (ET.ParseError, OSError) as err:
        raise ModelError(str(err))
    model = root.find('.//model')
    if root.tag != 'sdf' or model is None:
        raise ModelError('not an SDF file (no <sdf>/<model>)')

    def pose(node: Optional[ET.Element]) -> Tuple[np.ndarray, np.ndarray]:
        v = _floats(node.text if node is not None else None, 6)
        return rpy_to_R(v[3], v[4], v[5]), np.array(v[:3])

    Rm, pm = pose(model.find('pose'))
    links, world = {}, {}
    for ln in model.findall('link'):
        l = _Link(ln.get('name'))
        Rl, pl = pose(ln.find('pose'))
        world[l.name] = (Rm @ Rl, pm + Rm @ pl)
        inert = ln.find('inertial')
        if inert is not None:
            Ri, pi = pose(inert.find('pose'))
            l.com, l.R_inertial = pi, Ri
            l.mass = float(inert.findtext('mass', '0'))
            it = inert.find('inertia')
            if it is not None:
                g = lambda k: float(it.findtext(k, '0'))
                l.inertia = np.array([[g('ixx'), g('ixy'), g('ixz')], [g('ixy'), g('iyy'), g('iyz')],
                                      [g('ixz'), g('iyz'), g('izz')]])
        for col in ln.findall('collision'):
            Rc, pc = pose(col.find('pose'))
            sh = _parse_geometry(col.find('geometry'), Rc, pc, True, os.path.dirname(os.path.abspath(path)))
            if sh is not None:
                l.shapes.append(sh)
        links[l.name] = l
    joints = []
    for jn in model.findall('joint'):
        t = jn.get('type')
        if t not in _JTYPES:
            raise ModelError(f'unsupported joint type {t}')
        parent, child = jn.findtext('parent'), jn.findtext('child')
        if parent == 'world':
            continue                                  # base is fixed to the world already
        Rp, pp = world[parent]
        Rc, pc = world[child]
        Rj, pj = pose(jn.find('pose'))                # joint frame relative to the child link
        if np.abs(pj).max() > 1e-12 or np.abs(Rj - np.eye(3)).max() > 1e-12:
            raise ModelError('SDF joints with a non-identity <pose> are not supported')
        R = Rp.T @ Rc
        p = Rp.T @ (pc - pp)
        axn = jn.find('axis')
        axis = np.array([0.0, 0.0, 1.0])
        lower, upper, damping = 0.0, -1.0, 0.0
        if axn is not None:
            axis = np.array(_floats(axn.findtext('xyz'), 3))
            if axn.findtext('use_parent_model_frame', '0').strip() in ('1', 'true'):
                axis = Rc.T @ (Rm @ axis)
            lim = axn.find('limit')
            if lim is not None and t != 'continuous':
                lower, upper = float(lim.findtext('lower', '0')), float(lim.findtext('upper', '0'))
            dyn = axn.find('dynamics')
            if dyn is not None:
                damping = float(dyn.findtext('damping', '0'))
        joints.append(_Joint(jn.get('name'), _JTYPES[t], parent, child, R, p, axis, lower, upper, damping))
    child_names = {j.child for j in joints}
    roots = [n for n in links if n not in child_names]
    if len(roots) != 1:
        raise ModelError(f'expected exactly one root link, found {roots}')
    Rb, pb = world[roots[0]]
    return _compile(links, joints, Rb, pb)


def _pybullet_data_path() -> Optional[str]:
    """The real asset directory when the user's machine has it (the reference adds it to PyBullet's search path,
    environment.py:210); RLOA_ASSETS=standin forces the models shipped with this package."""
    if os.environ.get('RLOA_ASSETS', '').lower() == 'standin':
        return None
    try:
        import pybullet_data
        return pybullet_data.getDataPath()
    except Exception:       # not installed (the build image), or a broken install
        return None


def resolve_manipulator_file(manipulator_file: str) -> str:
    """File lookup: as given, else relative to pybullet_data when it is installed, else relative to the package data
    dir (the stand-in assets)."""
    if os.path.isfile(manipulator_file):
        return manipulator_file
    roots = [d for d in (_pybullet_data_path(), DATA_PATH) if d]
    for root in roots:
        cand = os.path.join(root, manipulator_file)
        if os.path.isfile(cand):
            if root is not DATA_PATH:
                logger.info(f'loading {manipulator_file} from pybullet_data ({root})')
            return cand
    # a path that ends in a known relative name, e.g. <somewhere>/pybullet_data/kuka_iiwa/x.sdf
    parts = manipulator_file.replace('\\', '/').split('/')
    for root in roots:
        for k in range(len(parts) - 1, 0, -1):
            cand = os.path.join(root, *parts[k - 1:])
            if os.path.isfile(cand):
                return cand
    raise ModelError(f'file not found: {manipulator_file}')


def _load_file(path: str) -> RobotModel:
    if path.endswith('.urdf'):
        return load_urdf(path)
    if path.endswith('.sdf'):
        return load_sdf(path)
    raise ModelError('The file extension is neither .sdf nor .urdf')


def load_manipulator(manipulator_file: str) -> RobotModel:
    path = resolve_manipulator_file(manipulator_file)
    try:
        return _load_file(path)
    except ModelError as err:
        # a pybullet_data file this front end cannot digest: fall back to the shipped stand-in of the same name, loudly
        standin = os.path.join(DATA_PATH, manipulator_file)
        if not path.startswith(DATA_PATH) and not os.path.isfile(manipulator_file) and os.path.isfile(standin):
            logger.warning(f'{path} could not be loaded ({err}); using the stand-in model {standin} instead')
            return _load_file(standin)
        raise
