"""
Generates the stand-in robot assets shipped under robotic_manipulator_rloa_b200/data/.

pybullet_data (kuka_iiwa/kuka_with_gripper2.sdf, franka_panda/panda.urdf, their meshes) is NOT
available in this image, so these files are written from the kinematic /
inertial parameters recalled in SURVEY.md Appendix C, with primitive collision geometry in place of the
convex-hulled meshes.  They keep pybullet_data's relative file names so the reference's demo calls
(rl_framework.py:547, 641) resolve, and they are clearly labelled as stand-ins inside the files.

Run:  python tools/make_standin_assets.py
"""
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(os.path.dirname(HERE), 'robotic_manipulator_rloa_b200', 'data')
PI = math.pi


def rpy_to_R(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def R_to_rpy(R):
    sp = -R[2, 0]
    p = math.asin(max(-1.0, min(1.0, sp)))
    if abs(sp) < 1 - 1e-12:
        return math.atan2(R[2, 1], R[2, 2]), p, math.atan2(R[1, 0], R[0, 0])
    return 0.0, p, math.atan2(-R[0, 1], R[1, 1])


# name, parent, joint name, type, origin xyz, origin rpy, axis, (lower, upper)|None, damping,
# mass, com, inertia diag, collisions [(kind, xyz, rpy, dims)]
KUKA = [
    ('lbr_iiwa_link_1', 'lbr_iiwa_link_0', 'J0', 'revolute', (0, 0, 0.1575), (0, 0, 0), (0, 0, 1), (-2.96706, 2.96706), 0.5,
     4.0, (0, -0.03, 0.12), (0.1, 0.09, 0.02), [('capsule', (0, 0, 0.10), (0, 0, 0), (0.07, 0.14))]),
    ('lbr_iiwa_link_2', 'lbr_iiwa_link_1', 'J1', 'revolute', (0, 0, 0.2025), (PI / 2, 0, PI), (0, 0, 1), (-2.09440, 2.09440), 0.5,
     4.0, (0.0003, 0.059, 0.042), (0.05, 0.018, 0.044), [('capsule', (0, 0.10, 0), (PI / 2, 0, 0), (0.07, 0.14))]),
    ('lbr_iiwa_link_3', 'lbr_iiwa_link_2', 'J2', 'revolute', (0, 0.2045, 0), (PI / 2, 0, PI), (0, 0, 1), (-2.96706, 2.96706), 0.5,
     3.0, (0, 0.03, 0.13), (0.08, 0.075, 0.01), [('capsule', (0, 0, 0.11), (0, 0, 0), (0.065, 0.14))]),
    ('lbr_iiwa_link_4', 'lbr_iiwa_link_3', 'J3', 'revolute', (0, 0, 0.2155), (PI / 2, 0, 0), (0, 0, 1), (-2.09440, 2.09440), 0.5,
     2.7, (0, 0.067, 0.034), (0.03, 0.01, 0.029), [('capsule', (0, 0.09, 0), (PI / 2, 0, 0), (0.065, 0.12))]),
    ('lbr_iiwa_link_5', 'lbr_iiwa_link_4', 'J4', 'revolute', (0, 0.1845, 0), (-PI / 2, PI, 0), (0, 0, 1), (-2.96706, 2.96706), 0.5,
     1.7, (0.0001, 0.021, 0.076), (0.02, 0.018, 0.005), [('capsule', (0, 0, 0.11), (0, 0, 0), (0.06, 0.14))]),
    ('lbr_iiwa_link_6', 'lbr_iiwa_link_5', 'J5', 'revolute', (0, 0, 0.2155), (PI / 2, 0, 0), (0, 0, 1), (-2.09440, 2.09440), 0.5,
     1.8, (0, 0.0006, 0.0004), (0.005, 0.0036, 0.0047), [('capsule', (0, 0.03, 0), (PI / 2, 0, 0), (0.06, 0.06))]),
    ('lbr_iiwa_link_7', 'lbr_iiwa_link_6', 'J6', 'revolute', (0, 0.081, 0), (-PI / 2, PI, 0), (0, 0, 1), (-3.05433, 3.05433), 0.5,
     0.3, (0, 0, 0.02), (0.001, 0.001, 0.001), [('sphere', (0, 0, 0.02), (0, 0, 0), (0.045,))]),
    ('base_link', 'lbr_iiwa_link_7', 'gripper_to_arm', 'continuous', (0, 0, 0.044), (0, 0, 0), (0, 0, 1), None, 0.0,
     0.2, (0, 0, 0), (2e-4, 2e-4, 2e-4), [('box', (0, 0, 0.01), (0, 0, 0), (0.05, 0.10, 0.04))]),
    ('left_finger', 'base_link', 'base_left_finger_joint', 'revolute', (0, 0.024, 0.045), (0, 0, 0), (1, 0, 0), (-0.6, 0.6), 0.0,
     0.2, (0, 0, 0.04), (1.2e-4, 1.2e-4, 2e-5), [('capsule', (0, 0, 0.04), (0, 0, 0), (0.01, 0.06))]),
    ('left_finger_base', 'left_finger', 'left_finger_base_joint', 'fixed', (0, 0, 0.08), (0, 0, 0), (0, 0, 1), None, 0.0,
     0.2, (0, 0, 0.02), (5e-5, 5e-5, 2e-5), [('capsule', (0, 0, 0.03), (0, 0, 0), (0.01, 0.04))]),
    ('left_finger_tip', 'left_finger_base', 'left_base_tip_joint', 'revolute', (0, 0, 0.06), (0, 0, 0), (1, 0, 0), (-0.6, 0.6), 0.0,
     0.2, (0, 0, 0.02), (5e-5, 5e-5, 2e-5), [('capsule', (0, 0, 0.02), (0, 0, 0), (0.01, 0.03))]),
    ('right_finger', 'base_link', 'base_right_finger_joint', 'revolute', (0, -0.024, 0.045), (0, 0, 0), (1, 0, 0), (-0.6, 0.6), 0.0,
     0.2, (0, 0, 0.04), (1.2e-4, 1.2e-4, 2e-5), [('capsule', (0, 0, 0.04), (0, 0, 0), (0.01, 0.06))]),
    ('right_finger_base', 'right_finger', 'right_finger_base_joint', 'fixed', (0, 0, 0.08), (0, 0, 0), (0, 0, 1), None, 0.0,
     0.2, (0, 0, 0.02), (5e-5, 5e-5, 2e-5), [('capsule', (0, 0, 0.03), (0, 0, 0), (0.01, 0.04))]),
    ('right_finger_tip', 'right_finger_base', 'right_base_tip_joint', 'revolute', (0, 0, 0.06), (0, 0, 0), (1, 0, 0), (-0.6, 0.6), 0.0,
     0.2, (0, 0, 0.02), (5e-5, 5e-5, 2e-5), [('capsule', (0, 0, 0.02), (0, 0, 0), (0.01, 0.03))]),
]

PANDA = [
    ('panda_link1', 'panda_link0', 'panda_joint1', 'revolute', (0, 0, 0.333), (0, 0, 0), (0, 0, 1), (-2.9671, 2.9671), 0.5,
     2.74, (0, -0.04, -0.05), (0.3, 0.3, 0.3), [('capsule', (0, 0, -0.10), (0, 0, 0), (0.07, 0.16))]),
    ('panda_link2', 'panda_link1', 'panda_joint2', 'revolute', (0, 0, 0), (-PI / 2, 0, 0), (0, 0, 1), (-1.8326, 1.8326), 0.5,
     2.74, (0, -0.04, 0.06), (0.3, 0.3, 0.3), [('capsule', (0, -0.08, 0), (PI / 2, 0, 0), (0.07, 0.12))]),
    ('panda_link3', 'panda_link2', 'panda_joint3', 'revolute', (0, -0.316, 0), (PI / 2, 0, 0), (0, 0, 1), (-2.9671, 2.9671), 0.5,
     2.38, (0.04, 0.03, -0.03), (0.3, 0.3, 0.3), [('capsule', (0, 0, -0.10), (0, 0, 0), (0.065, 0.14))]),
    ('panda_link4', 'panda_link3', 'panda_joint4', 'revolute', (0.0825, 0, 0), (PI / 2, 0, 0), (0, 0, 1), (-3.0718, -0.0698), 0.5,
     2.38, (-0.04, 0.04, 0.0), (0.3, 0.3, 0.3), [('capsule', (-0.04, 0.04, 0), (PI / 2, 0, 0), (0.065, 0.10))]),
    ('panda_link5', 'panda_link4', 'panda_joint5', 'revolute', (-0.0825, 0.384, 0), (-PI / 2, 0, 0), (0, 0, 1), (-2.9671, 2.9671), 0.5,
     2.74, (0, 0.04, -0.12), (0.3, 0.3, 0.3), [('capsule', (0, 0.02, -0.16), (0, 0, 0), (0.06, 0.20))]),
    ('panda_link6', 'panda_link5', 'panda_joint6', 'revolute', (0, 0, 0), (PI / 2, 0, 0), (0, 0, 1), (-0.0873, 3.8223), 0.5,
     1.55, (0.06, -0.01, 0.0), (0.1, 0.1, 0.1), [('capsule', (0.04, 0, 0), (0, PI / 2, 0), (0.06, 0.08))]),
    ('panda_link7', 'panda_link6', 'panda_joint7', 'revolute', (0.088, 0, 0), (PI / 2, 0, 0), (0, 0, 1), (-2.9671, 2.9671), 0.5,
     0.54, (0, 0, 0.08), (0.05, 0.05, 0.05), [('sphere', (0, 0, 0.07), (0, 0, 0), (0.05,))]),
    ('panda_link8', 'panda_link7', 'panda_joint8', 'fixed', (0, 0, 0.107), (0, 0, 0), (0, 0, 1), None, 0.0,
     0.0, (0, 0, 0), (0, 0, 0), []),
    ('panda_hand', 'panda_link8', 'panda_hand_joint', 'fixed', (0, 0, 0), (0, 0, -PI / 4), (0, 0, 1), None, 0.0,
     0.73, (0, 0, 0.04), (0.01, 0.01, 0.01), [('box', (0, 0, 0.03), (0, 0, 0), (0.04, 0.20, 0.07))]),
    ('panda_leftfinger', 'panda_hand', 'panda_finger_joint1', 'prismatic', (0, 0, 0.0584), (0, 0, 0), (0, 1, 0), (0.0, 0.04), 0.0,
     0.1, (0, 0.01, 0.02), (1e-4, 1e-4, 1e-4), [('capsule', (0, 0.01, 0.025), (0, 0, 0), (0.008, 0.04))]),
    ('panda_rightfinger', 'panda_hand', 'panda_finger_joint2', 'prismatic', (0, 0, 0.0584), (0, 0, 0), (0, -1, 0), (0.0, 0.04), 0.0,
     0.1, (0, -0.01, 0.02), (1e-4, 1e-4, 1e-4), [('capsule', (0, -0.01, 0.025), (0, 0, 0), (0.008, 0.04))]),
    ('panda_grasptarget', 'panda_hand', 'panda_grasptarget_hand', 'fixed', (0, 0, 0.105), (0, 0, 0), (0, 0, 1), None, 0.0,
     0.0, (0, 0, 0), (0, 0, 0), [('sphere', (0, 0, 0), (0, 0, 0), (0.01,))]),
]

# UFACTORY xArm 6 + gripper, joint order of pybullet_data/xarm/xarm6_with_gripper.urdf as the reference's demo uses it
# (rl_framework.py:571-580: endeffector 12, involved joints 1-6, fixed [0, 7..13]): joint 0 is the fixed world joint,
# 1-6 the arm, 7 the fixed gripper mount, 8-13 the six revolute gripper joints.  Arm kinematics / masses from the
# xArm 6 description (medium confidence), gripper low confidence.
XARM6 = [
    ('link_base', 'world', 'world_joint', 'fixed', (0, 0, 0), (0, 0, 0), (0, 0, 1), None, 0.0,
     2.7, (0, 0, 0.07), (0.005, 0.005, 0.003), [('capsule', (0, 0, 0.07), (0, 0, 0), (0.06, 0.08))]),
    ('link1', 'link_base', 'joint1', 'revolute', (0, 0, 0.267), (0, 0, 0), (0, 0, 1), (-6.28318, 6.28318), 1.0,
     2.16, (0.0002, 0.0270, -0.0135), (0.0054, 0.0049, 0.0032), [('capsule', (0, 0, -0.06), (0, 0, 0), (0.055, 0.12))]),
    ('link2', 'link1', 'joint2', 'revolute', (0, 0, 0), (-PI / 2, 0, 0), (0, 0, 1), (-2.059, 2.0944), 1.0,
     1.71, (0.0367, -0.2209, 0.0335), (0.0271, 0.0041, 0.0263), [('capsule', (0.03, -0.14, 0.02), (PI / 2, 0, 0), (0.05, 0.26))]),
    ('link3', 'link2', 'joint3', 'revolute', (0.0535, -0.2845, 0), (0, 0, 0), (0, 0, 1), (-3.927, 0.19198), 1.0,
     1.384, (0.0680, 0.2278, 0.0108), (0.0064, 0.0016, 0.0067), [('capsule', (0.07, 0.12, 0), (PI / 2, 0, 0), (0.045, 0.16))]),
    ('link4', 'link3', 'joint4', 'revolute', (0.0775, 0.3425, 0), (-PI / 2, 0, 0), (0, 0, 1), (-6.28318, 6.28318), 1.0,
     1.115, (-0.0002, 0.0205, -0.0264), (0.0046, 0.0043, 0.0012), [('capsule', (0, 0, -0.10), (0, 0, 0), (0.04, 0.16))]),
    ('link5', 'link4', 'joint5', 'revolute', (0, 0, 0), (PI / 2, 0, 0), (0, 0, 1), (-1.69297, 3.14159), 1.0,
     1.275, (0.0646, 0.0290, 0.0064), (0.0014, 0.0023, 0.0029), [('capsule', (0.05, 0.04, 0), (0, PI / 2, 0), (0.04, 0.08))]),
    ('link6', 'link5', 'joint6', 'revolute', (0.076, 0.097, 0), (-PI / 2, 0, 0), (0, 0, 1), (-6.28318, 6.28318), 1.0,
     0.1096, (0, -0.0025, -0.0166), (5e-5, 5e-5, 8e-5), [('sphere', (0, 0, -0.01), (0, 0, 0), (0.04,))]),
    ('xarm_gripper_base_link', 'link6', 'gripper_fix', 'fixed', (0, 0, 0), (0, 0, 0), (0, 0, 1), None, 0.0,
     0.54, (-0.0003, 0.0002, 0.0538), (4.7e-4, 3.0e-4, 3.7e-4), [('box', (0, 0, 0.045), (0, 0, 0), (0.06, 0.09, 0.09))]),
    ('left_outer_knuckle', 'xarm_gripper_base_link', 'drive_joint', 'revolute', (0, 0.035, 0.059098), (0, 0, 0), (1, 0, 0),
     (0.0, 0.85), 0.0, 0.033, (0, 0.021, 0.016), (1.9e-5, 6.7e-6, 1.3e-5), [('capsule', (0, 0.02, 0.015), (0, 0, 0), (0.008, 0.03))]),
    ('left_finger', 'left_outer_knuckle', 'left_finger_joint', 'revolute', (0, 0.035465, 0.042039), (0, 0, 0), (-1, 0, 0),
     (0.0, 0.85), 0.0, 0.048, (0, -0.016, 0.014), (2.1e-5, 1.7e-5, 1.1e-5), [('capsule', (0, -0.015, 0.02), (0, 0, 0), (0.008, 0.04))]),
    ('left_inner_knuckle', 'xarm_gripper_base_link', 'left_inner_knuckle_joint', 'revolute', (0, 0.02, 0.074098), (0, 0, 0),
     (1, 0, 0), (0.0, 0.85), 0.0, 0.033, (0, 0.017, 0.019), (1.3e-5, 8.7e-6, 8.3e-6), [('capsule', (0, 0.017, 0.019), (0, 0, 0), (0.007, 0.03))]),
    ('right_outer_knuckle', 'xarm_gripper_base_link', 'right_outer_knuckle_joint', 'revolute', (0, -0.035, 0.059098), (0, 0, 0),
     (-1, 0, 0), (0.0, 0.85), 0.0, 0.033, (0, -0.021, 0.016), (1.9e-5, 6.7e-6, 1.3e-5), [('capsule', (0, -0.02, 0.015), (0, 0, 0), (0.008, 0.03))]),
    ('right_finger', 'right_outer_knuckle', 'right_finger_joint', 'revolute', (0, -0.035465, 0.042039), (0, 0, 0), (1, 0, 0),
     (0.0, 0.85), 0.0, 0.048, (0, 0.016, 0.014), (2.1e-5, 1.7e-5, 1.1e-5), [('capsule', (0, 0.015, 0.02), (0, 0, 0), (0.008, 0.04))]),
    ('right_inner_knuckle', 'xarm_gripper_base_link', 'right_inner_knuckle_joint', 'revolute', (0, -0.02, 0.074098), (0, 0, 0),
     (-1, 0, 0), (0.0, 0.85), 0.0, 0.033, (0, -0.017, 0.019), (1.3e-5, 8.7e-6, 8.3e-6), [('capsule', (0, -0.017, 0.019), (0, 0, 0), (0.007, 0.03))]),
]

NOTE = ('STAND-IN ASSET written by tools/make_standin_assets.py: pybullet_data is not available in this image. '
        'Kinematics/inertials recalled from pybullet_data (SURVEY.md Appendix C), primitive collision shapes '
        'instead of convex-hulled meshes. Replace with the real file when available.')


def geom_sdf(kind, dims):
    if kind == 'sphere':
        return f'<sphere><radius>{dims[0]}</radius></sphere>'
    if kind == 'capsule':
        return f'<capsule><radius>{dims[0]}</radius><length>{dims[1]}</length></capsule>'
    return f'<box><size>{dims[0]} {dims[1]} {dims[2]}</size></box>'


def geom_urdf(kind, dims):
    if kind == 'sphere':
        return f'<sphere radius="{dims[0]}"/>'
    if kind == 'capsule':
        return f'<capsule radius="{dims[0]}" length="{dims[1]}"/>'
    return f'<box size="{dims[0]} {dims[1]} {dims[2]}"/>'


def fmt(v):
    return ' '.join(f'{0.0 if abs(x) < 5e-17 else x:.12g}' for x in v)


def write_sdf(path, model_name, root, table):
    world = {root: (np.eye(3), np.zeros(3))}
    out = [f'<?xml version="1.0"?>\n<!-- {NOTE} -->\n<sdf version="1.6">\n  <model name="{model_name}">\n'
           f'    <pose>0 0 0 0 0 0</pose>\n'
           f'    <link name="{root}">\n      <pose>0 0 0 0 0 0</pose>\n      <inertial><mass>0</mass></inertial>\n    </link>\n']
    joints = []
    for (name, parent, jname, jtype, xyz, rpy, axis, lim, damp, mass, com, inert, cols) in table:
        Rp, pp = world[parent]
        R = Rp @ rpy_to_R(*rpy)
        p = pp + Rp @ np.array(xyz, float)
        world[name] = (R, p)
        out.append(f'    <link name="{name}">\n      <pose>{fmt(p)} {fmt(R_to_rpy(R))}</pose>\n'
                   f'      <inertial>\n        <pose>{fmt(com)} 0 0 0</pose>\n        <mass>{mass}</mass>\n'
                   f'        <inertia><ixx>{inert[0]}</ixx><ixy>0</ixy><ixz>0</ixz><iyy>{inert[1]}</iyy><iyz>0</iyz>'
                   f'<izz>{inert[2]}</izz></inertia>\n      </inertial>\n')
        for k, (kind, cxyz, crpy, dims) in enumerate(cols):
            out.append(f'      <collision name="{name}_collision_{k}">\n        <pose>{fmt(cxyz)} {fmt(crpy)}</pose>\n'
                       f'        <geometry>{geom_sdf(kind, dims)}</geometry>\n      </collision>\n')
        out.append('    </link>\n')
        j = [f'    <joint name="{jname}" type="{jtype}">\n      <parent>{parent}</parent>\n      <child>{name}</child>\n']
        if jtype != 'fixed':
            j.append(f'      <axis>\n        <xyz>{fmt(axis)}</xyz>\n')
            if lim is not None:
                j.append(f'        <limit><lower>{lim[0]}</lower><upper>{lim[1]}</upper><effort>300</effort>'
                         f'<velocity>10</velocity></limit>\n')
            j.append(f'        <dynamics><damping>{damp}</damping></dynamics>\n      </axis>\n')
        j.append('    </joint>\n')
        joints.append(''.join(j))
    out.extend(joints)
    out.append('  </model>\n</sdf>\n')
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, 'w') as f:
        f.write(''.join(out))


def write_urdf(path, robot_name, root, table):
    out = [f'<?xml version="1.0"?>\n<!-- {NOTE} -->\n<robot name="{robot_name}">\n  <link name="{root}"/>\n']
    for (name, parent, jname, jtype, xyz, rpy, axis, lim, damp, mass, com, inert, cols) in table:
        out.append(f'  <link name="{name}">\n')
        if mass > 0:
            out.append(f'    <inertial>\n      <origin xyz="{fmt(com)}" rpy="0 0 0"/>\n      <mass value="{mass}"/>\n'
                       f'      <inertia ixx="{inert[0]}" ixy="0" ixz="0" iyy="{inert[1]}" iyz="0" izz="{inert[2]}"/>\n'
                       f'    </inertial>\n')
        for (kind, cxyz, crpy, dims) in cols:
            out.append(f'    <collision>\n      <origin xyz="{fmt(cxyz)}" rpy="{fmt(crpy)}"/>\n'
                       f'      <geometry>{geom_urdf(kind, dims)}</geometry>\n    </collision>\n')
        out.append('  </link>\n')
        out.append(f'  <joint name="{jname}" type="{jtype}">\n    <parent link="{parent}"/>\n    <child link="{name}"/>\n'
                   f'    <origin xyz="{fmt(xyz)}" rpy="{fmt(rpy)}"/>\n')
        if jtype != 'fixed':
            out.append(f'    <axis xyz="{fmt(axis)}"/>\n')
            if lim is not None:
                out.append(f'    <limit lower="{lim[0]}" upper="{lim[1]}" effort="300" velocity="10"/>\n')
            out.append(f'    <dynamics damping="{damp}"/>\n')
        out.append('  </joint>\n')
    out.append('</robot>\n')
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, 'w') as f:
        f.write(''.join(out))


if __name__ == '__main__':
    write_sdf(os.path.join(DATA, 'kuka_iiwa', 'kuka_with_gripper2.sdf'), 'lbr_iiwa_with_gripper_standin',
              'lbr_iiwa_link_0', KUKA)
    write_urdf(os.path.join(DATA, 'kuka_iiwa', 'model_standin.urdf'), 'lbr_iiwa_with_gripper_standin',
               'lbr_iiwa_link_0', KUKA)
    write_urdf(os.path.join(DATA, 'franka_panda', 'panda.urdf'), 'panda_standin', 'panda_link0', PANDA)
    write_urdf(os.path.join(DATA, 'xarm', 'xarm6_with_gripper.urdf'), 'xarm6_with_gripper_standin', 'world', XARM6)
    print('written under', DATA)
