"""
ctypes binding of librloa_b200.so (C ABI in include/rloa_b200.h).

The library is the product: there is no CPU fallback.  ``lib()`` raises ``NativeLibraryError`` when
the shared object is missing or does not export the ABI, and every call that returns a negative
status raises with ``rloa_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'librloa_b200.so')
MAX_LINKS = 32
MAX_SHAPES = 32
MAX_DOF = 16
MAX_HULL_VERTS = 16384


class NativeLibraryError(RuntimeError):
    """librloa_b200.so is missing, stale or a call into it failed."""


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_fp = C.c_void_p      # device pointers travel as integers (tensor.data_ptr())


class ModelDesc(C.Structure):
    _fields_ = [
        ('nl', C.c_int32), ('parent', _ip), ('jtype', _ip), ('E0', _dp), ('e', _dp), ('d', _dp), ('axis', _dp),
        ('mass', _dp), ('inertia', _dp), ('damping', _dp), ('lower', _dp), ('upper', _dp), ('has_limit', _ip),
        ('base_R', C.c_double * 9), ('base_p', C.c_double * 3), ('lin_damp', C.c_double), ('ang_damp', C.c_double),
        ('gravity', C.c_double * 3), ('dt', C.c_double), ('iters', C.c_int32), ('resid_thresh', C.c_double),
        ('erp', C.c_double), ('max_vel', C.c_double), ('limit_max_impulse', C.c_double),
        ('ns', C.c_int32), ('s_link', _ip), ('s_type', _ip), ('s_R', _dp), ('s_p', _dp), ('s_dim', _dp),
        ('obstacle_radius', C.c_double), ('target_half', C.c_double * 3), ('ee_link', C.c_int32),
        ('n_obs_joints', C.c_int32),
        ('n_verts', C.c_int32), ('s_vert_first', _ip), ('s_vert_count', _ip), ('verts', _dp),
    ]


class StepConfig(C.Structure):
    _fields_ = [('n_act', C.c_int32), ('act_joint', C.c_int32 * MAX_LINKS), ('n_fixed', C.c_int32),
                ('fixed_joint', C.c_int32 * MAX_LINKS), ('max_force', C.c_float), ('target_threshold', C.c_float),
                ('obstacle_threshold', C.c_float), ('contact_threshold', C.c_float)]


class NafParams(C.Structure):
    _fields_ = [('state_size', C.c_int32), ('action_size', C.c_int32), ('hidden', C.c_int32),
                ('w1', _fp), ('b1', _fp), ('bn1_w', _fp), ('bn1_b', _fp), ('bn1_mean', _fp), ('bn1_var', _fp),
                ('bn1_batches', _fp),
                ('w2', _fp), ('b2', _fp), ('bn2_w', _fp), ('bn2_b', _fp), ('bn2_mean', _fp), ('bn2_var', _fp),
                ('bn2_batches', _fp),
                ('w_mu', _fp), ('b_mu', _fp), ('w_v', _fp), ('b_v', _fp), ('w_l', _fp), ('b_l', _fp)]


class NafHyper(C.Structure):
    _fields_ = [('gamma', C.c_float), ('tau', C.c_float), ('lr', C.c_float), ('beta1', C.c_float),
                ('beta2', C.c_float), ('eps', C.c_float), ('clip_norm', C.c_float), ('trunc_action', C.c_int32),
                ('use_done_mask', C.c_int32), ('grad_scale', C.c_float)]


class AdamState(C.Structure):
    _fields_ = [('m', _fp), ('v', _fp), ('step', _fp)]


class Replay(C.Structure):
    _fields_ = [('capacity', C.c_int32), ('state_size', C.c_int32), ('action_size', C.c_int32),
                ('states', _fp), ('actions', _fp), ('rewards', _fp), ('next_states', _fp), ('dones', _fp),
                ('cursor', _fp), ('scratch', _fp)]


_VP = C.c_void_p
_I = C.c_int32
_U64 = C.c_uint64
_F = C.c_float

# name -> (restype, argtypes); the non-gpu test checks this table against include/rloa_b200.h
SIGNATURES = {
    'rloa_last_error': (C.c_char_p, []),
    'rloa_version': (C.c_int, []),
    'rloa_launch_count': (_U64, []),
    'rloa_model_create': (C.c_int, [C.POINTER(ModelDesc), C.POINTER(_VP)]),
    'rloa_model_destroy': (None, [_VP]),
    'rloa_sim_create': (C.c_int, [_VP, _I, C.POINTER(_VP)]),
    'rloa_sim_destroy': (None, [_VP]),
    'rloa_sim_num_envs': (C.c_int, [_VP]),
    'rloa_sim_obs_size': (C.c_int, [_VP]),
    'rloa_sim_set_task': (C.c_int, [_VP, _fp, _fp, _VP]),
    'rloa_sim_set_state': (C.c_int, [_VP, _fp, _fp, _VP]),
    'rloa_sim_get_state': (C.c_int, [_VP, _fp, _fp, _VP]),
    'rloa_sim_set_motors': (C.c_int, [_VP, _fp, _fp, _fp, _fp, _VP]),
    'rloa_sim_clear': (C.c_int, [_VP, _VP]),
    'rloa_sim_step': (C.c_int, [_VP, C.POINTER(StepConfig), _fp, _fp, _fp, _fp, _fp, _fp, _VP]),
    'rloa_sim_set_contacts': (C.c_int, [_VP, _F]),
    'rloa_sim_contact_counts': (C.c_int, [_VP, _VP, _VP]),
    'rloa_sim_prepare': (C.c_int, [_VP, _VP]),
    'rloa_sim_join': (C.c_int, [_VP, _VP]),
    'rloa_sim_begin_reset': (C.c_int, [_VP, _fp, _fp, _I, _I, _VP]),
    'rloa_sim_begin_reset_random': (C.c_int, [_VP, _fp, _fp, _fp, _I, _I, _U64, _fp, _VP]),
    'rloa_sim_reset': (C.c_int, [_VP, _fp, _fp, _I, _I, _fp, _VP]),
    'rloa_sim_observe': (C.c_int, [_VP, _fp, _fp, _fp, _VP]),
    'rloa_sim_self_distances': (C.c_int, [_VP, _fp, _VP]),
    'rloa_sim_last_iterations': (C.c_int, [_VP, _fp, _VP]),
    'rloa_episode_update': (C.c_int, [_I, _I, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _I, _fp, _fp, _fp, _VP]),
    'rloa_episode_update_reset': (C.c_int, [_VP, _I, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _I, _fp, _fp, _fp,
                                            _fp, _fp, _fp, _I, _I, _U64, _VP]),
    'rloa_naf_ws_create': (C.c_int, [_I, _I, _I, _I, C.POINTER(_VP)]),
    'rloa_naf_ws_destroy': (None, [_VP]),
    'rloa_naf_ws_set_trunk': (C.c_int, [_VP, _I]),
    'rloa_naf_hidden_layer': (C.c_int, [_VP, _fp, _fp, _fp, _fp, _fp, _fp, _I, _VP]),
    'rloa_umma_probe': (C.c_int, [_I, _fp, _fp, _fp, _VP]),
    'rloa_naf_ws_set_debug': (C.c_int, [_VP, _fp, _fp]),
    'rloa_naf_forward': (C.c_int, [_VP, C.POINTER(NafParams), _fp, _fp, _I, _I, _I, _fp, _fp, _fp, _fp, _VP]),
    'rloa_naf_act': (C.c_int, [_VP, C.POINTER(NafParams), _fp, _I, _U64, _U64, _fp, _F, _fp, _VP]),
    'rloa_naf_num_params': (C.c_int, [_I, _I, _I]),
    'rloa_naf_learn_grads': (C.c_int, [_VP, C.POINTER(NafParams), C.POINTER(NafParams), _fp, _fp, _fp, _fp, _fp, _I,
                                       C.POINTER(NafHyper), _fp, _fp, _VP]),
    'rloa_naf_learn_apply': (C.c_int, [_VP, C.POINTER(NafParams), C.POINTER(NafParams), C.POINTER(AdamState),
                                       C.POINTER(NafHyper), _fp, _fp, _VP]),
    'rloa_naf_learn_step': (C.c_int, [_VP, C.POINTER(NafParams), C.POINTER(NafParams), C.POINTER(AdamState), _fp, _fp, _fp,
                                      _fp, _fp, _I, C.POINTER(NafHyper), _fp, _fp, _fp, _VP]),
    'rloa_naf_learn_fused_supported': (C.c_int, [_VP, _I]),
    'rloa_naf_learn_step_replay': (C.c_int, [_VP, C.POINTER(NafParams), C.POINTER(NafParams), C.POINTER(AdamState), _VP,
                                             C.POINTER(Replay), _U64, _U64, _fp, _I, C.POINTER(NafHyper), _fp, _fp, _fp, _VP]),
    'rloa_naf_learn_step_pending': (C.c_int, [_VP, C.POINTER(NafParams), C.POINTER(NafParams), C.POINTER(AdamState), _VP,
                                              C.POINTER(Replay), _U64, _U64, _fp, _I, C.POINTER(NafHyper), _I, _fp, _fp, _fp, _fp,
                                              _fp, _fp, _fp, _fp, _fp, _VP]),
    'rloa_naf_learn_prepack': (C.c_int, [_VP, C.POINTER(NafParams), C.POINTER(NafParams), _VP]),
    'rloa_xchg_create': (C.c_int, [_I, C.POINTER(_VP)]),
    'rloa_xchg_handle': (C.c_int, [_VP, C.c_char_p]),
    'rloa_xchg_connect': (C.c_int, [_VP, _I, _I, C.c_char_p]),
    'rloa_xchg_status': (C.c_int, [_VP]),
    'rloa_xchg_destroy': (None, [_VP]),
    'rloa_naf_learn_apply_xchg': (C.c_int, [_VP, C.POINTER(NafParams), C.POINTER(NafParams), C.POINTER(AdamState),
                                            C.POINTER(NafHyper), _VP, _fp, _fp, _VP]),
    'rloa_naf_learn_step_xchg': (C.c_int, [_VP, C.POINTER(NafParams), C.POINTER(NafParams), C.POINTER(AdamState), _VP, _fp, _fp,
                                           _fp, _fp, _fp, _I, C.POINTER(NafHyper), _fp, _fp, _fp, _VP]),
    'rloa_naf_soft_update': (C.c_int, [C.POINTER(NafParams), C.POINTER(NafParams), _F, _VP]),
    'rloa_replay_append': (C.c_int, [C.POINTER(Replay), _I, _fp, _fp, _fp, _fp, _fp, _fp, _VP]),
    'rloa_replay_append_rows': (C.c_int, [C.POINTER(Replay), _I, _fp, _fp, _fp, _fp, _fp, _fp, _VP]),
    'rloa_replay_commit': (C.c_int, [C.POINTER(Replay), _I, _fp, _VP]),
    'rloa_replay_sample': (C.c_int, [C.POINTER(Replay), _I, _U64, _U64, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _VP]),
}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """The loaded library with typed entry points; raises NativeLibraryError when unusable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NativeLibraryError(
            f'{LIB_PATH} is missing: build it with `python -m robotic_manipulator_rloa_b200.build_native` '
            f'(nvcc, sm_100a). There is no CPU fallback for the simulator / NAF hot path.')
    try:
        handle = C.CDLL(LIB_PATH)
    except OSError as err:
        raise NativeLibraryError(f'cannot load {LIB_PATH}: {err}')
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError:
            raise NativeLibraryError(f'{LIB_PATH} does not export {name}; rebuild it')
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return handle


def check(status: int, what: str = '') -> None:
    if status < 0:
        msg = lib().rloa_last_error().decode('utf-8', 'replace')
        raise NativeLibraryError(f'{what or "librloa_b200"} failed ({status}): {msg}')


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None passes NULL)."""
    return None if t is None else t.data_ptr()


def current_stream_handle(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def make_model_desc(model, ee_link: int, n_obs_joints: int, obstacle_radius: float = 0.075,
                    target_half=(0.025, 0.025, 0.025)):
    """RobotModel (environment/robot_model.py) -> (ModelDesc, keep-alive list of the host arrays)."""
    keep = {}

    def dp(name, a):
        keep[name] = _arr(a, np.float64)
        return keep[name].ctypes.data_as(_dp)

    def ip(name, a):
        keep[name] = _arr(a, np.int32)
        return keep[name].ctypes.data_as(_ip)

    d = ModelDesc()
    d.nl = model.nl
    d.parent, d.jtype, d.has_limit = ip('parent', model.parent), ip('jtype', model.jtype), ip('hl', model.has_limit)
    d.E0, d.e, d.d, d.axis = dp('E0', model.E0), dp('e', model.e), dp('d', model.d), dp('axis', model.axis)
    d.mass, d.inertia, d.damping = dp('mass', model.mass), dp('inertia', model.inertia), dp('damping', model.damping)
    d.lower, d.upper = dp('lower', model.lower), dp('upper', model.upper)
    for k in range(9):
        d.base_R[k] = float(model.base_R[k])
    for k in range(3):
        d.base_p[k] = float(model.base_p[k])
        d.gravity[k] = float(model.gravity[k])
        d.target_half[k] = float(target_half[k])
    d.lin_damp, d.ang_damp, d.dt, d.iters = model.lin_damp, model.ang_damp, model.dt, model.iters
    d.resid_thresh, d.erp, d.max_vel = model.resid_thresh, model.erp, model.max_vel
    d.limit_max_impulse = model.limit_max_impulse
    d.ns = model.ns
    d.s_link, d.s_type = ip('s_link', model.s_link), ip('s_type', model.s_type)
    d.s_R, d.s_p, d.s_dim = dp('s_R', model.s_R), dp('s_p', model.s_p), dp('s_dim', model.s_dim)
    d.obstacle_radius = obstacle_radius
    d.ee_link = ee_link
    d.n_obs_joints = n_obs_joints
    verts = np.asarray(getattr(model, 'verts', np.zeros((0, 3))), np.float64).reshape(-1, 3)
    d.n_verts = verts.shape[0]
    if d.n_verts:
        d.s_vert_first, d.s_vert_count = ip('s_v0', model.s_v0), ip('s_vn', model.s_vn)
        d.verts = dp('verts', verts)
    return d, keep
