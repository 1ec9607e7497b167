"""
BatchedSimulator — torch-tensor front end of the native batched simulator (rloa_sim_* in
include/rloa_b200.h).  It plays the role of the PyBullet physics client the reference's Environment
owns (/root/reference/robotic_manipulator_rloa/environment/environment.py:207-210): N independent
arms, one thread each (struct-of-arrays state resident in HBM).  PyTorch is plumbing here (device memory + streams).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import os

import torch

from .. import _native as N
from .robot_model import RobotModel

CONTACT_THRESHOLD = 0.02         # Bullet's gContactBreakingThreshold: contact rows for shapes closer than this
OBSTACLE_RADIUS = 0.075          # sphere_small.urdf (r = 0.03) x globalScaling 2.5, environment.py:252-253
TARGET_HALF = (0.025, 0.025, 0.025)   # cube_small.urdf (0.05 m), environment.py:254-255
RESET_SUBSTEPS = 50              # environment.py:300-301


class BatchedSimulator:
    def __init__(self, model: RobotModel, n_envs: int, endeffector_index: int, involved_joints: Sequence[int],
                 fixed_joints: Sequence[int], max_force: float = 200.0, device: Optional[torch.device] = None,
                 target_threshold: float = 0.05, obstacle_threshold: float = 0.0, contacts: bool = True):
        if not torch.cuda.is_available():
            raise N.NativeLibraryError('BatchedSimulator needs a CUDA device (sm_100a); there is no CPU fallback')
        self.lib = N.lib()
        self.device = torch.device(device if device is not None else 'cuda:0')
        self.model = model
        self.n_envs = int(n_envs)
        self.nl = model.nl
        self.n_act = len(involved_joints)
        self.obs_size = 9 + 2 * self.n_act
        self._h_model = C.c_void_p()
        self._h_sim = C.c_void_p()
        with torch.cuda.device(self.device):
            desc, self._keep = N.make_model_desc(model, endeffector_index, self.n_act, OBSTACLE_RADIUS, TARGET_HALF)
            N.check(self.lib.rloa_model_create(C.byref(desc), C.byref(self._h_model)), 'rloa_model_create')
            N.check(self.lib.rloa_sim_create(self._h_model, self.n_envs, C.byref(self._h_sim)), 'rloa_sim_create')
        cfg = N.StepConfig()
        cfg.n_act = self.n_act
        for k, j in enumerate(involved_joints):
            cfg.act_joint[k] = int(j)
        cfg.n_fixed = len(fixed_joints)
        for k, j in enumerate(fixed_joints):
            cfg.fixed_joint[k] = int(j)
        cfg.max_force = float(max_force)
        cfg.target_threshold = float(target_threshold)
        cfg.obstacle_threshold = float(obstacle_threshold)
        # the reference's obstacle sphere and target cube are collidable fixed bodies (environment.py:252-255): links that
        # touch them are stopped by contact rows inside stepSimulation.  contacts=False gives the free dynamics.
        if os.environ.get('RLOA_CONTACTS') == '0':       # developer A/B switch (cost of the contact rows in bench.py)
            contacts = False
        cfg.contact_threshold = CONTACT_THRESHOLD if contacts else 0.0
        self.cfg = cfg
        N.check(self.lib.rloa_sim_set_contacts(self._h_sim, cfg.contact_threshold), 'rloa_sim_set_contacts')
        f32 = dict(dtype=torch.float32, device=self.device)
        self.obs = torch.zeros(self.n_envs, self.obs_size, **f32)
        self.reward = torch.zeros(self.n_envs, **f32)
        self.done = torch.zeros(self.n_envs, dtype=torch.uint8, device=self.device)
        # bumped by every call that drops the prepared half of the pipelined step on the host side (rloa_sim_prepare's
        # `prepared` flag is baked into a captured CUDA graph): VectorLoop compares it before replaying a graph
        self.generation = 0

    # ------------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _f32(self, x, shape) -> torch.Tensor:
        t = torch.as_tensor(x, dtype=torch.float32, device=self.device)
        if t.dim() < len(shape):
            t = t.expand(*shape)
        t = t.contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f'expected shape {tuple(shape)}, got {tuple(t.shape)}')
        return t

    def close(self) -> None:
        if getattr(self, '_h_sim', None) is not None and self._h_sim:
            torch.cuda.synchronize(self.device)
            self.lib.rloa_sim_destroy(self._h_sim)
            self.lib.rloa_model_destroy(self._h_model)
            self._h_sim = C.c_void_p()
            self._h_model = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def set_task(self, target, obstacle) -> None:
        t = self._f32(target, (self.n_envs, 3))
        o = self._f32(obstacle, (self.n_envs, 3))
        N.check(self.lib.rloa_sim_set_task(self._h_sim, t.data_ptr(), o.data_ptr(), self._stream()), 'rloa_sim_set_task')

    def set_state(self, q, qd) -> None:
        q = self._f32(q, (self.n_envs, self.nl))
        qd = self._f32(qd, (self.n_envs, self.nl))
        self.generation += 1
        N.check(self.lib.rloa_sim_set_state(self._h_sim, q.data_ptr(), qd.data_ptr(), self._stream()),
                'rloa_sim_set_state')

    def get_state(self) -> Tuple[torch.Tensor, torch.Tensor]:
        q = torch.empty(self.n_envs, self.nl, dtype=torch.float32, device=self.device)
        qd = torch.empty_like(q)
        N.check(self.lib.rloa_sim_get_state(self._h_sim, q.data_ptr(), qd.data_ptr(), self._stream()),
                'rloa_sim_get_state')
        return q, qd

    def set_motors(self, kp=None, target_pos=None, target_vel=None, max_impulse=None) -> None:
        ts = [None if x is None else self._f32(x, (self.n_envs, self.nl)) for x in (kp, target_pos, target_vel, max_impulse)]
        N.check(self.lib.rloa_sim_set_motors(self._h_sim, *[N.ptr(t) for t in ts], self._stream()), 'rloa_sim_set_motors')

    def clear(self) -> None:
        self.generation += 1
        N.check(self.lib.rloa_sim_clear(self._h_sim, self._stream()), 'rloa_sim_clear')

    def step(self, actions: torch.Tensor, active: Optional[torch.Tensor] = None, out=None,
             valid: Optional[torch.Tensor] = None):
        """Environment.step for every env.  actions: fp32 [n_envs, n_act] on the device."""
        if actions.dtype != torch.float32 or not actions.is_contiguous() or actions.device != self.device:
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(actions.shape) != (self.n_envs, self.n_act):
            raise ValueError(f'actions must have shape {(self.n_envs, self.n_act)}')
        obs, reward, done = out if out is not None else (self.obs, self.reward, self.done)
        N.check(self.lib.rloa_sim_step(self._h_sim, C.byref(self.cfg), actions.data_ptr(), N.ptr(active), obs.data_ptr(),
                                       reward.data_ptr(), done.data_ptr(), N.ptr(valid), self._stream()),
                'rloa_sim_step')
        return obs, reward, done

    def prepare(self) -> None:
        """Start the action-independent half of the NEXT step (dynamics + M^-1) on the simulator's side stream."""
        N.check(self.lib.rloa_sim_prepare(self._h_sim, self._stream()), 'rloa_sim_prepare')

    def join(self) -> None:
        """Make the current stream wait for an outstanding prepare() (no-op when there is none)."""
        N.check(self.lib.rloa_sim_join(self._h_sim, self._stream()), 'rloa_sim_join')

    def begin_reset(self, init_targets: torch.Tensor, mask: Optional[torch.Tensor] = None,
                    substeps: int = RESET_SUBSTEPS) -> None:
        """Schedule a lock-step asynchronous reset of the masked envs (one sub-step per step() call)."""
        n_init = init_targets.shape[1]
        N.check(self.lib.rloa_sim_begin_reset(self._h_sim, N.ptr(mask), init_targets.data_ptr() if n_init else None,
                                              n_init, int(substeps), self._stream()), 'rloa_sim_begin_reset')

    def begin_reset_random(self, pos: Optional[torch.Tensor], var: Optional[torch.Tensor], n_init: int,
                           mask: Optional[torch.Tensor] = None, seed: int = 0, tick: Optional[torch.Tensor] = None,
                           substeps: int = RESET_SUBSTEPS) -> None:
        """begin_reset with start poses pos + var * U(-1, 1) drawn on the device (pos / var: fp32 [n_init])."""
        N.check(self.lib.rloa_sim_begin_reset_random(self._h_sim, N.ptr(mask), N.ptr(pos), N.ptr(var), int(n_init),
                                                     int(substeps), int(seed) & 0xFFFFFFFFFFFFFFFF, N.ptr(tick),
                                                     self._stream()), 'rloa_sim_begin_reset_random')

    def reset(self, init_targets, mask: Optional[torch.Tensor] = None, substeps: int = RESET_SUBSTEPS,
              obs: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Environment.reset for the masked envs; init_targets fp32 [n_envs, n_init]."""
        init_targets = torch.as_tensor(init_targets, dtype=torch.float32, device=self.device)
        if init_targets.dim() == 1:
            init_targets = init_targets.expand(self.n_envs, -1)
        init_targets = init_targets.contiguous()
        n_init = init_targets.shape[1]
        obs = self.obs if obs is None else obs
        self.generation += 1
        N.check(self.lib.rloa_sim_reset(self._h_sim, N.ptr(mask), init_targets.data_ptr() if n_init else None, n_init,
                                        int(substeps), obs.data_ptr(), self._stream()), 'rloa_sim_reset')
        return obs

    def observe(self, want_distances: bool = False):
        obs = torch.empty(self.n_envs, self.obs_size, dtype=torch.float32, device=self.device)
        if not want_distances:
            N.check(self.lib.rloa_sim_observe(self._h_sim, obs.data_ptr(), None, None, self._stream()), 'rloa_sim_observe')
            return obs
        link = torch.empty(self.n_envs, self.nl, dtype=torch.float32, device=self.device)
        ee = torch.empty(self.n_envs, dtype=torch.float32, device=self.device)
        N.check(self.lib.rloa_sim_observe(self._h_sim, obs.data_ptr(), link.data_ptr(), ee.data_ptr(), self._stream()),
                'rloa_sim_observe')
        return obs, link, ee

    def self_distances(self) -> torch.Tensor:
        """[n_envs][nl][nl] link-link closest distances (10 on the diagonal / adjacent links / shapeless links)."""
        out = torch.empty(self.n_envs, self.nl, self.nl, dtype=torch.float32, device=self.device)
        N.check(self.lib.rloa_sim_self_distances(self._h_sim, out.data_ptr(), self._stream()), 'rloa_sim_self_distances')
        return out

    def contact_counts(self) -> torch.Tensor:
        """Contact rows each env's next step will carry (-1: unknown until that step's own collision phase)."""
        out = torch.empty(self.n_envs, dtype=torch.int32, device=self.device)
        N.check(self.lib.rloa_sim_contact_counts(self._h_sim, out.data_ptr(), self._stream()), 'rloa_sim_contact_counts')
        return out

    def last_iterations(self) -> torch.Tensor:
        it = torch.empty(self.n_envs, dtype=torch.int32, device=self.device)
        N.check(self.lib.rloa_sim_last_iterations(self._h_sim, it.data_ptr(), self._stream()), 'rloa_sim_last_iterations')
        return it
