"""
NAFAgent — drop-in for /root/reference/robotic_manipulator_rloa/naf_components/naf_algorithm.py with the
whole loop device-resident: batched policy inference (rloa_naf_act), the batched simulator step, the
HBM replay ring, and NAFAgent.learn as two native calls (rloa_naf_learn_grads -> optional NCCL gradient
all-reduce -> rloa_naf_learn_apply = clip-norm + Adam + soft target update).

Constructor keywords, attributes, ``run()`` return value, checkpoint files and log lines follow the
reference (naf_algorithm.py:27-89, 228-292).  Reference quirks are reproduced by default (SURVEY.md
Appendix B): replayed actions are truncated like ``.long()``, the TD target ignores ``done``, both nets
use train-mode BatchNorm in ``learn``, ``act`` always adds N(0, P^-1) noise.
"""
from __future__ import annotations

import ctypes as C
import json
import logging
import os
import random
import time
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from numpy.typing import NDArray

from .. import _native as N
from ..utils.exceptions import MissingWeightsFile
from ..utils.logger import get_global_logger
from ..utils.replay_buffer import ReplayBuffer
from ..utils import distributed as rdist
from .naf_neural_network import NAF, NafWorkspace

logger = get_global_logger()


class FusedAdam:
    """State of the fused clip + Adam + soft-update kernel (flat moments in nn.Module.parameters() order).
    Stands where the reference keeps ``optim.Adam`` (naf_algorithm.py:83); ``step()`` is not a separate call —
    the update happens inside ``NAFAgent.learn``."""

    def __init__(self, n_params: int, lr: float, device: torch.device):
        self.defaults = dict(lr=lr, betas=(0.9, 0.999), eps=1e-8)
        self.param_groups = [dict(self.defaults)]
        self.device = device
        self.n_params = n_params
        self.exp_avg = None
        self.exp_avg_sq = None
        self.step_count = None
        if torch.device(device).type == 'cuda':
            self.exp_avg = torch.zeros(n_params, dtype=torch.float32, device=device)
            self.exp_avg_sq = torch.zeros(n_params, dtype=torch.float32, device=device)
            self.step_count = torch.zeros(1, dtype=torch.int64, device=device)

    def zero_grad(self) -> None:        # gradients are overwritten, never accumulated
        pass

    def state_dict(self) -> dict:
        return {'state': {'exp_avg': self.exp_avg, 'exp_avg_sq': self.exp_avg_sq, 'step': self.step_count},
                'param_groups': self.param_groups}

    def native(self) -> N.AdamState:
        a = N.AdamState()
        a.m, a.v, a.step = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.step_count.data_ptr()
        return a


class NAFAgent:
    MODEL_PATH = 'model.p'

    def __init__(self, environment, state_size: int, action_size: int, layer_size: int, batch_size: int,
                 buffer_size: int, learning_rate: float, tau: float, gamma: float, update_freq: int, num_updates: int,
                 checkpoint_frequency: int, device: torch.device, seed: int) -> None:
        os.makedirs('checkpoints/', exist_ok=True)
        self.environment = environment
        self.state_size = state_size
        self.action_size = action_size
        self.layer_size = layer_size
        self.buffer_size = buffer_size
        self.learning_rate = learning_rate
        random.seed(seed)
        self.device = torch.device(device)
        self.tau = tau
        self.gamma = gamma
        self.update_freq = update_freq
        self.num_updates = num_updates
        self.batch_size = batch_size
        self.checkpoint_frequency = checkpoint_frequency
        self.seed = seed
        # opt-in corrections of reference quirks (defaults reproduce the reference)
        self.truncate_replayed_actions = True       # replay_buffer.py:60 `.long()`
        self.use_done_mask = False                  # naf_algorithm.py:199 ignores `done`
        self.noise_scale = 1.0                      # naf_neural_network.py:119-121 noise at every act()

        self.qnetwork_main = NAF(state_size, action_size, layer_size, seed, self.device).to(self.device)
        self.qnetwork_target = NAF(state_size, action_size, layer_size, seed, self.device).to(self.device)
        self.n_params = sum(p.numel() for p in self.qnetwork_main.parameters())
        self.optimizer = FusedAdam(self.n_params, learning_rate, self.device)
        n_envs = getattr(environment, 'n_envs', 1) if environment is not None else 1
        # one process per GPU: the network initialisation above shares `seed` (parameters start identical on every
        # rank), exploration noise and replay sampling get per-rank streams — otherwise N ranks would compute the same
        # rollout and the same gradient N times
        self.rank, self.world_size = rdist.world()
        self.noise_seed = seed if self.world_size == 1 else rdist.rank_seed(seed, self.rank, 0)
        replay_seed = seed if self.world_size == 1 else rdist.rank_seed(seed, self.rank, 1)
        self.memory = ReplayBuffer(buffer_size, batch_size, self.device, replay_seed,
                                   state_size if self.device.type == 'cuda' else None,
                                   action_size if self.device.type == 'cuda' else None, max_append=n_envs)
        self.update_t_step = 0
        self._act_calls = 0
        self._tick_base = 0                         # where the next loop's device counter starts (resume)
        self._ws: Optional[NafWorkspace] = None
        self._bufs = None
        self.last_loss = None
        self.last_grad_norm = None
        self._ws_generation = 0                     # bumped when the workspace / trunk mode a captured graph baked in changes
        # N > 1: 'peer' = NVLink peer-memory exchange fused into the optimiser kernels (csrc/grad_exchange.cu),
        # 'nccl' = torch.distributed all-reduce between the two native calls (the baseline); RLOA_GRAD_EXCHANGE selects
        self.grad_exchange_mode = os.environ.get('RLOA_GRAD_EXCHANGE', 'peer')
        self._xchg: Optional[rdist.GradExchange] = None
        self.fused_learn = True                     # single rank: one-call learn (rloa_naf_learn_step)

    # ------------------------------------------------------------------------------------------
    # native plumbing
    def _require_cuda(self) -> None:
        if self.device.type != 'cuda':
            raise N.NativeLibraryError('NAFAgent runs on librloa_b200.so (sm_100a): a CUDA device is required, '
                                       'there is no CPU fallback')

    def _workspace(self, rows: int) -> NafWorkspace:
        if self._ws is None or self._ws.max_batch < rows:
            if self._ws is not None:
                self._ws.close()
            self._ws_generation += 1
            self._ws = NafWorkspace(self.state_size, self.action_size, self.layer_size, max(rows, self.batch_size, 256),
                                    self.device)
            if self.qnetwork_main.trunk_mode:
                self._ws.set_trunk(self.qnetwork_main.trunk_mode)
        return self._ws

    def set_trunk_mode(self, mode: int) -> None:
        """0 = fp32 CUDA-core trunk (reference-exact), 1 = tcgen05 tensor-core trunk (bf16 operands)."""
        self.qnetwork_main.set_trunk_mode(int(mode))
        self.qnetwork_target.set_trunk_mode(int(mode))
        self._ws_generation += 1
        if self._ws is not None:
            self._ws.set_trunk(int(mode))

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _hyper(self) -> N.NafHyper:
        h = N.NafHyper()
        h.gamma, h.tau, h.lr = float(self.gamma), float(self.tau), float(self.learning_rate)
        h.beta1, h.beta2, h.eps, h.clip_norm = 0.9, 0.999, 1e-8, 1.0
        h.trunc_action = 1 if self.truncate_replayed_actions else 0
        h.use_done_mask = 1 if self.use_done_mask else 0
        h.grad_scale = 1.0 / float(self.world_size)
        return h

    def _learn_buffers(self):
        if self._bufs is None:
            f32 = dict(dtype=torch.float32, device=self.device)
            B = self.batch_size
            self._bufs = dict(s=torch.empty(B, self.state_size, **f32), a=torch.empty(B, self.action_size, **f32),
                              r=torch.empty(B, **f32), s2=torch.empty(B, self.state_size, **f32),
                              d=torch.empty(B, **f32), grad=torch.zeros(self.n_params, **f32),
                              loss=torch.zeros(1, **f32), gnorm=torch.zeros(1, **f32))
        return self._bufs

    # ------------------------------------------------------------------------------------------
    # checkpoints (naf_algorithm.py:91-127)
    def initialize_pretrained_agent_from_episode(self, episode: int) -> None:
        if not os.path.isfile(f'checkpoints/{episode}/weights.p'):
            raise MissingWeightsFile
        logger.debug(f'Loading naf_components weights from trained naf_components on episode {episode}...')
        self.qnetwork_main.load_state_dict(torch.load(f'checkpoints/{episode}/weights.p', map_location=self.device))
        self.qnetwork_target.load_state_dict(torch.load(f'checkpoints/{episode}/weights.p', map_location=self.device))
        logger.info(f'Loaded weights from trained naf_components on episode {episode}')

    def initialize_pretrained_agent_from_weights_file(self, weights_path: str) -> None:
        if not os.path.isfile(weights_path):
            raise MissingWeightsFile
        logger.debug('Loading naf_components weights from trained naf_components...')
        self.qnetwork_main.load_state_dict(torch.load(weights_path, map_location=self.device))
        self.qnetwork_target.load_state_dict(torch.load(weights_path, map_location=self.device))
        logger.info('Loaded pre-trained weights for the NN')

    def _cpu_state_dict(self):
        return {k: v.detach().cpu() for k, v in self.qnetwork_main.state_dict().items()}

    # ------------------------------------------------------------------------------------------
    # true resume (an addition; the reference's checkpoints hold the main network's weights only — SURVEY.md
    # section 5: "Not saved: optimizer state, replay buffer, target net, RNG, episode counter")
    def save_training_state(self, path: str) -> None:
        """Everything ``learn`` / ``act`` depend on: both networks, the Adam moments and step, the replay ring, the
        update gate and the counters that key the device RNG streams."""
        self._require_cuda()
        cpu = lambda sd: {k: v.detach().cpu() for k, v in sd.items()}
        torch.save({'format': 1, 'state_size': self.state_size, 'action_size': self.action_size,
                    'layer_size': self.layer_size, 'seed': self.seed,
                    'main': cpu(self.qnetwork_main.state_dict()), 'target': cpu(self.qnetwork_target.state_dict()),
                    'adam': {'exp_avg': self.optimizer.exp_avg.cpu(), 'exp_avg_sq': self.optimizer.exp_avg_sq.cpu(),
                             'step': int(self.optimizer.step_count.item())},
                    'replay': self.memory.state_dict(), 'update_t_step': self.update_t_step,
                    'act_calls': self._act_calls, 'tick_base': self._tick_base}, path)

    def load_training_state(self, path: str) -> None:
        self._require_cuda()
        if not os.path.isfile(path):
            raise MissingWeightsFile
        sd = torch.load(path, map_location='cpu')
        if (sd['state_size'], sd['action_size'], sd['layer_size']) != (self.state_size, self.action_size, self.layer_size):
            raise ValueError('training state was saved for a different network shape')
        self.qnetwork_main.load_state_dict(sd['main'])
        self.qnetwork_target.load_state_dict(sd['target'])
        self.optimizer.exp_avg.copy_(sd['adam']['exp_avg'])
        self.optimizer.exp_avg_sq.copy_(sd['adam']['exp_avg_sq'])
        self.optimizer.step_count.fill_(int(sd['adam']['step']))
        self.memory.load_state_dict(sd['replay'])
        self.update_t_step, self._act_calls = int(sd['update_t_step']), int(sd['act_calls'])
        self._tick_base = int(sd['tick_base'])

    # ------------------------------------------------------------------------------------------
    def act_batch(self, states: torch.Tensor, out: Optional[torch.Tensor] = None,
                  tick: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Eval-mode policy for a batch of device states -> noisy clamped actions [n, A] (stays on the device).
        `tick`: device uint64 counter added to the Philox step (so a captured CUDA graph draws fresh noise)."""
        self._require_cuda()
        n = states.shape[0]
        ws = self._workspace(n)
        if out is None:
            out = torch.empty(n, self.action_size, dtype=torch.float32, device=self.device)
        p = self.qnetwork_main.native_params()
        # Philox step: the host call counter, or — inside the vectorised loop — the device counter alone, so that an
        # eager launch and a replay of the captured graph draw the same noise
        N.check(ws.lib.rloa_naf_act(ws.handle, C.byref(p), states.data_ptr(), n, self.noise_seed,
                                    self._act_calls if tick is None else 0, N.ptr(tick), float(self.noise_scale),
                                    out.data_ptr(), self._stream()), 'rloa_naf_act')
        if tick is None:
            self._act_calls += 1
        return out

    def act(self, state):
        """Reference contract (naf_algorithm.py:158-178): ndarray[S] -> ndarray[A]; device tensors [n, S] map to
        device tensors [n, A]."""
        if isinstance(state, torch.Tensor) and state.dim() == 2:
            return self.act_batch(state.to(device=self.device, dtype=torch.float32).contiguous())
        s = torch.from_numpy(np.asarray(state, dtype=np.float32)).to(self.device).reshape(1, -1)
        return self.act_batch(s).cpu().squeeze().numpy()

    def step(self, state: NDArray, action: NDArray, reward: float, next_state: NDArray, done: int) -> None:
        """Store one transition and learn when due (naf_algorithm.py:129-156)."""
        self.memory.add(state, action, reward, next_state, done)
        self._maybe_learn()

    def _learn_due(self, pending: int = 0, pending_exact: bool = True) -> bool:
        """The update gate of NAFAgent.step (naf_algorithm.py:147-150); advances the update_freq counter.
        `pending` = transitions the caller is about to append before it learns (an upper bound when the append is
        masked: pending_exact False)."""
        self.update_t_step = (self.update_t_step + 1) % self.update_freq
        return self.update_t_step == 0 and self.memory.gate_open(pending, pending_exact)

    def _maybe_learn(self, tick: Optional[torch.Tensor] = None) -> None:
        if self._learn_due():
            for u in range(self.num_updates):
                self.learn_from_memory(tick=tick, sub=u)

    MAX_PENDING_ROWS = 16384

    def one_launch_learn(self) -> bool:
        """True when learn_from_memory is a single kernel (tensor-core trunk, supported batch, and - with N > 1 ranks - the
        peer-memory exchange): the precondition of `pending=` below."""
        m = self.memory
        if not self.fused_learn or m._rb is None or self._ws is None:
            return False
        xchg = self._exchange()
        # N > 1 without the peer-memory exchange (RLOA_GRAD_EXCHANGE=nccl): the all-reduce sits between two native calls, so
        # the update cannot be one launch
        if not (self.world_size == 1 or (xchg is not None and self.world_size <= 8)):
            return False
        ws = self._workspace(self.batch_size)
        return bool(ws.lib.rloa_naf_learn_fused_supported(ws.handle, self.batch_size))

    def learn_from_memory(self, tick: Optional[torch.Tensor] = None, sub: int = 0, pending=None) -> None:
        """sample + learn without leaving the device (replay_buffer.py:47-67 + naf_algorithm.py:180-213).  With the tensor-core
        trunk and a batch of at most 1024 rows this is ONE kernel: the fused learn kernel draws the sampler's slots itself
        and reads its rows straight from the ring (rloa_naf_learn_step_replay).
        pending = (states, actions, rewards, next_states, dones, valid | None): rows being copied into the ring right now
        (ReplayBuffer.add_batch(..., commit=False) on another stream) and committed after this call - the update samples the
        ring as it will be (rloa_naf_learn_step_pending); needs one_launch_learn()."""
        b = self._learn_buffers()
        m = self.memory
        ws = self._workspace(self.batch_size)
        xchg = self._exchange()
        one_launch = self.world_size == 1 or xchg is not None
        if self.fused_learn and one_launch and m._rb is not None and \
                ws.lib.rloa_naf_learn_fused_supported(ws.handle, self.batch_size):
            hp = self._hyper()
            pm, pt = self.qnetwork_main.native_params(), self.qnetwork_target.native_params()
            adam = self.optimizer.native()
            draw = m._draws if tick is None else (int(sub) << 32)
            if pending is not None:
                ps, pa, pr, ps2, pd, pv = pending
                N.check(ws.lib.rloa_naf_learn_step_pending(ws.handle, C.byref(pm), C.byref(pt), C.byref(adam),
                                                           xchg.handle if xchg is not None else None, C.byref(m._rb), m.seed,
                                                           draw, N.ptr(tick), self.batch_size, C.byref(hp), ps.shape[0],
                                                           ps.data_ptr(), pa.data_ptr(), pr.data_ptr(), ps2.data_ptr(),
                                                           N.ptr(pd), N.ptr(pv), b['grad'].data_ptr(), b['loss'].data_ptr(),
                                                           b['gnorm'].data_ptr(), self._stream()), 'rloa_naf_learn_step_pending')
            else:
                N.check(ws.lib.rloa_naf_learn_step_replay(ws.handle, C.byref(pm), C.byref(pt), C.byref(adam),
                                                          xchg.handle if xchg is not None else None, C.byref(m._rb), m.seed, draw,
                                                          N.ptr(tick), self.batch_size, C.byref(hp), b['grad'].data_ptr(),
                                                          b['loss'].data_ptr(), b['gnorm'].data_ptr(), self._stream()),
                        'rloa_naf_learn_step_replay')
            m._draws += 1
            self.last_loss, self.last_grad_norm = b['loss'], b['gnorm']
            return
        if pending is not None:
            raise RuntimeError('learn_from_memory(pending=...) needs the one-kernel update (one_launch_learn())')
        self.memory.sample_into(b['s'], b['a'], b['r'], b['s2'], b['d'], tick=tick, sub=sub)
        self._learn_device(b['s'], b['a'], b['r'], b['s2'], b['d'])

    def prepack(self) -> None:
        """Start writing the tensor-core weight images of the next learn() on the workspace's side stream (a no-op on the
        fp32 / multi-launch paths).  The parameters must not be modified between this call and that learn()."""
        if self._ws is None or not self.qnetwork_main.trunk_mode:
            return
        pm, pt = self.qnetwork_main.native_params(), self.qnetwork_target.native_params()
        N.check(self._ws.lib.rloa_naf_learn_prepack(self._ws.handle, C.byref(pm), C.byref(pt), self._stream()),
                'rloa_naf_learn_prepack')

    def learn(self, experiences: Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]) -> None:
        """Reference signature: a 5-tuple of tensors as ReplayBuffer.sample returns them."""
        self._require_cuda()
        states, actions, rewards, next_states, dones = experiences
        f = lambda t: t.to(device=self.device, dtype=torch.float32).contiguous()
        self._learn_device(f(states), f(actions), f(rewards).reshape(-1), f(next_states), f(dones).reshape(-1))

    def _learn_device(self, s, a, r, s2, d) -> None:
        self._require_cuda()
        B = s.shape[0]
        ws = self._workspace(B)
        b = self._learn_buffers()
        hp = self._hyper()
        pm, pt = self.qnetwork_main.native_params(), self.qnetwork_target.native_params()
        st = self._stream()
        adam = self.optimizer.native()
        xchg = self._exchange()
        if xchg is None and self.world_size == 1 and self.fused_learn:
            # single rank: backward tail, gradient norm and the optimiser in one kernel (rloa_naf_learn_step)
            N.check(ws.lib.rloa_naf_learn_step(ws.handle, C.byref(pm), C.byref(pt), C.byref(adam), s.data_ptr(),
                                               a.data_ptr(), r.data_ptr(), s2.data_ptr(), d.data_ptr(), B, C.byref(hp),
                                               b['grad'].data_ptr(), b['loss'].data_ptr(), b['gnorm'].data_ptr(), st),
                    'rloa_naf_learn_step')
            self.last_loss, self.last_grad_norm = b['loss'], b['gnorm']
            return
        if xchg is not None and self.fused_learn:
            # N ranks: backward, then the gradient exchange over NVLink peer memory inside the optimiser kernels
            N.check(ws.lib.rloa_naf_learn_step_xchg(ws.handle, C.byref(pm), C.byref(pt), C.byref(adam), xchg.handle,
                                                    s.data_ptr(), a.data_ptr(), r.data_ptr(), s2.data_ptr(), d.data_ptr(), B,
                                                    C.byref(hp), b['grad'].data_ptr(), b['loss'].data_ptr(),
                                                    b['gnorm'].data_ptr(), st), 'rloa_naf_learn_step_xchg')
            self.last_loss, self.last_grad_norm = b['loss'], b['gnorm']
            return
        N.check(ws.lib.rloa_naf_learn_grads(ws.handle, C.byref(pm), C.byref(pt), s.data_ptr(), a.data_ptr(),
                                            r.data_ptr(), s2.data_ptr(), d.data_ptr(), B, C.byref(hp),
                                            b['grad'].data_ptr(), b['loss'].data_ptr(), st), 'rloa_naf_learn_grads')
        if xchg is not None:             # the one exchange step of the data-parallel path, inside the optimiser kernels
            N.check(ws.lib.rloa_naf_learn_apply_xchg(ws.handle, C.byref(pm), C.byref(pt), C.byref(adam), C.byref(hp),
                                                     xchg.handle, b['grad'].data_ptr(), b['gnorm'].data_ptr(), st),
                    'rloa_naf_learn_apply_xchg')
        else:
            if self.world_size > 1:      # baseline: NCCL all-reduce between the two native calls
                rdist.allreduce_gradient(b['grad'])
            N.check(ws.lib.rloa_naf_learn_apply(ws.handle, C.byref(pm), C.byref(pt), C.byref(adam), C.byref(hp),
                                                b['grad'].data_ptr(), b['gnorm'].data_ptr(), st), 'rloa_naf_learn_apply')
        self.last_loss, self.last_grad_norm = b['loss'], b['gnorm']

    def _exchange(self) -> Optional[rdist.GradExchange]:
        """The peer-memory exchange when it applies (N > 1 and mode 'peer', or forced with mode 'peer-always');
        created lazily — collectively — on the first learn()."""
        if self._xchg is None and (
                (self.world_size > 1 and self.grad_exchange_mode == 'peer') or self.grad_exchange_mode == 'peer-always'):
            self._xchg = rdist.GradExchange.create(self.n_params, self.device)
            if self._xchg is None:      # some rank cannot export / map the blocks: every rank falls back together
                logger.warning('NVLink peer-memory gradient exchange unavailable, using the NCCL all-reduce')
                self.grad_exchange_mode = 'nccl'
        return self._xchg

    def soft_update(self, main_nn: NAF, target_nn: NAF) -> None:
        """theta_target = tau theta_main + (1 - tau) theta_target over parameters (naf_algorithm.py:217-226)."""
        self._require_cuda()
        pm, pt = main_nn.native_params(), target_nn.native_params()
        N.check(N.lib().rloa_naf_soft_update(C.byref(pm), C.byref(pt), float(self.tau), self._stream()),
                'rloa_naf_soft_update')

    # ------------------------------------------------------------------------------------------
    def make_loop(self, frames: int, log_capacity: int, learn: bool = True, store: bool = True) -> 'VectorLoop':
        """`store=False` (evaluation rollouts): transitions are not appended to the replay ring."""
        return VectorLoop(self, frames, log_capacity, learn, store)

    def run(self, frames: int = 1000, episodes: int = 1000, verbose: bool = True) -> Dict[int, Tuple[float, int]]:
        """Training loop (naf_algorithm.py:228-292) over ``environment.n_envs`` arms in lock step.
        ``episodes`` counts completed episodes over all envs, numbered in completion order; with one env the
        sequence of resets, steps, updates, log lines and checkpoints is the reference's.

        One process per GPU (torchrun): ``episodes`` counts over ALL ranks.  Every learn() is a collective, so the exit
        test uses the all-reduced count read at the same iteration on every rank — all ranks run the same number of
        updates; each rank logs its own episodes, rank 0 alone writes checkpoints and model.p, and a gradient
        exchange that timed out raises instead of letting the ranks drift apart."""
        self._require_cuda()
        env = self.environment
        n = env.n_envs
        multi = self.world_size > 1
        logger.info('Training started')
        scores = {episode: (0, 0) for episode in range(1, episodes + 1)}
        loop = self.make_loop(frames, episodes + 2 * n + 1)
        completed, start = 0, time.time()       # local episodes recorded so far
        total = 0                               # completed episodes over all ranks (== completed on a single rank)
        if n == 1:
            logger.info(f'Running Episode {completed + 1}')
        loop.reset_all(verbose)
        sync_every = 1 if n == 1 else (16 if n >= 256 else 4)      # iterations between host reads of the episode log
        it = 0
        while total < episodes:
            if n == 1 and verbose:
                logger.info(f'Running frame {int(loop.frame.item()) + 1} in episode {completed + 1}')
                logger.info(f'Current State: {loop.state[0].cpu().numpy()}')
            if n == 1:
                loop.step(auto_reset=False)
                it += 1
            else:               # pairs of iterations replay the captured CUDA graph once the loop is in steady state
                loop.run_steps(sync_every)
                it += sync_every
            if n == 1 and verbose:
                logger.info(f'Action chosen for the given state is: {loop.actions[0].cpu().numpy()}')
                logger.info(f'Reward: {float(loop.reward.item())}\n')
            if it % sync_every:
                continue
            if multi:           # same iteration on every rank: one tiny all-reduce, then the host read
                cnt = loop.log_count.to(torch.int64)
                glob = cnt.clone()
                rdist.dist.all_reduce(glob, op=rdist.dist.ReduceOp.SUM)
                n_done, total = (int(x) for x in torch.stack([cnt[0], glob[0]]).cpu())
                self._check_exchange()
            else:
                n_done = total = int(loop.log_count.item())          # the one host read of the loop
            if n_done <= completed:
                continue
            hi = min(n_done, episodes)
            ls, lf = loop.log_score[completed:hi].cpu().numpy(), loop.log_frame[completed:hi].cpu().numpy()
            info = logger.isEnabledFor(logging.INFO)
            for k in range(hi - completed):
                ep = completed + k + 1
                scores[ep] = (float(ls[k]), int(lf[k]))
                if info:
                    logger.info(f'Reward:                             {float(ls[k])}')
                    logger.info(f'Number of frames:                   {int(lf[k])}')
                    logger.info(f'Mean of rewards on this episode:    {float(ls[k]) / frames}')
                    logger.info(f'Time taken for this episode:        {round(time.time() - start, 3)} secs\n')
                if ep % self.checkpoint_frequency == 0 and self.rank == 0:
                    os.makedirs(f'checkpoints/{ep}/', exist_ok=True)
                    torch.save(self._cpu_state_dict(), f'checkpoints/{ep}/weights.p')
                    with open(f'checkpoints/{ep}/scores.txt', 'w') as f:
                        f.write(json.dumps(scores))
            completed = hi
            start = time.time()
            if total < episodes and n == 1:
                logger.info(f'Running Episode {completed + 1}')
                loop.reset_all(verbose)
        self._tick_base = int(loop.tick.item())      # a later run() / a resumed agent continues the RNG streams
        self._check_exchange()
        if self.rank == 0:
            torch.save(self._cpu_state_dict(), self.MODEL_PATH)
            logger.info(f'Model has been successfully saved in {self.MODEL_PATH}')
        return scores

    def _check_exchange(self) -> None:
        """Raise when a device-side wait of the peer-memory gradient exchange gave up: from that update on the
        optimiser kernels skip their update (grad_exchange.cu), so parameters are stale, not divergent."""
        if self._xchg is not None and self._xchg.timed_out():
            raise RuntimeError(f'rank {self.rank}: the NVLink gradient exchange timed out waiting for a peer '
                               '(a rank left the training loop or died); parameters stopped updating at that point')


class VectorLoop:
    """Device-resident buffers and one sync-free iteration of the vectorised act -> step -> store -> learn loop
    (the body of naf_algorithm.py:249-270 for every env at once).  Nothing in ``step`` reads back to the host,
    every random draw is keyed by the device counter ``tick``, and two consecutive iterations (the state buffers
    ping-pong) can therefore be captured once into a CUDA graph and replayed (``run_steps``)."""

    def __init__(self, agent: NAFAgent, frames: int, log_capacity: int, learn: bool = True, store: bool = True):
        self.agent, self.env = agent, agent.environment
        self.frames, self.learn = int(frames), learn
        self.store = bool(store) or bool(learn)
        n, dev = self.env.n_envs, agent.device
        self.n = n
        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.state = torch.zeros(n, agent.state_size, **f32)
        self.next_state = torch.zeros(n, agent.state_size, **f32)
        self.reward = torch.zeros(n, **f32)
        self.done = torch.zeros(n, **u8)
        self.valid = torch.ones(n, **u8)
        self.actions = torch.zeros(n, agent.action_size, **f32)
        self.score = torch.zeros(n, **f32)
        self.frame = torch.zeros(n, **i32)
        self.reset_mask = torch.zeros(n, **u8)
        self.cap = int(log_capacity)
        self.log_score = torch.zeros(self.cap, **f32)
        self.log_frame = torch.zeros(self.cap, **i32)
        self.log_last = torch.zeros(self.cap, **f32)
        self.log_env = torch.zeros(self.cap, **i32)
        self.log_count = torch.zeros(1, **i32)
        self.transitions = torch.zeros(1, dtype=torch.int64, device=dev)
        # device loop counter (uint64 bits) as a ping-pong pair: iteration p reads _ticks[p] and its bookkeeping kernel
        # writes _ticks[1 - p] = _ticks[p] + 1, so that kernel can run beside the update that still reads _ticks[p]
        self._ticks = torch.full((2,), int(agent._tick_base), dtype=torch.int64, device=dev)
        self._par = 0                   # which counter is current
        self._odd = 0                   # iterations since the state buffers last matched the captured graph (mod 2)
        self._graph_par = 0
        self.lib = N.lib()
        self.phase_events = None        # optional [6 events] x 2 iterations: act | sim | append | learn | tail
        self._graph = None              # CUDA graph of two consecutive iterations
        self._graph_gen = None          # _generation() at capture time
        self._host_gen = None
        self._graph_failed = False      # a capture attempt raised: stay on eager launches
        self._graph_learn = False
        self.graph_kernels = 0
        self.graph_error: Optional[str] = None
        self.pipeline_sim = True        # rloa_sim_prepare: the next step's dynamics run beside this step's learn
        self._host = None               # pinned host buffers bound by bind_host_buffers
        self._host_ready = None         # event behind the D2H of (state, reward, done) in the host-facing iteration
        self._copy_stream = None
        # one-kernel update: the copy into the replay ring runs beside it (see _body); RLOA_OVERLAP_STORE=0: in sequence (A/B)
        self.overlap_store = os.environ.get('RLOA_OVERLAP_STORE', '1') != '0'
        self._store_stream = None
        self._fork_stream = None
        self._store_pending = False     # a commit on the store stream that the main stream has not waited for yet
        self._host_graphs = None

    @property
    def tick(self) -> torch.Tensor:
        """The current loop counter (a 1-element view; pass it as `tick=` / read it with .item())."""
        return self._ticks[self._par:self._par + 1]

    def reset_all(self, verbose: bool = False) -> None:
        """Synchronous Environment.reset of every env (50 sub-steps) -> self.state."""
        if self.n == 1:
            s = np.asarray(self.env.reset(verbose), dtype=np.float32)
            self.state.copy_(torch.from_numpy(s).to(self.state.device).reshape(1, -1))
        else:
            self.env.reset_batch(obs=self.state)
        self.score.zero_()
        self.frame.zero_()

    def _mark(self, parity: int, k: int) -> None:
        """Record phase-boundary event k of iteration `parity` (bench.py installs the events; no-op otherwise)."""
        if self.phase_events is not None:
            self.phase_events[parity][k].record()

    def _join_store(self) -> None:
        if self._store_pending:
            torch.cuda.current_stream(self.agent.device).wait_stream(self._store_stream)
            self._store_pending = False

    def _body(self, auto_reset: bool, learn_now: bool, parity: int, join_store: bool = True) -> None:
        a, env = self.agent, self.env
        self._mark(parity, 0)
        if learn_now:             # the weights are final since the last update: their images are written beside act / step
            a.prepack()
        a.act_batch(self.state, out=self.actions, tick=self.tick)
        self._mark(parity, 1)
        env.sim.step(self.actions, out=(self.next_state, self.reward, self.done), valid=self.valid)
        # One-kernel update (tensor-core trunk): its 16 CTAs need a whole SM each (all 64 K registers, 231 KB of shared memory),
        # so it is enqueued FIRST after the step - behind it, on forked streams that only wait for the step: the dynamics +
        # M^-1 of the next step, the episode bookkeeping and the copy of the step's rows into the replay ring (the update
        # reads drawn slots of the pending range from the loop's own buffers; the cursor moves once it has finished: same
        # samples, same ring).  Launched the other way round, the 128 one-warp blocks of the dynamics kernel sit on 128 SMs
        # and the update waits ~15 us for 16 empty ones.
        valid = self.valid if auto_reset else None
        overlap = (self.store and learn_now and self.overlap_store and a.num_updates == 1 and self.n <= a.MAX_PENDING_ROWS
                   and a.memory._rb is not None and self.n <= a.memory.buffer_size and a.one_launch_learn())
        cur = torch.cuda.current_stream(a.device)
        tick = self.tick
        common = (self.reward.data_ptr(), self.done.data_ptr(), self.valid.data_ptr(), self.score.data_ptr(),
                  self.frame.data_ptr(), self.reset_mask.data_ptr(), self.log_score.data_ptr(), self.log_frame.data_ptr(),
                  self.log_last.data_ptr(), self.log_env.data_ptr(), self.cap, self.log_count.data_ptr(),
                  self.transitions.data_ptr(), tick.data_ptr())

        def bookkeeping_reset():
            # bookkeeping + the finished envs start their 50 reset sub-steps (one per following step); the counter
            # for the next iteration goes to the other slot (see _ticks)
            nxt = self._ticks[1 - self._par:2 - self._par]
            N.check(self.lib.rloa_episode_update_reset(env.sim._h_sim, self.frames, *common, nxt.data_ptr(),
                                                       env._d_pos.data_ptr(), env._d_var.data_ptr(), env._n_init(), 50,
                                                       (env.seed + 0x5EED) & 0xFFFFFFFFFFFFFFFF, a._stream()),
                    'rloa_episode_update_reset')

        early = auto_reset and self.pipeline_sim      # beside the update, on the simulator's side stream
        self._join_store()                            # the previous step's commit (long finished)
        if overlap:
            if self._store_stream is None:
                self._store_stream, self._fork_stream = torch.cuda.Stream(a.device), torch.cuda.Stream(a.device)
            stepped = cur.record_event()
            self._mark(parity, 2)
            self._mark(parity, 3)
            a.learn_from_memory(tick=tick, sub=0, pending=(self.state, self.actions, self.reward, self.next_state, self.done, valid))
            self._mark(parity, 4)
            self._fork_stream.wait_event(stepped)
            with torch.cuda.stream(self._fork_stream):
                if self.pipeline_sim:
                    env.sim.prepare()
                if early:
                    bookkeeping_reset()
            self._store_stream.wait_event(stepped)
            with torch.cuda.stream(self._store_stream):
                a.memory.add_batch(self.state, self.actions, self.reward, self.next_state, self.done, valid=valid, commit=False)
            self._store_stream.wait_stream(cur)
            with torch.cuda.stream(self._store_stream):
                a.memory.commit_rows(self.n, valid)
            self._store_pending = True
            cur.wait_stream(self._fork_stream)        # nothing but the fork itself lives there; the work is joined below
        else:
            if self.pipeline_sim:     # dynamics + M^-1 of the NEXT step overlap the replay / learn phase of this one
                env.sim.prepare()
            self._mark(parity, 2)
            if self.store:
                a.memory.add_batch(self.state, self.actions, self.reward, self.next_state, self.done, valid=valid)
            self._mark(parity, 3)
            if early:
                bookkeeping_reset()
            if learn_now:
                for u in range(a.num_updates):
                    a.learn_from_memory(tick=tick, sub=u)
            self._mark(parity, 4)
        if auto_reset and not early:
            bookkeeping_reset()
        elif not auto_reset:
            N.check(self.lib.rloa_episode_update(self.n, self.frames, *common, a._stream()), 'rloa_episode_update')
        if self.pipeline_sim:
            env.sim.join()
        if join_store:            # eager steps leave the ring committed on the caller's stream; inside the captured pair only
            self._join_store()    # the last iteration joins (the next one waits where it touches the ring)
        self._mark(parity, 5)
        self.state, self.next_state = self.next_state, self.state
        self._odd ^= 1
        if auto_reset:
            self._par ^= 1

    def step(self, auto_reset: bool = True) -> None:
        learn_now = self.learn and self.agent._learn_due(pending=self.n, pending_exact=not auto_reset)
        self._body(auto_reset, learn_now, 0)

    # ---- CUDA graph of two iterations ---------------------------------------------------------------------
    def _graphable(self) -> bool:
        a = self.agent
        return (not self.learn) or (a.update_freq == 1 and a.memory.gate_open())

    def _generation(self):
        """Host-side state a captured graph bakes in: the simulator's prepared flag (dropped by reset / set_state /
        clear) and the NAF workspace pointers / trunk mode."""
        return (self.env.sim.generation, self.agent._ws_generation, id(self.agent._ws))

    def capture(self) -> bool:
        """Capture two consecutive auto-reset iterations (eager warm-up must have happened: workspaces exist).
        Returns False (and keeps the eager path) when the loop is not in steady state or capture fails."""
        if self._graph is not None:
            return True
        a = self.agent
        if not self._graphable() or a._ws is None or (self.learn and a._bufs is None):
            return False
        host_len = a.memory._len
        try:
            torch.cuda.synchronize(self.agent.device)
            g = torch.cuda.CUDAGraph()
            launched = self.lib.rloa_launch_count()
            # thread_local: other host threads (NCCL watchdog, clock sampler) may keep calling the CUDA API
            with torch.cuda.graph(g, capture_error_mode='thread_local'):
                self._body(True, self.learn, 0, join_store=False)
                self._body(True, self.learn, 1)
            self._graph, self._graph_learn = g, self.learn
            self._odd, self._graph_par = 0, self._par
            self._graph_gen = self._generation()
            self.graph_kernels = int(self.lib.rloa_launch_count() - launched)    # librloa kernels per replay
            a.memory._len = host_len          # capturing does not execute: roll the host mirror back
            return True
        except Exception as err:          # pragma: no cover - depends on driver / NCCL capture support
            self.graph_error = f'{type(err).__name__}: {err}'
            self._graph = None
            self._graph_failed = True
            a.memory._len = host_len
            torch.cuda.synchronize(self.agent.device)
            return False

    def replay_pair(self) -> None:
        """Two iterations through the captured graph."""
        if self._odd or self._par != self._graph_par:
            raise RuntimeError('the loop buffers are not in the phase the graph was captured in; use run_steps')
        if self._graph_gen != self._generation():
            raise RuntimeError('the simulator state or the NAF workspace changed since the graph was captured '
                               '(reset / set_state / set_trunk_mode / a larger batch); use run_steps, which re-captures')
        self._graph.replay()
        m = self.agent.memory
        if self.store:
            m._len = min(m.buffer_size, m._len + 2 * self.n)

    # ---- host-facing iteration: states / actions / results cross pinned HOST buffers every step ---------------
    def bind_host_buffers(self, h_state: torch.Tensor, h_action: torch.Tensor, h_reward: torch.Tensor,
                          h_done: torch.Tensor) -> None:
        """Pinned host tensors [n,S] f32, [n,A] f32, [n] f32, [n] u8 that `step_host` reads and writes: the caller
        owns the observations and sees every action, reward and done flag, as with the reference's
        ``env.step(agent.act(state))`` (naf_algorithm.py:249-262), for all envs at once."""
        for t in (h_state, h_action, h_reward, h_done):
            if not t.is_pinned():
                raise ValueError('step_host needs pinned host tensors')
        self._host = (h_state, h_action, h_reward, h_done)
        self._host_graphs = None
        try:        # `external`: the record can sit inside the captured graph as an event-record node
            self._host_ready = torch.cuda.Event(external=True)
        except TypeError:      # pragma: no cover - older torch: fall back to a full stream synchronisation
            self._host_ready = None

    def _host_phase_act(self) -> None:
        h_state, h_action, _, _ = self._host
        self.state.copy_(h_state, non_blocking=True)                      # H2D: this step's states
        self.agent.act_batch(self.state, out=self.actions, tick=self.tick)
        h_action.copy_(self.actions, non_blocking=True)                   # D2H: the actions, for the caller

    def _host_phase_step(self, learn_now: bool) -> None:
        h_state, h_action, h_reward, h_done = self._host
        a, env = self.agent, self.env
        self.actions.copy_(h_action, non_blocking=True)                   # H2D: actions into Environment.step
        env.sim.step(self.actions, out=(self.next_state, self.reward, self.done), valid=self.valid)
        cur = torch.cuda.current_stream(a.device)
        # same schedule as _body: the one-kernel update first, everything that only needs the step forked behind it
        overlap = (learn_now and self.overlap_store and a.num_updates == 1 and self.n <= a.MAX_PENDING_ROWS
                   and a.memory._rb is not None and self.n <= a.memory.buffer_size and a.one_launch_learn())
        self._join_store()
        stepped = cur.record_event()
        # D2H of (state, reward, done) on a copy stream: it overlaps the store / learn phase and is joined at the end
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(a.device)
        self._copy_stream.wait_event(stepped)
        with torch.cuda.stream(self._copy_stream):
            h_state.copy_(self.next_state, non_blocking=True)
            h_reward.copy_(self.reward, non_blocking=True)
            h_done.copy_(self.done, non_blocking=True)
            if self._host_ready is not None:      # what step_host waits for: the caller's buffers are complete here
                self._host_ready.record(self._copy_stream)
        if overlap:
            if self._store_stream is None:
                self._store_stream, self._fork_stream = torch.cuda.Stream(a.device), torch.cuda.Stream(a.device)
            a.learn_from_memory(tick=self.tick, sub=0,
                                pending=(self.state, self.actions, self.reward, self.next_state, self.done, self.valid))
            if self.pipeline_sim:
                self._fork_stream.wait_event(stepped)
                with torch.cuda.stream(self._fork_stream):
                    env.sim.prepare()
                cur.wait_stream(self._fork_stream)
            self._store_stream.wait_event(stepped)
            with torch.cuda.stream(self._store_stream):
                a.memory.add_batch(self.state, self.actions, self.reward, self.next_state, self.done, valid=self.valid, commit=False)
            self._store_stream.wait_stream(cur)
            with torch.cuda.stream(self._store_stream):
                a.memory.commit_rows(self.n, self.valid)
            self._store_pending = True
        else:
            if self.pipeline_sim:
                env.sim.prepare()
            a.memory.add_batch(self.state, self.actions, self.reward, self.next_state, self.done, valid=self.valid)
            if learn_now:
                for u in range(a.num_updates):
                    a.learn_from_memory(tick=self.tick, sub=u)
        N.check(self.lib.rloa_episode_update_reset(
            env.sim._h_sim, self.frames, self.reward.data_ptr(), self.done.data_ptr(), self.valid.data_ptr(),
            self.score.data_ptr(), self.frame.data_ptr(), self.reset_mask.data_ptr(), self.log_score.data_ptr(),
            self.log_frame.data_ptr(), self.log_last.data_ptr(), self.log_env.data_ptr(), self.cap,
            self.log_count.data_ptr(), self.transitions.data_ptr(), self.tick.data_ptr(), None, env._d_pos.data_ptr(),
            env._d_var.data_ptr(), env._n_init(), 50, (env.seed + 0x5EED) & 0xFFFFFFFFFFFFFFFF, a._stream()),
            'rloa_episode_update_reset')
        if self.pipeline_sim:
            env.sim.join()
        cur.wait_stream(self._copy_stream)
        self._join_store()

    def step_host(self, use_graph: bool = True) -> None:
        """One iteration through the bound host buffers: [H2D states, act, D2H actions] sync
        [H2D actions, Environment.step, D2H state/reward/done | store, learn, bookkeeping] wait for the D2H.  In steady
        state the two halves are CUDA graphs (copies included).  The call returns as soon as the caller's buffers hold
        this step's (state, reward, done): store / learn / bookkeeping of this step may still be running, and the next
        call's work queues behind them on the same stream (torch.cuda.synchronize() to drain)."""
        a = self.agent
        stream = torch.cuda.current_stream(a.device)
        if self._host_graphs and self._host_gen != self._generation():
            self._host_graphs = None
        if use_graph and self._host_graphs is None and self._graphable() and a._ws is not None and \
                (not self.learn or a._bufs is not None):
            host_len = a.memory._len
            try:
                if self.pipeline_sim:      # the graphs bake in whether Environment.step finds its first half prepared: capture
                    self.env.sim.prepare()  # the steady state (every replay leaves the next step prepared), not a cold start
                    self.env.sim.join()
                torch.cuda.synchronize(a.device)
                g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1, capture_error_mode='thread_local'):
                    self._host_phase_act()
                with torch.cuda.graph(g2, capture_error_mode='thread_local'):
                    self._host_phase_step(self.learn)
                self._host_graphs = (g1, g2)
                self._host_gen = self._generation()
            except Exception as err:      # pragma: no cover
                self.graph_error = f'{type(err).__name__}: {err}'
                self._host_graphs = False
                torch.cuda.synchronize(a.device)
            a.memory._len = host_len
        if use_graph and self._host_graphs:
            self._host_graphs[0].replay()
            stream.synchronize()
            self._host_graphs[1].replay()
            a.memory._len = min(a.memory.buffer_size, a.memory._len + self.n)
        else:
            self._host_phase_act()
            stream.synchronize()
            self._host_phase_step(self.learn and a._learn_due(pending=self.n, pending_exact=False))
        if self._host_ready is not None:
            self._host_ready.synchronize()
        else:
            stream.synchronize()

    def run_steps(self, k: int, use_graph: bool = True) -> None:
        """k iterations with auto-reset; pairs go through the CUDA graph when it is (or can be) captured."""
        if self._graph is not None and self._graph_gen != self._generation():
            self._graph = None             # stale: baked-in prepared flag / workspace pointers; capture again below
        if use_graph and self._graph is not None:
            off_state, off_tick = bool(self._odd), self._par != self._graph_par
            if off_state != off_tick:      # step(auto_reset=False) calls since the capture moved only one of the two
                self._graph = None         # phases: no eager iteration can realign them, capture again instead
            elif off_state and k >= 1:
                self.step()                # an odd number of eager iterations since the capture: realign the buffers
                k -= 1
        if use_graph and k >= 2 and not self._graph_failed and self.capture():
            while k >= 2:
                self.replay_pair()
                k -= 2
        for _ in range(k):
            self.step()
