"""
ReplayBuffer — drop-in for /root/reference/robotic_manipulator_rloa/utils/replay_buffer.py, with the
storage moved into HBM: a struct-of-arrays ring (rloa_replay_append) instead of a host deque of
namedtuples, and a device sampler + gather (rloa_replay_sample) instead of random.sample + np.stack
+ five H2D copies.  ``sample()`` keeps the reference dtypes: states f32 [B,S], actions int64 [B,A]
(the .long() cast of replay_buffer.py:60), rewards f32 [B,1], next_states f32 [B,S], dones f32 [B,1].
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from .. import _native as N


class ReplayBuffer:

    def __init__(self, buffer_size: int, batch_size: int, device: torch.device, seed: int,
                 state_size: Optional[int] = None, action_size: Optional[int] = None, max_append: int = 1):
        self.device = torch.device(device)
        self.buffer_size = int(buffer_size)
        self.batch_size = int(batch_size)
        self.seed = int(seed)
        self._draws = 0
        self._len = 0                       # host mirror of min(cursor, capacity)
        self._len_exact = True              # False after a masked append of unknown size (the mirror is an upper bound)
        self._gate_open = False             # sticky: the ring holds more than batch_size rows (it never shrinks)
        self._max_append = int(max_append)
        self._rb: Optional[N.Replay] = None
        self.state_size, self.action_size = state_size, action_size
        if state_size is not None and action_size is not None:
            self._allocate(state_size, action_size)

    def _allocate(self, state_size: int, action_size: int) -> None:
        if self.device.type != 'cuda':
            raise N.NativeLibraryError('ReplayBuffer storage lives in HBM: a CUDA device is required (no CPU fallback)')
        self.lib = N.lib()
        self.state_size, self.action_size = int(state_size), int(action_size)
        f32 = dict(dtype=torch.float32, device=self.device)
        cap = self.buffer_size
        self.states = torch.zeros(cap, self.state_size, **f32)
        self.next_states = torch.zeros(cap, self.state_size, **f32)
        self.actions = torch.zeros(cap, self.action_size, **f32)
        self.rewards = torch.zeros(cap, **f32)
        self.dones = torch.zeros(cap, **f32)
        self.cursor = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.scratch = torch.zeros((self._max_append + 63) // 64 + 2, dtype=torch.int32, device=self.device)
        rb = N.Replay()
        rb.capacity, rb.state_size, rb.action_size = cap, self.state_size, self.action_size
        rb.states, rb.actions, rb.rewards = self.states.data_ptr(), self.actions.data_ptr(), self.rewards.data_ptr()
        rb.next_states, rb.dones, rb.cursor = self.next_states.data_ptr(), self.dones.data_ptr(), self.cursor.data_ptr()
        rb.scratch = self.scratch.data_ptr()
        self._rb = rb
        B = self.batch_size
        self._out = (torch.empty(B, self.state_size, **f32), torch.empty(B, self.action_size, **f32),
                     torch.empty(B, 1, **f32), torch.empty(B, self.state_size, **f32), torch.empty(B, 1, **f32))
        self._idx = torch.empty(B, dtype=torch.int32, device=self.device)

    @property
    def native(self) -> N.Replay:
        return self._rb

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ---- reference API ------------------------------------------------------------------------
    def add(self, state, action, reward, next_state, done) -> None:
        """One transition from host values (replay_buffer.py:32-45)."""
        s = np.asarray(state[0] if isinstance(state, tuple) else state, dtype=np.float32).reshape(1, -1)
        a = np.asarray(action, dtype=np.float32).reshape(1, -1)
        if self._rb is None:
            self._allocate(s.shape[1], a.shape[1])
        dev = self.device
        self.add_batch(torch.from_numpy(s).to(dev), torch.from_numpy(a).to(dev),
                       torch.tensor([float(reward)], dtype=torch.float32, device=dev),
                       torch.from_numpy(np.asarray(next_state, dtype=np.float32).reshape(1, -1)).to(dev),
                       torch.tensor([1 if done else 0], dtype=torch.uint8, device=dev))

    def add_batch(self, states: torch.Tensor, actions: torch.Tensor, rewards: torch.Tensor, next_states: torch.Tensor,
                  dones: torch.Tensor, valid: Optional[torch.Tensor] = None, n_valid: Optional[int] = None,
                  commit: bool = True) -> None:
        """n transitions already on the device (fp32 rows, uint8 dones / valid).  commit=False copies the rows without moving
        the cursor (rloa_replay_append_rows): the caller finishes with `commit_rows` once the ring's concurrent readers are
        done (VectorLoop runs the copy beside the update, which reads the pending rows from the step's own buffers)."""
        n = states.shape[0]
        if self._rb is None:
            self._allocate(states.shape[1], actions.shape[1])
        if n > self._max_append and valid is not None:
            self._max_append = n
            self.scratch = torch.zeros((n + 63) // 64 + 2, dtype=torch.int32, device=self.device)
            self._rb.scratch = self.scratch.data_ptr()
        fn = self.lib.rloa_replay_append if commit else self.lib.rloa_replay_append_rows
        N.check(fn(C.byref(self._rb), n, states.data_ptr(), actions.data_ptr(), rewards.data_ptr(), next_states.data_ptr(),
                   N.ptr(dones), N.ptr(valid), self._stream()), 'rloa_replay_append')
        # host mirror of the live count; with a valid mask and no n_valid it is an upper bound (exact value:
        # sync_len()) — the `len(memory) > batch_size` gate only matters before the ring first fills a batch
        added = n if (valid is None or n_valid is None) else int(n_valid)
        if valid is not None and n_valid is None:
            self._len_exact = False
        self._len = min(self.buffer_size, self._len + added)

    def commit_rows(self, n: int, valid: Optional[torch.Tensor] = None) -> None:
        """Second half of add_batch(..., commit=False): the cursor moves past the rows copied then."""
        N.check(self.lib.rloa_replay_commit(C.byref(self._rb), int(n), N.ptr(valid), self._stream()), 'rloa_replay_commit')

    def gate_open(self, pending: int = 0, pending_exact: bool = True) -> bool:
        """The `len(memory) > batch_size` gate of NAFAgent.step (naf_algorithm.py:150), evaluated on the EXACT live count:
        `pending` rows are about to be appended before the update samples; with a masked append of unknown size
        (pending_exact False) only rows already in the ring count, so learning can start one vectorised step later than
        the upper bound would allow, never earlier — the sampler never sees fewer than batch_size live rows.  Costs
        one device read the first time the host mirror crosses the threshold, nothing afterwards."""
        if self._gate_open:
            return True
        if self._len + pending <= self.batch_size:
            return False
        if not (self._len_exact and pending_exact):
            self.sync_len()
            self._len_exact = True
        live = self._len + (pending if pending_exact else 0)
        self._gate_open = live > self.batch_size
        return self._gate_open

    def sync_len(self) -> int:
        """Exact live count read back from the device cursor (synchronises)."""
        if self._rb is not None:
            self._len = min(self.buffer_size, int(self.cursor.item()))
        return self._len

    def sample_into(self, states, actions, rewards, next_states, dones, indices=None, tick=None, sub: int = 0) -> None:
        """`tick` (device uint64 counter) + `sub` replace the host draw counter when the call is captured in a CUDA graph."""
        draw = self._draws if tick is None else (int(sub) << 32)
        N.check(self.lib.rloa_replay_sample(C.byref(self._rb), states.shape[0], self.seed, draw, N.ptr(tick),
                                            states.data_ptr(), actions.data_ptr(), rewards.data_ptr(),
                                            next_states.data_ptr(), N.ptr(dones), N.ptr(indices), self._stream()),
                'rloa_replay_sample')
        self._draws += 1

    def sample(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        if self._rb is None or (self._len if self._len_exact else self.sync_len()) < self.batch_size:
            raise ValueError('Sample larger than population or is negative')     # what random.sample raises
        s, a, r, s2, d = self._out
        self.sample_into(s, a, r, s2, d, self._idx)
        return s.clone(), a.long(), r.clone(), s2.clone(), d.clone()

    def __len__(self) -> int:
        return self._len

    # ---- resume (an addition: the reference does not persist its deque) ------------------------------
    def state_dict(self) -> dict:
        """Live rows, write cursor and draw counter as CPU tensors (NAFAgent.save_training_state)."""
        if self._rb is None:
            return {'allocated': False, 'draws': self._draws}
        n = self.sync_len()
        rows = lambda t: t[:n].detach().cpu()
        return {'allocated': True, 'state_size': self.state_size, 'action_size': self.action_size,
                'buffer_size': self.buffer_size, 'len': n, 'cursor': int(self.cursor.item()), 'draws': self._draws,
                'states': rows(self.states), 'next_states': rows(self.next_states), 'actions': rows(self.actions),
                'rewards': rows(self.rewards), 'dones': rows(self.dones)}

    def load_state_dict(self, sd: dict) -> None:
        self._draws = int(sd['draws'])
        if not sd['allocated']:
            return
        if int(sd['buffer_size']) != self.buffer_size:
            raise ValueError(f"replay capacity mismatch: saved {sd['buffer_size']}, this buffer {self.buffer_size}")
        if self._rb is None:
            self._allocate(int(sd['state_size']), int(sd['action_size']))
        if (self.state_size, self.action_size) != (int(sd['state_size']), int(sd['action_size'])):
            raise ValueError('replay row layout mismatch')
        n = int(sd['len'])
        for name in ('states', 'next_states', 'actions', 'rewards', 'dones'):
            getattr(self, name)[:n].copy_(sd[name].to(self.device))
        self.cursor.fill_(int(sd['cursor']))
        self._len, self._len_exact, self._gate_open = n, True, n > self.batch_size
