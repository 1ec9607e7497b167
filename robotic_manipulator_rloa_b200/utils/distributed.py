"""
Host-side plumbing of the data-parallel path: one process per GPU, env instances sharded by
rank, no data-path collective in the rollout, and ONE exchange per NAF update — the all-reduce of the flat main-net
gradient between rloa_naf_learn_grads and rloa_naf_learn_apply.  The reference is single-process
(naf_components/naf_algorithm.py:180-213); these helpers define what "the same update on every rank" means.

Backend-agnostic on purpose: NCCL on the GPUs, gloo in the CPU tests (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Env index range [lo, hi) owned by `rank`: contiguous, disjoint, covering, sizes differing by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError(f'rank {rank} outside world of {world_size}')
    base, extra = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_seed(seed: int, rank: int, stream: int = 0) -> int:
    """Per-rank seed for exploration noise / replay sampling / start poses (stream 0, 1, 2).  Network
    initialisation does NOT use this: every rank seeds the NAF weights with the same `seed`, so that parameters
    start identical and stay identical without a broadcast."""
    return (int(seed) + 1000003 * (int(rank) + 1) + 7919 * int(stream)) & 0x7FFFFFFFFFFFFFFF


def allreduce_gradient(flat_grad: torch.Tensor) -> float:
    """Sum the flat gradient over ranks in place; returns the scale (1 / world_size) the optimiser must apply
    (rloa_naf_hyper.grad_scale), i.e. the ranks step with the MEAN gradient — clip norm and Adam update are
    then computed from identical numbers on every rank."""
    _, w = world()
    if w > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / float(w)


class GradExchange:
    """NVLink peer-memory gradient exchange (csrc/grad_exchange.cu): one exchange block per rank, exported with CUDA
    IPC, mapped by every peer; the optimiser's own kernels publish, sum (in rank order) and consume the gradient.
    Construction is collective: every rank of the process group must create it at the same point.  `create`
    returns None on EVERY rank when any rank cannot export or map the blocks (IPC disabled, no peer access), so the
    callers fall back to the NCCL all-reduce together."""

    def __init__(self, n_params: int, device: torch.device):
        import ctypes as C
        from .. import _native as N
        self.lib = N.lib()
        self.device = torch.device(device)
        self.rank, self.world = world()
        self._h = C.c_void_p()
        self.error = None
        with torch.cuda.device(self.device):
            mine = C.create_string_buffer(64)
            ok = True
            try:
                N.check(self.lib.rloa_xchg_create(int(n_params), C.byref(self._h)), 'rloa_xchg_create')
                N.check(self.lib.rloa_xchg_handle(self._h, mine), 'rloa_xchg_handle')
            except N.NativeLibraryError as err:
                ok, self.error = False, str(err)
            if not self._all_ok(ok):
                self.close()
                return
            handles = None
            if self.world > 1:
                local = torch.tensor(list(mine.raw), dtype=torch.uint8, device=self.device)
                gathered = [torch.empty_like(local) for _ in range(self.world)]
                dist.all_gather(gathered, local)
                handles = bytes(torch.cat(gathered).cpu().numpy().tobytes())
            try:
                N.check(self.lib.rloa_xchg_connect(self._h, self.rank, self.world, handles), 'rloa_xchg_connect')
            except N.NativeLibraryError as err:
                ok, self.error = False, str(err)
            if not self._all_ok(ok):
                self.close()
                return
            if self.world > 1:
                dist.barrier()

    def _all_ok(self, ok: bool) -> bool:
        """Logical AND of `ok` over the ranks (one tiny all-reduce), so that every rank takes the same branch."""
        if self.world == 1:
            return ok
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item())

    @classmethod
    def create(cls, n_params: int, device: torch.device):
        x = cls(n_params, device)
        return x if x._h else None

    @property
    def handle(self):
        return self._h

    def timed_out(self) -> bool:
        """True after a device-side wait for a peer gave up (~2 s): the ranks have diverged."""
        return self.lib.rloa_xchg_status(self._h) == 1

    def close(self) -> None:
        if self._h:
            self.lib.rloa_xchg_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
