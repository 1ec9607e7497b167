"""
CPU tests of the N > 1 path with world_size-2 `gloo` process groups: env sharding, per-rank
seeding, and the one exchange step of the data-parallel NAF update — all-reduce of the flat gradient, then the
identical clip + Adam + soft update on every rank.  The arithmetic of the update is the oracle's (this is the
host-side protocol under test, not the CUDA kernels; those are covered by the -m gpu tests).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import assert_params_close
from oracle.naf_restatement import NAFRef
from robotic_manipulator_rloa_b200.utils import distributed as rdist

S, A, H, B = 21, 6, 64, 96


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_shard_range_is_a_partition():
    for n, w in [(4096, 1), (4096, 8), (65536, 8), (10, 4), (3, 8), (1048576, 8)]:
        seen = []
        for r in range(w):
            lo, hi = rdist.shard_range(n, r, w)
            assert 0 <= lo <= hi <= n
            seen.extend(range(lo, hi)) if n <= 4096 else seen.append((lo, hi))
        if n <= 4096:
            assert seen == list(range(n))
        else:
            assert seen[0][0] == 0 and seen[-1][1] == n and all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
        sizes = [rdist.shard_range(n, r, w)[1] - rdist.shard_range(n, r, w)[0] for r in range(w)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        rdist.shard_range(8, 2, 2)
    assert rdist.world() == (0, 1)
    assert len({rdist.rank_seed(0, r, k) for r in range(8) for k in range(3)}) == 24


def _batch(rank):
    g = torch.Generator().manual_seed(100 + rank)
    s = torch.randn(B, S, generator=g)
    return (s, torch.clamp(torch.randn(B, A, generator=g) * 1.5, -1, 1).long(), -torch.rand(B, 1, generator=g),
            s + 0.1 * torch.randn(B, S, generator=g), torch.zeros(B, 1))


def _local_gradient(main, target, batch, gamma=0.99):
    """Flat gradient of one NAFAgent.learn on this rank's replay shard (naf_algorithm.py:194-208)."""
    s, a, r, s2, _ = batch
    main.train(); target.train()
    main.zero_grad()
    with torch.no_grad():
        v_next = target.heads(s2)[3]
    q = main.heads(s, a)[2]
    torch.nn.functional.mse_loss(q, r + gamma * v_next).backward()
    return torch.cat([p.grad.reshape(-1) for p in main.parameters()]).clone()


def _apply(main, target, opt, flat, scale, tau=1e-3):
    """clip_grad_norm_(1) + Adam + soft update from the (summed) flat gradient times `scale` (naf_algorithm.py:209-226)."""
    off = 0
    for p in main.parameters():
        p.grad = (flat[off:off + p.numel()] * scale).reshape(p.shape).clone()
        off += p.numel()
    torch.nn.utils.clip_grad_norm_(main.parameters(), 1)
    opt.step()
    with torch.no_grad():
        for pt, pm in zip(target.parameters(), main.parameters()):
            pt.copy_(tau * pm + (1.0 - tau) * pt)


def _worker(rank, world_size, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    torch.set_num_threads(1)
    assert rdist.world() == (rank, world_size)
    lo, hi = rdist.shard_range(4096, rank, world_size)
    assert hi - lo == 4096 // world_size
    main, target = NAFRef(S, A, H, seed=0), NAFRef(S, A, H, seed=0)       # same seed on every rank: no broadcast
    opt = torch.optim.Adam(main.parameters(), lr=1e-3)
    for step in range(3):
        flat = _local_gradient(main, target, _batch(rank + 10 * step))
        scale = rdist.allreduce_gradient(flat)
        assert scale == 1.0 / world_size
        _apply(main, target, opt, flat, scale)
    vec = torch.cat([p.detach().reshape(-1) for p in list(main.parameters()) + list(target.parameters())])
    gathered = [torch.zeros_like(vec) for _ in range(world_size)]
    dist.all_gather(gathered, vec)
    if rank == 0:
        np.save(os.path.join(out_dir, 'params.npy'), torch.stack(gathered).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_gradient_allreduce_keeps_ranks_identical(tmp_path):
    world_size = 2
    mp.spawn(_worker, args=(world_size, _free_port(), str(tmp_path)), nprocs=world_size, join=True)
    params = np.load(tmp_path / 'params.npy')
    assert np.array_equal(params[0], params[1]), 'ranks diverged: parameters must stay bit-identical without a broadcast'
    # single-process restatement of the same protocol: both shards' gradients, averaged
    main, target = NAFRef(S, A, H, seed=0), NAFRef(S, A, H, seed=0)
    opt = torch.optim.Adam(main.parameters(), lr=1e-3)
    for step in range(3):
        flat = sum(_local_gradient(main, target, _batch(r + 10 * step)) for r in range(world_size))
        _apply(main, target, opt, flat, 1.0 / world_size)
    want = torch.cat([p.detach().reshape(-1) for p in list(main.parameters()) + list(target.parameters())]).numpy()
    # Elements whose gradient is rounding noise (the linear biases in front of a BatchNorm) take lr-sized Adam steps in
    # an arbitrary direction, in the workers and here alike: allow 2 % of elements to differ by up to 3 steps of lr.
    assert_params_close(params[0], want, 'params after 3 data-parallel updates', max_frac=0.02)
