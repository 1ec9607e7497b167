"""NAFAgent.learn timing on one GPU: eager vs CUDA graph, warm vs L2-flushed, cluster kernel vs multi-launch path.
    python tools/learn_timing.py [batch=1024]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from test_naf_learn_cluster_gpu import make_agent, make_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
agent, _, _ = make_agent(batch=B)
s, a, r, s2, d = make_batch(B, seed=1)
f = lambda t: t.to(device='cuda', dtype=torch.float32).contiguous()
args = (f(s), f(a.long()), f(r).reshape(-1), f(s2), f(d).reshape(-1))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, n=50, flushed=False):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    if not flushed:
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    tot = 0.0
    for i in range(n):
        flush.fill_(i & 255)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


fn = lambda: agent._learn_device(*args)
print(f'eager, warm L2:    {timed(fn):7.1f} us')
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    fn()
print(f'graph, warm L2:    {timed(g.replay):7.1f} us')
print(f'graph, flushed L2: {timed(g.replay, flushed=True):7.1f} us')
print(f'eager, flushed L2: {timed(fn, flushed=True):7.1f} us')
