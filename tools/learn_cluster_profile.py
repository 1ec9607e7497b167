"""Phase timeline of the fused learn kernel (csrc/naf_learn_cluster.cu) from its clock64 stamps.
    python tools/learn_cluster_profile.py [batch=1024]
Prints, per phase, the cycles between consecutive stamps of thread 0 (median over the 8 CTAs of each cluster) in us at the
SM clock read from nvidia-smi.  The stamps cost a few cycles each (one store by thread 0)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from robotic_manipulator_rloa_b200 import _native as N
from test_naf_learn_cluster_gpu import make_agent, make_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1024
if int(os.environ.get('WORLD_SIZE', '1')) > 1:      # under torchrun: the in-kernel NVLink gradient exchange is part of the tail
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
FLUSH = 'flush' in sys.argv          # 256 MiB fill before the stamped launch: cold L2, as between bench.py's timed pairs
import test_naf_learn_cluster_gpu as T
T.DEV = torch.device('cuda', torch.cuda.current_device())
agent, _, _ = make_agent(batch=B)
s, a, r, s2, d = make_batch(B, seed=1)
f = lambda t: t.to(device=T.DEV, dtype=torch.float32).contiguous()
args = (f(s), f(a.long()), f(r).reshape(-1), f(s2), f(d).reshape(-1))
for _ in range(5):
    agent._learn_device(*args)
stamps = torch.zeros(16, 32, dtype=torch.int64, device=T.DEV)
N.check(agent._ws.lib.rloa_naf_ws_set_debug(agent._ws.handle, None, stamps.data_ptr()), 'dbg')
for _ in range(3):
    agent._learn_device(*args)
if FLUSH:
    torch.empty(256 << 20, dtype=torch.uint8, device='cuda').fill_(1)
    agent._learn_device(*args)
torch.cuda.synchronize()
N.check(agent._ws.lib.rloa_naf_ws_set_debug(agent._ws.handle, None, None), 'dbg')
st = stamps.cpu().numpy()
try:
    mhz = float(subprocess.check_output(['nvidia-smi', '--query-gpu=clocks.sm', '--format=csv,noheader,nounits']).split()[0])
except Exception:
    mhz = 1965.0
if int(os.environ.get('RANK', '0')) != 0:
    torch.cuda.synchronize()
    for _ in range(50):
        agent._learn_device(*args)
    torch.cuda.synchronize()
    dist.barrier(); dist.destroy_process_group(); sys.exit(0)
names = {0: 'start', 1: 'prologue+L1 mma', 2: 'BN1 stats+xchg', 3: 'a1 tile+L2 mma', 4: 'BN2 stats+xchg', 5: 'a2 tile+head mma', 6: 'head read + wait y',
         7: 'head math/dzh/hb', 8: 'dWh mma+dump', 9: 'da2 mma', 10: 'BN2 bwd stats+xchg', 11: 'dz2 tile', 12: 'da1 mma+x restage', 13: 'z1 mma',
         14: 'BN1 bwd stats+xchg', 15: 'dz1/a1 tiles+x tile', 16: 'dW1 mma+dump', 17: 'dW2 mma+staged dump', 18: 'cluster sync (partials)',
         19: 'gather partials', 20: 'norm xchg', 21: 'adam', 22: 'final cluster sync'}
for net, label in ((1, 'main'), (0, 'target')):
    rows = st[net * 8:(net + 1) * 8]
    print(f'--- {label} cluster (8 CTAs), us at {mhz:.0f} MHz: median [min..max] ---')
    prev = 0
    for i in range(1, 23):
        if (rows[:, i] == 0).all():
            continue
        dt = (rows[:, i] - rows[:, prev]) / mhz
        print(f'{names.get(i, i):28s} {np.median(dt):7.2f} [{dt.min():6.2f} .. {dt.max():6.2f}]')
        prev = i
    tot = (rows[:, prev] - rows[:, 0]) / mhz
    print(f'{"total":28s} {np.median(tot):7.2f}')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    agent._learn_device(*args)
e1.record(); torch.cuda.synchronize()
print(f'eager learn (pack + cluster kernel), no debug: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per update')

if int(os.environ.get('WORLD_SIZE', '1')) > 1:
    dist.barrier(); dist.destroy_process_group()
