"""Quick simulator throughput probe (sim-only, U(-1,1) actions); prints env-steps/s for a few N.
usage: bench_sim_quick.py [panda] [mesh[=VERTS]] [N ...]   (mesh: every collision primitive as a VERTS-point hull, GJK path)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from robotic_manipulator_rloa_b200.environment.robot_model import load_manipulator
from robotic_manipulator_rloa_b200.environment.simulator import BatchedSimulator

args = sys.argv[1:]
panda = bool(args) and args[0] == 'panda'
if panda:
    args = args[1:]
mesh = 0
if args and args[0].startswith('mesh'):
    mesh = int(args[0].split('=')[1]) if '=' in args[0] else 40
    args = args[1:]
model = load_manipulator('franka_panda/panda.urdf' if panda else 'kuka_iiwa/kuka_with_gripper2.sdf')
na = 7 if panda else 6
if mesh:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
    from helpers import hullified
    model = hullified(model, n=mesh)
    print(f'collision shapes: convex hulls, {model.verts.shape[0]} vertices in {model.ns} shapes')
print('model:', 'panda (12 joints / 9 dof, BASELINE configs[4])' if panda else 'kuka iiwa + gripper (14 joints / 12 dof)')
for n in [int(x) for x in (args or ['4096', '16384', '65536', '262144'])]:
    if panda:
        sim = BatchedSimulator(model, n, 11, list(range(7)), [7, 8, 9, 10, 11], contacts=os.environ.get('RLOA_CONTACTS', '1') != '0')
        sim.set_task([0.4, 0.3, 0.5], [0.3, 0.0, 0.6])
        base = torch.tensor([0, 0, 0, -1.5, 0, 1.5, 0], device='cuda', dtype=torch.float32)
    else:
        sim = BatchedSimulator(model, n, 13, [0, 1, 2, 3, 4, 5], list(range(6, 14)), contacts=os.environ.get('RLOA_CONTACTS', '1') != '0')
        sim.set_task([0.4, 0.85, 0.71], [0.45, 0.55, 0.55])
        base = torch.tensor([0.9, 0.45, 0, 0, 0, 0], device='cuda')
    start = base + 0.5 * (torch.rand(n, na, device='cuda') - 0.5)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); sim.reset(start); t1.record(); torch.cuda.synchronize()
    print(f'N={n}: reset (50 substeps) {t0.elapsed_time(t1):.3f} ms -> {n*50/t0.elapsed_time(t1)*1e3:.3e} sim-steps/s')
    acts = [2 * torch.rand(n, na, device='cuda') - 1 for _ in range(8)]
    for i in range(5):
        sim.step(acts[i % 8])
    torch.cuda.synchronize()
    K = 50
    t0.record()
    for i in range(K):
        sim.step(acts[i % 8])
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / K
    print(f'N={n}: step {ms*1e3:.1f} us -> {n/ms*1e3:.3e} env-steps/s; mean PGS iters {sim.last_iterations().float().mean().item():.1f}')
    sim.close()
