"""Short simulator-only run for ncu: a few Environment.step launches at N envs (U(-1,1) actions)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robotic_manipulator_rloa_b200.environment.robot_model import load_manipulator
from robotic_manipulator_rloa_b200.environment.simulator import BatchedSimulator

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
model = load_manipulator('kuka_iiwa/kuka_with_gripper2.sdf')
if len(sys.argv) > 3:       # mesh=<verts>: every collision primitive as a convex vertex cloud (the GJK instantiation)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
    from helpers import hullified
    model = hullified(model, n=int(sys.argv[3].split('=')[1]))
sim = BatchedSimulator(model, n, 13, [0, 1, 2, 3, 4, 5], list(range(6, 14)))
sim.set_task([0.4, 0.85, 0.71], [0.45, 0.55, 0.55])
q = torch.zeros(n, model.nl, device='cuda')
q[:, :6] = torch.tensor([0.9, 0.45, 0, 0, 0, 0], device='cuda') + 0.5 * (torch.rand(n, 6, device='cuda') - 0.5)
sim.set_state(q, torch.zeros_like(q))
for i in range(steps):
    sim.step(2 * torch.rand(n, 6, device='cuda') - 1)
torch.cuda.synchronize()
print('mean PGS iters', sim.last_iterations().float().mean().item())
