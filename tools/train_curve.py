"""Learning evidence on the device-resident loop (BASELINE.json north_star: "a NAF agent that learns the
obstacle-avoidance task end to end on device").

    python tools/train_curve.py [n_envs=4096] [iterations=40000] [chunk=2000] [batch=1024] [frames=400]

KUKA demo task (rl_framework.py:547-555: target [0.4, 0.85, 0.71], obstacle [0.45, 0.55, 0.55], start
[0.9, 0.45, 0, 0, 0, 0]), reference hyper-parameters, one NAF update per vectorised step, CUDA-graph loop.  Every
`chunk` iterations prints the episodes that finished in the window: mean return, % reaching the target (last reward
250), % hitting the obstacle (last reward -1000), mean length; ends with greedy (noise-free) rollouts."""
import logging
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from robotic_manipulator_rloa_b200 import ManipulatorFramework

arg = lambda i, d: int(sys.argv[i]) if len(sys.argv) > i else d
n_envs, iters, chunk, batch, frames = arg(1, 4096), arg(2, 40000), arg(3, 2000), arg(4, 1024), arg(5, 400)
logging.getLogger().setLevel(logging.ERROR)
mf = ManipulatorFramework()
mf.set_hyperparameter('batch_size', batch)
mf.set_hyperparameter('buffer_size', max(100000, 16 * n_envs))
mf.initialize_environment(manipulator_file='kuka_iiwa/kuka_with_gripper2.sdf', endeffector_index=13,
                          fixed_joints=[6, 7, 8, 9, 10, 11, 12, 13], involved_joints=[0, 1, 2, 3, 4, 5],
                          target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
                          initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0],
                          initial_positions_variation_range=[0, 0, 0, 0, 0, 0], visualize=False, n_envs=n_envs)
mf.initialize_naf_agent()
agent = mf.naf_agent
agent.set_trunk_mode(1)
loop = agent.make_loop(frames, 4 * n_envs + chunk * n_envs // 20 + 16)
loop.reset_all()
print(f'{n_envs} envs, batch {batch}, {frames}-step episodes, {iters} iterations (= NAF updates), chunks of {chunk}')
print('iteration | episodes | mean return | target % | obstacle % | timeout % | mean frames | loss | s')
t0 = time.perf_counter()
done_iters = 0
while done_iters < iters:
    loop.log_count.zero_()
    loop.run_steps(chunk)
    torch.cuda.synchronize()
    done_iters += chunk
    k = min(int(loop.log_count.item()), loop.cap)
    sc, fr, last = loop.log_score[:k].cpu().numpy(), loop.log_frame[:k].cpu().numpy(), loop.log_last[:k].cpu().numpy()
    if k:
        hit_t, hit_o = (last == 250).mean() * 100, (last == -1000).mean() * 100
        print(f'{done_iters:9d} | {k:8d} | {sc.mean():11.2f} | {hit_t:8.2f} | {hit_o:10.2f} | {100 - hit_t - hit_o:9.2f} | '
              f'{fr.mean():11.1f} | {float(agent.last_loss.item()):.4g} | {time.perf_counter() - t0:.1f}', flush=True)
    else:
        print(f'{done_iters:9d} | no episode finished in this window', flush=True)
# greedy evaluation: the same envs, noise off
agent.noise_scale = 0.0
ev = agent.make_loop(frames, 4 * n_envs, learn=False)
ev.reset_all()
for _ in range(frames + 60):
    ev.step()
torch.cuda.synchronize()
k = min(int(ev.log_count.item()), ev.cap)
last, fr = ev.log_last[:k].cpu().numpy(), ev.log_frame[:k].cpu().numpy()
print(f'greedy policy after training: {k} episodes, target {100 * (last == 250).mean():.2f} %, obstacle '
      f'{100 * (last == -1000).mean():.2f} %, mean frames {fr.mean():.1f}')
