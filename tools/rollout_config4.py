"""BASELINE.json configs[3]: test_trained_model rollouts with per-env randomised target / obstacle positions on the
demo KUKA weights (weights_kuka.p), through the public ManipulatorFramework API.

    python tools/rollout_config4.py [n_envs=16384] [frames=750]

target = [0.4, 0.85, 0.71] + U(-0.1, 0.1)^3, obstacle = [0.45, 0.55, 0.55] + U(-0.1, 0.1)^3 per env (seed 4321),
start pose with var [0, 0, .5, .5, .5, .5] (rl_framework.py:648-649).  Reports success %, collision %, mean frames
(the quantities of rl_framework.py:362-367) and rollout throughput."""
import logging
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from robotic_manipulator_rloa_b200 import ManipulatorFramework

n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 750
logging.getLogger().setLevel(logging.ERROR)
mf = ManipulatorFramework()
mf.initialize_environment(manipulator_file='kuka_iiwa/kuka_with_gripper2.sdf', endeffector_index=13,
                          fixed_joints=[6, 7, 8, 9, 10, 11, 12, 13], involved_joints=[0, 1, 2, 3, 4, 5],
                          target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
                          initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0],
                          initial_positions_variation_range=[0, 0, .5, .5, .5, .5], visualize=False, n_envs=n_envs)
mf.initialize_naf_agent()
here = os.path.dirname(os.path.abspath(__file__))
mf.load_pretrained_parameters_from_weights_file(
    os.path.join(here, '..', 'robotic_manipulator_rloa_b200', 'naf_components', 'demo_weights', 'weights_kuka.p'))
mf.naf_agent.set_trunk_mode(1)
g = torch.Generator().manual_seed(4321)
dev = mf.env.device
tgt = torch.tensor([0.4, 0.85, 0.71]) + 0.2 * (torch.rand(n_envs, 3, generator=g) - 0.5)
obs = torch.tensor([0.45, 0.55, 0.55]) + 0.2 * (torch.rand(n_envs, 3, generator=g) - 0.5)
mf.env.set_task_positions(tgt.to(dev), obs.to(dev))
torch.cuda.synchronize()
t0 = time.perf_counter()
mf.test_trained_model(n_envs, frames)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
res = mf.last_test_results
ok = np.array([r[0] for r in res]); fr = np.array([r[1] for r in res])
coll = sum(1 for r in res if (not r[0]) and r[1] < frames - 1)
print(f'config 4: {len(res)} episodes x <= {frames} frames in {dt:.2f} s; success {100 * ok.mean():.1f} %, '
      f'collisions {100 * coll / len(res):.1f} %, mean frames of successes {fr[ok].mean() if ok.any() else float("nan"):.1f}; '
      f'~{(fr + 1).sum() / dt:.3e} env-steps/s (wall clock, incl. the 50-sub-step resets)')
