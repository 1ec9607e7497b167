"""BASELINE.json configs[0]: the reference's own kuka_training demo (1 env, NAF batch 128) through the drop-in facade,
timed.  python tools/demo_config1_timing.py [episodes=10] [frames=400]"""
import logging
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from robotic_manipulator_rloa_b200 import ManipulatorFramework

episodes = int(sys.argv[1]) if len(sys.argv) > 1 else 10
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 400
os.chdir(tempfile.mkdtemp())
mf = ManipulatorFramework()
mf.set_log_level(logging.ERROR)
mf.initialize_environment(manipulator_file='kuka_iiwa/kuka_with_gripper2.sdf', endeffector_index=13,
                          fixed_joints=[6, 7, 8, 9, 10, 11, 12, 13], involved_joints=[0, 1, 2, 3, 4, 5],
                          target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
                          initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0],
                          initial_positions_variation_range=[0, 0, 0, 0, 0, 0], visualize=False)
mf.initialize_naf_agent()
torch.cuda.synchronize()
t0 = time.perf_counter()
scores = mf.run_training(episodes, frames, verbose=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
steps = sum(v[1] for v in scores.values())
print(f'configs[0]: {episodes} episodes, {steps} env-steps (+ {episodes * 50} reset sub-steps), one NAF update (batch 128) per step '
      f'once the buffer holds a batch: {dt:.2f} s wall = {dt / max(steps, 1) * 1e3:.3f} ms per step, {steps / dt:.0f} env-steps/s')
