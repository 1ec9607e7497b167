"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): a few loop iterations at 256 envs in both
trunk modes, eager launches (developer tool; slow under the sanitizer)."""
import logging
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robotic_manipulator_rloa_b200.environment.environment import Environment, EnvironmentConfiguration
from robotic_manipulator_rloa_b200.naf_components.naf_algorithm import NAFAgent

logging.getLogger().setLevel(logging.ERROR)
dev = torch.device('cuda:0')
cfg = EnvironmentConfiguration(endeffector_index=13, fixed_joints=list(range(6, 14)), involved_joints=list(range(6)),
                               target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
                               initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0],
                               initial_positions_variation_range=[0, 0, .5, .5, .5, .5], visualize=False)
for trunk in (0, 1):
    env = Environment('kuka_iiwa/kuka_with_gripper2.sdf', cfg, n_envs=256, device=dev)
    agent = NAFAgent(env, 21, 6, 256, 128, 4096, 1e-3, 1e-3, 0.99, 1, 1, 500, dev, seed=0)
    agent.set_trunk_mode(trunk)
    loop = agent.make_loop(6, 1 << 12)
    loop.reset_all()
    for _ in range(8):
        loop.step()
    torch.cuda.synchronize()
    print('trunk', trunk, 'ok: loss', float(agent.last_loss.item()), 'episodes', int(loop.log_count.item()))
