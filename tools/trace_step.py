"""Warm per-launch times of the vectorised loop (developer tool): RLOA_TRACE=1 python tools/trace_step.py [fp32|tc]
Each line is the mean interval that ENDS at a launch site (file:line of the RLOA_LAUNCHED after the <<<>>>), i.e. that
kernel's execution plus the gap before it, measured with CUDA events in eager mode on the real loop."""
import ctypes as C
import os
import sys
os.environ['RLOA_TRACE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import logging
import torch
from robotic_manipulator_rloa_b200 import _native
from robotic_manipulator_rloa_b200.environment.environment import Environment, EnvironmentConfiguration
from robotic_manipulator_rloa_b200.naf_components.naf_algorithm import NAFAgent

logging.getLogger().setLevel(logging.ERROR)
dev = torch.device('cuda:0')
cfg = EnvironmentConfiguration(endeffector_index=13, fixed_joints=list(range(6, 14)), involved_joints=list(range(6)),
                               target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
                               initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0],
                               initial_positions_variation_range=[0, 0, 0, 0, 0, 0], visualize=False)
env = Environment('kuka_iiwa/kuka_with_gripper2.sdf', cfg, n_envs=4096, device=dev)
agent = NAFAgent(env, 21, 6, 256, 1024, 100000, 1e-3, 1e-3, 0.99, 1, 1, 500, dev, seed=0)
if (sys.argv[1:] or ['tc'])[0] == 'tc':
    agent.set_trunk_mode(1)
loop = agent.make_loop(400, 1 << 20)
loop.reset_all()
lib = C.CDLL(_native.LIB_PATH)
lib.rloa_trace_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
for _ in range(10):
    loop.step()
torch.cuda.synchronize()
lib.rloa_trace_dump(0)          # discard warm-up
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(40):
    if i % 2 == 0:
        flush.fill_(i)
    loop.step()
lib.rloa_trace_dump(0)
