"""Where the host-facing step (VectorLoop.step_host, bench.py's e2e) spends its time: host wall clock around the two graph
launches and the two waits, 200 steps of configs[1]."""
import logging, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robotic_manipulator_rloa_b200 import ManipulatorFramework

mf = ManipulatorFramework(); mf.set_log_level(logging.ERROR)
mf.set_hyperparameter('batch_size', 1024)
mf.initialize_environment(manipulator_file='kuka_iiwa/kuka_with_gripper2.sdf', endeffector_index=13,
                          fixed_joints=[6, 7, 8, 9, 10, 11, 12, 13], involved_joints=[0, 1, 2, 3, 4, 5],
                          target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
                          initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0],
                          initial_positions_variation_range=[0, 0, 0, 0, 0, 0], visualize=False, n_envs=4096)
mf.initialize_naf_agent(seed=0)
a = mf.naf_agent; a.set_trunk_mode(1)
loop = a.make_loop(400, 1 << 20); loop.reset_all()
loop.frame.copy_(torch.randint(0, 400, (4096,), device='cuda', dtype=torch.int32))      # episodes out of phase, as in bench.py
loop.run_steps(500)
n = 4096
hs, ha = torch.zeros(n, 21).pin_memory(), torch.zeros(n, 6).pin_memory()
hr, hd = torch.zeros(n).pin_memory(), torch.zeros(n, dtype=torch.uint8).pin_memory()
hs.copy_(loop.state)
loop.bind_host_buffers(hs, ha, hr, hd)
for _ in range(5):
    loop.step_host()
torch.cuda.synchronize()
g1, g2 = loop._host_graphs
stream = torch.cuda.current_stream()
T = [0.0] * 4
K = 300
t_all = time.perf_counter()
for _ in range(K):
    t0 = time.perf_counter(); g1.replay()
    t1 = time.perf_counter(); stream.synchronize()
    t2 = time.perf_counter(); g2.replay()
    t3 = time.perf_counter(); loop._host_ready.synchronize()
    t4 = time.perf_counter()
    T[0] += t1 - t0; T[1] += t2 - t1; T[2] += t3 - t2; T[3] += t4 - t3
t_all = time.perf_counter() - t_all
torch.cuda.synchronize()
loop.transitions.zero_() if False else None
print('per step us: launch g1 %.1f | wait act+D2H actions %.1f | launch g2 %.1f | wait step+D2H results %.1f | total %.1f' % (
    *(x / K * 1e6 for x in T), t_all / K * 1e6))
t_all = time.perf_counter()
for _ in range(K):
    loop.step_host()
t_all = time.perf_counter() - t_all
torch.cuda.synchronize()
print('through VectorLoop.step_host: %.1f us per step' % (t_all / K * 1e6))
