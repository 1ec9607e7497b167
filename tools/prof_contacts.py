"""The training loop of bench.py's headline config, eager launches, for ncu: the solve kernel's duration as contacts appear
towards the end of the first 400-frame episode (prints the fraction of arms / warps that carry contact rows)."""
import logging
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robotic_manipulator_rloa_b200 import ManipulatorFramework

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 350
graph = len(sys.argv) > 2 and sys.argv[2] == 'graph'       # timed chunks through the CUDA graph instead of eager launches
mf = ManipulatorFramework()
mf.set_log_level(logging.ERROR)
mf.set_hyperparameter('batch_size', 1024)
mf.initialize_environment(manipulator_file='kuka_iiwa/kuka_with_gripper2.sdf', endeffector_index=13,
                          fixed_joints=[6, 7, 8, 9, 10, 11, 12, 13], involved_joints=[0, 1, 2, 3, 4, 5],
                          target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
                          initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0],
                          initial_positions_variation_range=[0, 0, 0, 0, 0, 0], visualize=False, n_envs=4096)
mf.initialize_naf_agent(seed=0)
a = mf.naf_agent
a.set_trunk_mode(1)
loop = a.make_loop(400, 1 << 20)
loop.reset_all()
sim = mf.env.sim
done = 0
while done < steps:
    k = min(50 if graph else 25, steps - done)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop.run_steps(k, use_graph=graph)
    e1.record()
    done += k
    torch.cuda.synchronize()
    if graph:
        print('%.1f us/step ' % (e0.elapsed_time(e1) / k * 1e3), end='')
    near = sim.contact_counts()
    print(done, 'arms with rows %.4f  warps with rows %.3f  max rows %d  mean sweeps %.1f' % (
        float((near > 0).float().mean()), float((near.view(-1, 32) > 0).any(1).float().mean()), int(near.max()),
        sim.last_iterations().float().mean().item()), flush=True)
