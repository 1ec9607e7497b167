import logging, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robotic_manipulator_rloa_b200 import ManipulatorFramework
logging.getLogger().setLevel(logging.ERROR)
mf = ManipulatorFramework()
mf.set_log_level(logging.ERROR)           # 4 log lines per finished episode otherwise (the reference's format)
mf.set_hyperparameter('batch_size', 1024)
mf.initialize_environment(manipulator_file='kuka_iiwa/kuka_with_gripper2.sdf', endeffector_index=13,
                          fixed_joints=[6, 7, 8, 9, 10, 11, 12, 13], involved_joints=[0, 1, 2, 3, 4, 5],
                          target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
                          initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0],
                          initial_positions_variation_range=[0, 0, .5, .5, .5, .5], visualize=False, n_envs=4096)
mf.initialize_naf_agent(checkpoint_frequency=10 ** 9)     # the reference's default (500 episodes) would write 81 checkpoints here
mf.naf_agent.set_trunk_mode(1)
os.chdir(os.environ.get('TMPDIR', '/tmp'))
torch.cuda.synchronize(); t0 = time.perf_counter()
scores = mf.run_training(40960, 400, verbose=False)      # ~10 episodes per env ~ 4500 iterations
torch.cuda.synchronize(); dt = time.perf_counter() - t0
frames = sum(v[1] for v in scores.values())
print(f'run_training(40960 episodes x 400 frames, 4096 envs): {dt:.2f} s wall, {frames} logged env-steps -> {frames / dt:.3e} env-steps/s through the public API')
