"""Phase times of the training loop AS SCHEDULED (simulator pipelined, update first, forks beside it): CUDA-event nodes inside
the captured graph at the phase boundaries of configs[1], L2 warm.  Each event node costs ~2-4 us, so the sum exceeds the
un-instrumented step; what matters is where the time sits.
    python tools/phase_timeline.py [pairs=200]"""
import logging, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robotic_manipulator_rloa_b200 import ManipulatorFramework

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
mf = ManipulatorFramework(); mf.set_log_level(logging.ERROR)
mf.set_hyperparameter('batch_size', 1024)
mf.initialize_environment(manipulator_file='kuka_iiwa/kuka_with_gripper2.sdf', endeffector_index=13,
                          fixed_joints=[6, 7, 8, 9, 10, 11, 12, 13], involved_joints=[0, 1, 2, 3, 4, 5],
                          target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
                          initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0],
                          initial_positions_variation_range=[0, 0, 0, 0, 0, 0], visualize=False, n_envs=4096)
mf.initialize_naf_agent(seed=0)
a = mf.naf_agent; a.set_trunk_mode(1)
loop = a.make_loop(400, 1 << 22); loop.reset_all()
loop.frame.copy_(torch.randint(0, 400, (4096,), device='cuda', dtype=torch.int32))
loop.run_steps(40)          # eager until the replay holds a batch ...
loop.run_steps(460)         # ... then the captured pair
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(pairs):
    loop.replay_pair()
e1.record(); torch.cuda.synchronize()
print('un-instrumented: %.1f us per step' % (e0.elapsed_time(e1) / (2 * pairs) * 1e3))
ev = [[torch.cuda.Event(enable_timing=True, external=True) for _ in range(6)] for _ in range(2)]
loop._graph = None
loop.phase_events = ev
assert loop.capture(), loop.graph_error
acc = [0.0] * 6
for _ in range(pairs):
    loop.replay_pair()
    torch.cuda.synchronize()
    for par in range(2):
        for k in range(5):
            acc[k] += ev[par][k].elapsed_time(ev[par][k + 1])
        acc[5] += ev[par][0].elapsed_time(ev[par][5])
names = ('act (pack + policy kernel)', 'Environment.step (solve only: first half prepared)', 'fork points', 'update (sample + learn)',
         'joins (dynamics, bookkeeping, replay commit)', 'iteration')
for n, t in zip(names, acc):
    print('%-52s %7.1f us' % (n, t / (2 * pairs) * 1e3))
