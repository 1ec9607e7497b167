"""Cost of the contact rows in the training loop's simulator step (configs[1] shape): fraction of arms flagged near / in
contact and the env_step phase time.  RLOA_CONTACT_DEBUG=1: collision phase only; =2: near flags only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, logging
from robotic_manipulator_rloa_b200.environment.environment import Environment, EnvironmentConfiguration
from robotic_manipulator_rloa_b200.naf_components.naf_algorithm import NAFAgent
logging.getLogger().setLevel(logging.ERROR)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = EnvironmentConfiguration(endeffector_index=13, fixed_joints=[6, 7, 8, 9, 10, 11, 12, 13], involved_joints=[0, 1, 2, 3, 4, 5],
                               target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
                               initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0], initial_positions_variation_range=[0] * 6, visualize=False)
env = Environment('kuka_iiwa/kuka_with_gripper2.sdf', cfg, n_envs=n, seed=0)
agent = NAFAgent(env, 21, 6, 256, 1024, 100000, 1e-3, 1e-3, 0.99, 1, 1, 500, env.device, seed=0)
agent.set_trunk_mode(1)
loop = agent.make_loop(400, 1 << 22)
loop.reset_all()
loop.frame.copy_(torch.randint(0, 400, (n,), device=env.device, dtype=torch.int32))
loop.run_steps(600)
torch.cuda.synchronize()
import ctypes as C
near = torch.empty(n, dtype=torch.int32, device=env.device)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
acts = torch.zeros(n, 6, device=env.device)
tot = 0.0
for i in range(50):
    a = agent.act_batch(loop.state)
    e0.record(); env.sim.step(a, out=(loop.next_state, loop.reward, loop.done)); e1.record(); torch.cuda.synchronize()
    loop.state, loop.next_state = loop.next_state, loop.state
    tot += e0.elapsed_time(e1)
print(f'contacts dbg={os.environ.get("RLOA_CONTACT_DEBUG", "0")} thr={env.sim.cfg.contact_threshold}: sim.step {tot / 50 * 1e3:.1f} us (3 kernels, eager)')
