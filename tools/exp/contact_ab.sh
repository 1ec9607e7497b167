for dbg in 0 1 2; do
echo "contact_debug=$dbg"
RLOA_CONTACT_DEBUG=$dbg python bench.py --steps 200 --warmup 20 --no-cpu --no-extras --no-flush | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], d['phases_ms'], d['roofline']['at_measured_sweeps']['mean_pgs_sweeps'])"
done
python - <<'PY'
import torch, logging
from robotic_manipulator_rloa_b200 import ManipulatorFramework
mf = ManipulatorFramework(); mf.set_log_level(logging.ERROR)
mf.set_hyperparameter('batch_size', 1024)
mf.initialize_environment(manipulator_file='kuka_iiwa/kuka_with_gripper2.sdf', endeffector_index=13,
    fixed_joints=[6,7,8,9,10,11,12,13], involved_joints=[0,1,2,3,4,5], target_position=[0.4,0.85,0.71],
    obstacle_position=[0.45,0.55,0.55], initial_joint_positions=[0.9,0.45,0,0,0,0],
    initial_positions_variation_range=[0,0,0,0,0,0], visualize=False, n_envs=4096)
mf.initialize_naf_agent(seed=0)
a = mf.naf_agent; a.set_trunk_mode(1)
loop = a.make_loop(400, 1 << 20); loop.reset_all()
sim = mf.env.sim
import ctypes
for k in range(12):
    loop.run_steps(50); torch.cuda.synchronize()
    near = sim.contact_counts()
    print(k, None if near is None else (float((near>0).float().mean()), float((near.view(-1,32)>0).any(1).float().mean()), int(near.max())))
PY
