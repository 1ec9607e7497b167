import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from helpers import KUKA, make_oracle, random_states, step_motors
from robotic_manipulator_rloa_b200.environment.simulator import BatchedSimulator
np.set_printoptions(linewidth=200, precision=3)
cfg = KUKA
model, orc = make_oracle(cfg)
n = 4096
q, qd = random_states(model, n, seed=7, vel=2.0)
rng = np.random.default_rng(11)
actions = rng.uniform(-1, 1, (n, 6)); actions[: n // 8] = np.sign(actions[: n // 8])
sim = BatchedSimulator(model, n, cfg['ee'], cfg['involved'], cfg['fixed']); sim.set_task(cfg['target'], cfg['obstacle'])
sim.set_state(q, qd)
sim.step(torch.as_tensor(actions, dtype=torch.float32, device='cuda'))
qg, qdg = sim.get_state(); it_g = sim.last_iterations().cpu().numpy()
step_motors(orc, cfg)
q32, qd32 = q.astype(np.float32).astype(np.float64), qd.astype(np.float32).astype(np.float64)
q0, qd0 = q32.copy(), qd32.copy()
obs_o, rew_o, done_o, it_o = orc.batch_step(q32, qd32, actions.astype(np.float32).astype(np.float64), cfg['involved'], 200.0, cfg['obstacle'], cfg['target'], nthreads=8)
err = np.abs(qdg.cpu().numpy() - qd32)
print('per-joint max |dqd|:', err.max(axis=0))
print('per-joint 99.9pct |dqd|:', np.quantile(err, 0.999, axis=0))
print('per-joint median |dqd|:', np.median(err, axis=0))
same = it_g == it_o
print('iter mismatch', (~same).sum(), 'iters hist', np.bincount(it_o)[-10:])
w = np.argsort(-err.max(axis=1))[:5]
for e in w:
    print('env', e, 'it', it_g[e], it_o[e], 'err', err[e], 'qd_o', qd32[e], 'qd_in', qd0[e])
# single substep with no PGS influence: compare free acceleration through a zero-impulse motor table
from helpers import PANDA
cfg = PANDA
model, orc = make_oracle(cfg)
q, qd = random_states(model, n, seed=7, vel=1.0, near_limit=0.25, held=cfg['fixed'])
rng = np.random.default_rng(11)
actions = rng.uniform(-1, 1, (n, 7)); actions[: n // 8] = np.sign(actions[: n // 8])
sim = BatchedSimulator(model, n, cfg['ee'], cfg['involved'], cfg['fixed']); sim.set_task(cfg['target'], cfg['obstacle'])
sim.set_state(q, qd)
sim.step(torch.as_tensor(actions, dtype=torch.float32, device='cuda'))
qg, qdg = sim.get_state(); it_g = sim.last_iterations().cpu().numpy()
step_motors(orc, cfg)
q32, qd32 = q.astype(np.float32).astype(np.float64), qd.astype(np.float32).astype(np.float64)
q0, qd0 = q32.copy(), qd32.copy()
orc.batch_step(q32, qd32, actions.astype(np.float32).astype(np.float64), cfg['involved'], 200.0, cfg['obstacle'], cfg['target'], nthreads=8)
err = np.abs(qdg.cpu().numpy() - qd32)
print('PANDA per-joint max |dqd|:', err.max(axis=0))
w = np.argsort(-err.max(axis=1))[:3]
for e in w:
    print('env', e, 'it', it_g[e], 'err', err[e], '\n qd_o', qd32[e], '\n q_in', q0[e], '\n lower', model.lower, '\n upper', model.upper)
