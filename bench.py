#!/usr/bin/env python
"""
bench.py — env-steps/s of the vectorised rollout + NAF training hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4|5]

--config numbers follow SURVEY.md section 8(d) (config k = BASELINE.json configs[k-1]):
  1  kuka_training demo through the drop-in facade (1 env, batch 128, eager): latency, not throughput
  2  KUKA IIWA, fixed target / obstacle, 4096 envs per GPU, 400-step episodes, NAF batch 1024      <- default, the headline
  3  as 2 with initial_positions_variation_range [0,0,.5,.5,.5,.5] and 8192 envs per GPU (65,536 at 8 GPUs), batch 1024 per
     GPU (8192 global) with the gradient exchange every update
  4  test_trained_model rollouts on weights_kuka.p, 750 steps x 16,384 envs per GPU, per-env randomised target / obstacle
     (replicas only: no collective)
  5  Franka Panda, 131,072 arms per GPU, sim-only U(-1,1) actions (replicas only)
The default run prints ONE line for config 2 and, inside it, short sub-records of configs 3, 4 and 5 measured by the same
processes (`other_configs`), so that every BASELINE configuration shows up in the driver's 1/2/4/8-GPU records while the
headline stays one workload with fixed per-GPU work (weak scaling).

One training "step" = act (all envs) -> Environment.step -> replay append -> episode bookkeeping / reset scheduling -> replay
sample -> NAFAgent.learn.  `value` = policy-visible transitions (reset sub-steps excluded) per second over all GPUs in STEADY
STATE (the loop is rolled 460 iterations past the start-up transient before the timed region); inputs resident in HBM;
per-pair CUDA events, L2 flushed between timed pairs.  `e2e` = the same loop with every step's states / actions / results
crossing pinned HOST buffers.

--impl reference times the CPU restatement of the same loop (oracle/: fp64 Bullet restatement with OpenMP over all host
cores + torch-CPU NAF), because the reference's own PyBullet path cannot run in this image.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

KUKA = dict(file='kuka_iiwa/kuka_with_gripper2.sdf', ee=13, involved=[0, 1, 2, 3, 4, 5],
            fixed=[6, 7, 8, 9, 10, 11, 12, 13], target=[0.4, 0.85, 0.71], obstacle=[0.45, 0.55, 0.55],
            start=[0.9, 0.45, 0, 0, 0, 0], var=[0, 0, 0, 0, 0, 0])
PANDA = dict(file='franka_panda/panda.urdf', ee=11, involved=[0, 1, 2, 3, 4, 5, 6], fixed=[7, 8, 9, 10, 11],
             target=[0.4, 0.3, 0.5], obstacle=[0.3, 0.0, 0.6], start=[0, 0, 0, -1.5, 0, 1.5, 0])
VAR3 = [0, 0, .5, .5, .5, .5]      # README.md:49 / rl_framework.py:648-649
FRAMES = 400
CONTACT_THRESHOLD = 0.02         # contact rows against obstacle / target on BOTH arms (the product's default, environment/simulator.py)
PREROLL_PAIRS = 230              # 460 untimed iterations > one 400-step episode + its 50 reset sub-steps
FLOP_PER_ENV_STEP = 1.0e5        # KUKA 14 links / 12 dof / 34 rows / 50 iterations
FLOP_PER_SWEEP = 34 * (2 * 12 + 8)   # one Gauss-Seidel sweep over the 34 rows (same table)
BYTES_PER_ENV_STEP = 329         # sim_step kernel only: q,qd r/w 192 + action 24 + task 24 + obs 84 + reward/done 5
NAF_FLOP_PER_SAMPLE = 600064     # fwd main + fwd target (V only) + backward
NAF_BYTES = lambda B: 200 * B + 9 * 318576 + 8 * B * 256 * 4       # multi-launch path: activations cross HBM between kernels
NAF_BYTES_FUSED = lambda B: 200 * B + 9 * 318576 + 2 * 198912       # one-kernel update: sampled rows, parameters / optimiser
                                                                     # state read + written, the two weight images
METRIC = 'env-steps/sec KUKA IIWA (4096 envs/GPU, 400-step episodes, NAF batch 1024, 1 update/step)'

TRAIN_CONFIGS = {
    2: dict(envs=4096, batch=1024, var=KUKA['var'],
            workload='BASELINE.json configs[1]: KUKA IIWA kuka_with_gripper2 (stand-in asset), fixed target/obstacle, '
                     '{envs} envs per GPU, 400-step episodes, NAF batch {batch}, 1 update per vectorised step, live-policy actions'),
    3: dict(envs=8192, batch=1024, var=VAR3,
            workload='BASELINE.json configs[2]: KUKA IIWA with initial_positions_variation_range [0,0,.5,.5,.5,.5], {envs} envs '
                     'per GPU ({total} over {gpus} GPUs), NAF batch {batch} per GPU ({gbatch} global), gradient exchange every update'),
}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return p, 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'sm_max_mhz': 1965.0, 'bf16_tflops_sustained': 1373.4}, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(',')]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------------------
# CPU restatement of the loop (oracle/) — the --impl reference arm and the cpu_baseline leg
# ------------------------------------------------------------------------------------------------------------
class CpuLoop:
    def __init__(self, n_envs: int, batch: int, update_every: int, cores: int, var, seed: int = 0):
        import numpy as np
        import torch
        from helpers import make_oracle, step_motors
        from oracle.naf_restatement import NAFRef
        self.np, self.torch = np, torch
        torch.set_num_threads(cores)
        self.cores, self.n, self.batch, self.update_every = cores, n_envs, batch, update_every
        self.var = np.asarray(var, dtype=np.float64)
        self.model, self.orc = make_oracle(dict(file=KUKA['file'], ee=KUKA['ee'], involved=KUKA['involved']))
        self.rng = np.random.default_rng(seed)
        nl = self.model.nl
        self.q, self.qd = np.zeros((n_envs, nl)), np.zeros((n_envs, nl))
        self.orc.batch_reset(self.q, self.qd, self._starts(n_envs), 50, nthreads=cores, obstacle=KUKA['obstacle'],
                             target=KUKA['target'], contact_threshold=CONTACT_THRESHOLD)
        for j in range(len(KUKA['start'])):
            self.orc.set_position_control(j, KUKA['start'][j])
        step_motors(self.orc, KUKA)
        self.main, self.target = NAFRef(21, 6, 256, seed=0), NAFRef(21, 6, 256, seed=0)
        self.opt = torch.optim.Adam(self.main.parameters(), lr=1e-3)
        self.obs = np.zeros((n_envs, 21), dtype=np.float32)
        for e in range(n_envs):
            self.obs[e] = self.orc.observe(self.q[e], self.qd[e], KUKA['obstacle'], KUKA['target'])[0]
        cap = 100000
        self.rs, self.ra = np.zeros((cap, 21), np.float32), np.zeros((cap, 6), np.float32)
        self.rr, self.rs2, self.rd = np.zeros((cap, 1), np.float32), np.zeros((cap, 21), np.float32), np.zeros((cap, 1), np.float32)
        self.cursor, self.cap, self.t = 0, cap, 0
        self.frame = self.rng.integers(0, FRAMES, n_envs)
        self.transitions = 0

    def _starts(self, k):
        np = self.np
        base = np.tile(np.asarray(KUKA['start'], dtype=np.float64), (k, 1))
        return base + self.var * self.rng.uniform(-1, 1, base.shape)

    def step(self):
        np, torch = self.np, self.torch
        from oracle.naf_restatement import learn_ref
        self.main.eval()
        with torch.no_grad():
            mu, P, _, _ = self.main.heads(torch.from_numpy(self.obs))
            std = torch.rsqrt(torch.diagonal(P, dim1=1, dim2=2))
            act = torch.clamp(mu + std * torch.randn_like(mu), -1, 1).numpy()
        obs2, rew, done, _ = self.orc.batch_step(self.q, self.qd, act.astype(np.float64), KUKA['involved'], 200.0,
                                                 KUKA['obstacle'], KUKA['target'], nthreads=self.cores,
                                                 contact_threshold=CONTACT_THRESHOLD)
        idx = (self.cursor + np.arange(self.n)) % self.cap
        self.rs[idx], self.ra[idx], self.rr[idx, 0] = self.obs, act, rew
        self.rs2[idx], self.rd[idx, 0] = obs2, done
        self.cursor += self.n
        self.transitions += self.n
        self.t += 1
        if self.t % self.update_every == 0 and min(self.cursor, self.cap) > self.batch:
            pick = self.rng.choice(min(self.cursor, self.cap), self.batch, replace=False)
            b = (torch.from_numpy(self.rs[pick]), torch.from_numpy(self.ra[pick]).long(), torch.from_numpy(self.rr[pick]),
                 torch.from_numpy(self.rs2[pick]), torch.from_numpy(self.rd[pick]))
            learn_ref(self.main, self.target, self.opt, b, 0.99, 1e-3)
        self.frame += 1
        fin = (done != 0) | (self.frame >= FRAMES)
        self.obs = obs2.astype(np.float32)
        if fin.any():            # synchronous Environment.reset of the finished envs, as the reference does
            k = np.nonzero(fin)[0]
            q, qd = np.ascontiguousarray(self.q[k]), np.ascontiguousarray(self.qd[k])
            self.orc.batch_reset(q, qd, self._starts(len(k)), 50, nthreads=self.cores, obstacle=KUKA['obstacle'],
                                 target=KUKA['target'], contact_threshold=CONTACT_THRESHOLD)
            self.q[k], self.qd[k] = q, qd
            for i, e in enumerate(k):
                self.obs[e] = self.orc.observe(q[i], qd[i], KUKA['obstacle'], KUKA['target'])[0]
            self.frame[k] = 0


CPU_MIN_WARMUP = 3       # three NAF updates: the first torch-CPU backward / OpenMP team start-up are not steady state


def run_cpu(steps: int, warmup: int, n_envs: int, batch: int, update_every: int, var):
    cores = os.cpu_count() or 1
    loop = CpuLoop(n_envs, batch, update_every, cores, var)
    for _ in range(max(warmup, CPU_MIN_WARMUP)):
        loop.step()
    loop.transitions = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        loop.step()
    dt = time.perf_counter() - t0
    return loop.transitions / dt, dt / steps * 1e3, cores


def train_cfg(args):
    c = dict(TRAIN_CONFIGS[args.config if args.config in TRAIN_CONFIGS else 2])
    if args.envs:
        c['envs'] = args.envs
    if args.batch:
        c['batch'] = args.batch
    return c


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    c = train_cfg(args)
    # one CPU step = the whole per-GPU workload of one GPU step: c['envs'] arms stepped + one batch-B NAF update
    value, ms, cores = run_cpu(args.steps, args.warmup, c['envs'], c['batch'], 1, c['var'])
    sample = (f"the full per-GPU workload per step ({c['envs']} envs stepped, one batch-{c['batch']} NAF update), {args.steps} timed "
              f'steps after {max(args.warmup, CPU_MIN_WARMUP)} warm-up steps')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, c, 'cpu'),
        'cpu_baseline': {'value': value, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'PyBullet is absent from this image: the CPU arm is the fp64 C restatement of the Bullet step '
                '(oracle/, OpenMP over all host cores, one process) + the torch-CPU NAF restatement',
    }
    print(json.dumps(line))


def naf_update_leg(agent, dev, batch):
    """NAF updates/s (the second half of BASELINE.json's metric) on this GPU: NAFAgent.learn through librloa_b200
    against the reference's learn() arithmetic run by torch eager on the SAME GPU (oracle/naf_restatement.py on CUDA —
    the real incumbent for a user of the reference who owns a B200).  CUDA events, L2 not flushed, 50 updates each."""
    import torch
    from oracle.naf_restatement import NAFRef
    g = torch.Generator().manual_seed(0)
    s = torch.randn(batch, 21, generator=g).to(dev)
    s2 = (s + 0.1 * torch.randn(batch, 21, generator=g).to(dev))
    a = torch.clamp(torch.randn(batch, 6, generator=g) * 1.5, -1, 1).to(dev)
    r = -torch.rand(batch, 1, generator=g).to(dev)
    d = torch.zeros(batch, 1, device=dev)
    out = {'batch': batch}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    rr, dd = r.reshape(-1).contiguous(), d.reshape(-1).contiguous()
    out['ours_us_per_update'] = timed(lambda: agent._learn_device(s, a, rr, s2, dd))
    g1 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g1, capture_error_mode='thread_local'):
        agent._learn_device(s, a, rr, s2, dd)
    out['ours_graph_us_per_update'] = timed(g1.replay)
    main, target = NAFRef(21, 6, 256, seed=0).to(dev), NAFRef(21, 6, 256, seed=0).to(dev)
    opt = torch.optim.Adam(main.parameters(), lr=1e-3)
    al = a.long()
    main.train(); target.train()

    def ref_learn():             # naf_algorithm.py:180-226 without any host read (the reference's learn() has none either)
        opt.zero_grad()
        with torch.no_grad():
            v_next = target.heads(s2)[3]
        q = main.heads(s, al)[2]
        torch.nn.functional.mse_loss(q, r + 0.99 * v_next).backward()
        torch.nn.utils.clip_grad_norm_(main.parameters(), 1)
        opt.step()
        with torch.no_grad():
            for pt, pm in zip(target.parameters(), main.parameters()):
                pt.copy_(1e-3 * pm + (1.0 - 1e-3) * pt)

    out['torch_eager_us_per_update'] = timed(ref_learn, n=20)
    out['updates_per_s'] = 1e6 / out['ours_graph_us_per_update']
    out['speedup_vs_torch_eager_same_gpu'] = out['torch_eager_us_per_update'] / out['ours_graph_us_per_update']
    return out


def workload_config(args, c, where):
    gpus = args.gpus
    return {'workload': c['workload'].format(envs=c['envs'], batch=c['batch'], total=c['envs'] * gpus, gpus=gpus,
                                             gbatch=c['batch'] * gpus),
            'envs_per_gpu': c['envs'], 'replay_batch': c['batch'], 'frames': FRAMES, 'parallelism': f'env-dp{gpus}',
            'initial_positions_variation_range': list(c['var']),
            'contacts': 'normal contact rows against the obstacle sphere and the target cube (threshold 0.02 m), as in the reference world',
            'l2': 'flushed between timed step pairs (256 MiB fill, untimed)' if where == 'gpu' else 'n/a',
            'steady_state': f'{2 * getattr(args, "preroll", PREROLL_PAIRS)} untimed iterations before the timed region' if where == 'gpu' else 'n/a',
            'naf_trunk': ('tcgen05: bf16 operands (tf32 for the S-wide input layer), fp32 TMEM accumulate (fused policy kernel + trunk); heads, BatchNorm, '
                          'backward and optimiser fp32' if getattr(args, 'trunk', 'tc') == 'tc' and where == 'gpu' else 'fp32'),
            'launch': 'eager' if getattr(args, 'no_graph', False) or where == 'cpu' else 'cuda-graph of 2 loop iterations'}


# ------------------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide handles of one bench run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from robotic_manipulator_rloa_b200 import _native
        self.torch, self.dist = torch, dist
        self.rank, self.world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)
        assert self.world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={self.world}'
        self.lib = _native.lib()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev) if not args.no_flush else None
        import logging
        import robotic_manipulator_rloa_b200  # noqa: F401  (importing the package installs its logger at INFO, like the reference)
        from robotic_manipulator_rloa_b200.utils.logger import get_global_logger
        logging.getLogger().setLevel(logging.ERROR)
        get_global_logger().setLevel(logging.ERROR)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x: float, op: str) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())


def train_workload(ctx, args, c, steps, full):
    """The training loop of configs 2 / 3 on this process group.  Returns the record of the run; with `full` also the phase
    breakdown, the roofline inputs, e2e through host buffers, the fp32-trunk pass and the N > 1 consistency checks."""
    torch, dist = ctx.torch, ctx.dist
    from robotic_manipulator_rloa_b200.environment.environment import Environment, EnvironmentConfiguration
    from robotic_manipulator_rloa_b200.naf_components.naf_algorithm import NAFAgent
    dev, rank, world, lib = ctx.dev, ctx.rank, ctx.world, ctx.lib
    envs, batch = c['envs'], c['batch']
    cfg = EnvironmentConfiguration(endeffector_index=KUKA['ee'], fixed_joints=KUKA['fixed'],
                                   involved_joints=KUKA['involved'], target_position=KUKA['target'],
                                   obstacle_position=KUKA['obstacle'], initial_joint_positions=KUKA['start'],
                                   initial_positions_variation_range=list(c['var']), visualize=False)
    # seeds: network initialisation shared by all ranks; exploration noise, replay sampling and start poses per rank
    # (NAFAgent / Environment derive them from utils.distributed.rank_seed when world_size > 1)
    env = Environment(KUKA['file'], cfg, n_envs=envs, device=dev, seed=0)
    agent = NAFAgent(env, 21, 6, 256, batch, 100000, 1e-3, 1e-3, 0.99, 1, 1, 500, dev, seed=0)
    if args.trunk == 'tc':
        agent.set_trunk_mode(1)
    loop = agent.make_loop(FRAMES, 1 << 22)
    loop.reset_all()
    loop.frame.copy_(torch.randint(0, FRAMES, (envs,), device=dev, dtype=torch.int32))
    flush = ctx.flush
    use_graph = not args.no_graph
    mk = lambda **kw: [[torch.cuda.Event(enable_timing=True, **kw) for _ in range(6)] for _ in range(2)]
    phase_ev = None
    if use_graph:
        try:
            phase_ev = mk(external=True)
        except TypeError:
            use_graph = False
    if phase_ev is None:
        phase_ev = mk()
    loop.phase_events = None
    for _ in range(max(args.warmup, 3)):
        loop.step()
    graphed = bool(use_graph and loop.capture())
    # ---- roll past the start-up transient: every env has finished an episode and gone through its 50 reset sub-steps
    # at least once, so the share of envs emitting transitions is the stationary one ----
    for _ in range(args.preroll):
        if graphed:
            loop.replay_pair()
        else:
            loop.step(); loop.step()
    ctx.barrier()

    def run_pairs(count, odd, collect, use_g, do_flush=True):
        """`count` pairs (+ `odd` single iteration) of the loop, one CUDA-event pair and one L2 flush per pair."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(count + odd)]
        for i in range(count + odd):
            if flush is not None and do_flush:
                flush.fill_(i & 0xff)
            evs[i][0].record()
            if i < count:
                if use_g:
                    loop.replay_pair()
                else:
                    loop._body(True, True, 0)
                    loop._body(True, True, 1)
            else:
                loop._body(True, True, 0)
            evs[i][1].record()
            if collect is not None:
                evs[i][1].synchronize()
                for par in range(2 if i < count else 1):
                    for k in range(5):
                        collect[0][k].append(phase_ev[par][k].elapsed_time(phase_ev[par][k + 1]))
                    collect[1][par].append(phase_ev[par][0].elapsed_time(phase_ev[par][5]))
        return evs

    def timed(n_steps, do_flush=True):
        npairs, odd = n_steps // 2, n_steps % 2
        loop.transitions.zero_()
        l0 = lib.rloa_launch_count()
        ctx.barrier()
        t_wall = time.perf_counter()
        ev = run_pairs(npairs, odd, None, graphed and loop._graph is not None, do_flush)
        ctx.barrier()
        t_wall = time.perf_counter() - t_wall
        launches = lib.rloa_launch_count() - l0
        if graphed and loop._graph is not None:
            launches = loop.graph_kernels * npairs + launches
        trans_local = float(loop.transitions.item())
        if odd:                                # an odd K leaves the ping-pong buffers off the phase the graph was captured in:
            loop._body(True, True, 1)          # one untimed, uncounted iteration realigns them for the passes that follow
        total_ms = ctx.reduce(sum(a.elapsed_time(b) for a, b in ev), 'MAX')
        trans = ctx.reduce(trans_local, 'SUM')
        return dict(value=trans / (total_ms * 1e-3), ms_per_step=total_ms / n_steps, total_ms=total_ms,
                    sim_substeps_per_s=world * envs * n_steps / (total_ms * 1e-3),
                    valid_fraction=trans / (world * envs * n_steps), launches=int(launches),
                    wall_ms_per_step_incl_flush=t_wall / n_steps * 1e3)

    sampler = ClockSampler(ctx.local)
    if rank == 0 and full:
        sampler.start()
    rec = timed(steps)
    rec['clocks'] = sampler.stop() if (rank == 0 and full) else None
    rec['graphed'], rec['graph_error'] = graphed, loop.graph_error
    if full and flush is not None:
        # the same loop without the flush: what a long training run sees (its working set - weights, optimiser state, kernel
        # code, 4096 arm states - lives in the 126 MB L2); reported beside the headline, never instead of it
        w = timed(steps, do_flush=False)
        rec['warm_l2'] = dict(value=w['value'], ms_per_step=w['ms_per_step'], steps=steps,
                              note='same loop, same per-pair CUDA events, no L2 flush between pairs')
    rec['mean_pgs_sweeps'] = float(env.sim.last_iterations().float().mean().item())

    if full:
        # ---- second pass, same loop and state, WITH the phase-boundary events recorded inside it (re-captured graph):
        # where the step goes, and the live launch time of the simulator step for the roofline ----
        phase_ms = [[] for _ in range(5)]
        parity_ms = [[], []]
        loop.pipeline_sim = False              # the three kernels of the simulator step back to back, so that they can be timed
        loop.phase_events = None
        loop._body(True, True, 0)              # consumes the half-step the pipelined loop left prepared
        loop._body(True, True, 1)
        loop.phase_events = phase_ev
        graphed_phases = False
        if graphed:
            loop._graph = None
            graphed_phases = loop.capture()
        run_pairs(max(8, min(steps // 2, 50)), 0, (phase_ms, parity_ms), graphed_phases)
        ctx.barrier()
        loop.phase_events = None
        loop.pipeline_sim = True
        loop._graph = None
        rec['phases_ms'] = dict(zip(('act', 'env_step', 'replay_append', 'sample_learn', 'bookkeeping_reset'),
                                    [round(sum(x) / len(x), 5) for x in phase_ms]))
        rec['iteration_ms_after_flush_then_warm'] = [round(sum(x) / max(len(x), 1), 5) for x in parity_ms]
        rec['sim_avg_ms'] = sum(phase_ms[1]) / len(phase_ms[1])
        rec['learn_avg_ms'] = sum(phase_ms[3]) / len(phase_ms[3])
        rec['phase_sum_ms'] = sum(sum(x) / len(x) for x in phase_ms)

        # ---- e2e: same loop, every step's states/actions/results cross pinned host buffers -----------------------
        n, S, A = envs, 21, 6
        h_state = torch.zeros(n, S).pin_memory(); h_act = torch.zeros(n, A).pin_memory()
        h_rew = torch.zeros(n).pin_memory(); h_done = torch.zeros(n, dtype=torch.uint8).pin_memory()
        h_state.copy_(loop.state)
        loop.bind_host_buffers(h_state, h_act, h_rew, h_done)
        e2e_steps = max(20, min(5 * steps, 1000))      # ~0.2 s of host-driven steps: a single host hiccup of a few ms must not decide the number
        for _ in range(10):                    # warm-up of the host-facing path (captures its two graphs)
            loop.step_host(use_graph=use_graph)
        loop.transitions.zero_()
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(e2e_steps):
            loop.step_host(use_graph=use_graph)        # the public call: host states in, host actions / results out
        e1.record()
        ctx.barrier()
        e2e_ms = ctx.reduce(e0.elapsed_time(e1), 'MAX')
        e2e_tr = ctx.reduce(float(loop.transitions.item()), 'SUM')
        rec['e2e'] = {'value': e2e_tr / (e2e_ms * 1e-3), 'unit': 'env-steps/s', 'h2d_bytes_per_step': n * (S + A) * 4,
                      'd2h_bytes_per_step': n * (A * 4 + S * 4 + 4 + 1), 'steps': e2e_steps,
                      'ms_per_step': e2e_ms / e2e_steps, 'valid_fraction': e2e_tr / (world * n * e2e_steps),
                      'graphed': bool(loop._host_graphs), 'graph_error': loop.graph_error}

        # ---- the reference-exact arithmetic (all-fp32 NAF, no tensor cores) on the same loop, driver-run ----
        loop.state.copy_(loop.next_state)      # step_host keeps the newest observation in next_state / the host buffer
        if args.trunk == 'tc':
            agent.set_trunk_mode(0)
            loop._graph = None
            for _ in range(4):
                loop.step()
            g2 = bool(use_graph and loop.capture())
            graphed_save, graphed = graphed, g2
            r32 = timed(min(steps, 100))
            graphed = graphed_save
            rec['fp32_trunk'] = {'value': r32['value'], 'ms_per_step': r32['ms_per_step'], 'steps': min(steps, 100),
                                 'note': 'NAFAgent.set_trunk_mode(0): every contraction in fp32 on the CUDA cores — the '
                                         "arithmetic of the reference's torch-CPU path (1e-5 parity tests)"}
            agent.set_trunk_mode(1)
            loop._graph = None

    # ---- N > 1: the ranks' parameters must still be bit-identical and no exchange wait may have timed out ----
    if world > 1:
        with torch.no_grad():
            flat = torch.cat([p.detach().reshape(-1) for net in (agent.qnetwork_main, agent.qnetwork_target)
                              for p in net.parameters()])
            h = flat.view(torch.int32).to(torch.int64)
            k = torch.arange(1, h.numel() + 1, device=dev, dtype=torch.int64)
            sig = torch.stack([h.sum(), (h * (k % 8191)).sum()])
        sigs = [torch.empty_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        rec['ranks_identical'] = all(bool(torch.equal(s, sigs[0])) for s in sigs)
        to = 1.0 if (agent._xchg is not None and agent._xchg.timed_out()) else 0.0
        rec['xchg_timed_out'] = ctx.reduce(to, 'MAX') > 0
        fused = agent._xchg is not None and args.trunk == 'tc' and world <= 8
        rec['grad_exchange'] = ('reduce-scatter + all-gather over NVLink peer memory in tagged 64-bit words, inside the tail of the '
                                'learn kernel (csrc/naf_learn_cluster.cu)' if fused else
                                'NVLink peer memory inside the optimiser kernels (csrc/grad_exchange.cu)'
                                if agent._xchg is not None else 'NCCL all-reduce (torch.distributed)')
    rec['_agent'], rec['_loop'], rec['_env'] = agent, loop, env
    return rec


def release(rec):
    torch_objs = [rec.pop(k, None) for k in ('_loop', '_agent', '_env')]
    loop, agent, env = torch_objs
    if loop is not None:
        loop._graph = None
        loop._host_graphs = None
    if env is not None:
        import torch
        torch.cuda.synchronize()
        env.close()


def quiet_framework():
    """ManipulatorFramework with the package logger at ERROR: the reference logs every test episode at INFO on stdout,
    which must carry the one JSON line only."""
    import logging
    from robotic_manipulator_rloa_b200 import ManipulatorFramework
    from robotic_manipulator_rloa_b200.utils.logger import get_global_logger
    get_global_logger().setLevel(logging.ERROR)
    mf = ManipulatorFramework()
    get_global_logger().setLevel(logging.ERROR)
    logging.getLogger().setLevel(logging.ERROR)
    return mf


def rollout_workload(ctx, n_envs=16384, frames=750):
    """Config 4: test_trained_model on weights_kuka.p, per-env randomised target / obstacle, through the public facade."""
    torch = ctx.torch
    import numpy as np
    from robotic_manipulator_rloa_b200 import ManipulatorFramework
    mf = quiet_framework()
    mf.initialize_environment(manipulator_file=KUKA['file'], endeffector_index=13, fixed_joints=KUKA['fixed'],
                              involved_joints=KUKA['involved'], target_position=KUKA['target'],
                              obstacle_position=KUKA['obstacle'], initial_joint_positions=KUKA['start'],
                              initial_positions_variation_range=VAR3, visualize=False, n_envs=n_envs, device=ctx.dev)
    mf.initialize_naf_agent()
    mf.load_pretrained_parameters_from_weights_file(
        os.path.join(ROOT, 'robotic_manipulator_rloa_b200', 'naf_components', 'demo_weights', 'weights_kuka.p'))
    mf.naf_agent.set_trunk_mode(1)
    g = torch.Generator().manual_seed(4321 + ctx.rank)
    tgt = torch.tensor(KUKA['target']) + 0.2 * (torch.rand(n_envs, 3, generator=g) - 0.5)
    obs = torch.tensor(KUKA['obstacle']) + 0.2 * (torch.rand(n_envs, 3, generator=g) - 0.5)
    mf.env.set_task_positions(tgt.to(ctx.dev), obs.to(ctx.dev))
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mf.test_trained_model(n_envs, frames)
    e1.record()
    ctx.barrier()
    ms = ctx.reduce(e0.elapsed_time(e1), 'MAX')
    res = mf.last_test_results
    ok = np.array([r[0] for r in res]); fr = np.array([r[1] for r in res])
    coll = sum(1 for r in res if (not r[0]) and r[1] < frames - 1)
    steps = ctx.reduce(float((fr + 1).sum()), 'SUM')
    out = {'workload': f'BASELINE.json configs[3]: test_trained_model on weights_kuka.p, {n_envs} envs per GPU x <= {frames} steps, '
                       'per-env target / obstacle = nominal + U(-0.1, 0.1)^3, start pose var [0,0,.5,.5,.5,.5] (stand-in KUKA asset); '
                       'replicas only',
           'value': steps / (ms * 1e-3), 'unit': 'env-steps/s (policy steps of episodes still running, incl. the 50-sub-step reset in the time)',
           'ms_total': ms, 'episodes': int(ctx.reduce(float(len(res)), 'SUM')),
           'success_pct': 100.0 * ctx.reduce(float(ok.sum()), 'SUM') / ctx.reduce(float(len(res)), 'SUM'),
           'collision_pct': 100.0 * ctx.reduce(float(coll), 'SUM') / ctx.reduce(float(len(res)), 'SUM'),
           'mean_frames_of_successes': float(fr[ok].mean()) if ok.any() else None}
    mf.delete_environment()
    return out


def panda_workload(ctx, n=131072, steps=30):
    """Config 5: Panda-like 12-joint / 9-dof model, sim-only U(-1,1) actions, n arms per GPU (replicas only)."""
    torch = ctx.torch
    from robotic_manipulator_rloa_b200.environment.robot_model import load_manipulator
    from robotic_manipulator_rloa_b200.environment.simulator import BatchedSimulator
    model = load_manipulator(PANDA['file'])
    sim = BatchedSimulator(model, n, PANDA['ee'], PANDA['involved'], PANDA['fixed'], device=ctx.dev)
    sim.set_task(PANDA['target'], PANDA['obstacle'])
    g = torch.Generator(device=ctx.dev).manual_seed(1234 + ctx.rank)
    base = torch.tensor(PANDA['start'], device=ctx.dev, dtype=torch.float32)
    sim.reset(base + 0.5 * (torch.rand(n, 7, device=ctx.dev, generator=g) - 0.5))
    acts = [2 * torch.rand(n, 7, device=ctx.dev, generator=g) - 1 for _ in range(8)]
    for i in range(5):
        sim.step(acts[i % 8])
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        sim.step(acts[i % 8])
    e1.record()
    ctx.barrier()
    ms = ctx.reduce(e0.elapsed_time(e1), 'MAX')
    out = {'workload': f'BASELINE.json configs[4]: Franka Panda-like (stand-in asset, 12 joints / 9 dof), {n} arms per GPU '
                       f'({n * ctx.world} total), sim-only U(-1,1) actions; replicas only; state footprint > L2, no flush needed',
           'value': ctx.world * n * steps / (ms * 1e-3), 'unit': 'env-steps/s', 'us_per_step': ms / steps * 1e3,
           'mean_pgs_sweeps': float(sim.last_iterations().float().mean().item()), 'steps': steps}
    sim.close()
    return out


def demo_workload(ctx):
    """Config 1: the reference's kuka_training demo shape (1 env, batch 128, 10 episodes x 400 frames) through the facade."""
    torch = ctx.torch
    from robotic_manipulator_rloa_b200 import ManipulatorFramework
    mf = quiet_framework()
    mf.initialize_environment(manipulator_file=KUKA['file'], endeffector_index=13, fixed_joints=KUKA['fixed'],
                              involved_joints=KUKA['involved'], target_position=KUKA['target'],
                              obstacle_position=KUKA['obstacle'], initial_joint_positions=KUKA['start'],
                              initial_positions_variation_range=KUKA['var'], visualize=False, device=ctx.dev)
    mf.initialize_naf_agent()
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    try:
        mf.run_training(1, 400, verbose=False)           # warm-up episode
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        scores = mf.run_training(10, 400, verbose=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    finally:
        os.chdir(cwd)
    frames = sum(f + 1 for _, f in scores.values())
    mf.delete_environment()
    return {'workload': 'BASELINE.json configs[0]: kuka_training demo, 1 env, 10 episodes x <= 400 frames, NAF batch 128, through '
                        'ManipulatorFramework.run_training (eager launches, one host read per step as the reference loop has)',
            'value': frames / dt, 'unit': 'env-steps/s', 'ms_per_step': dt / frames * 1e3, 'frames': frames}


def ours(args):
    ctx = Ctx(args)
    torch, rank, world = ctx.torch, ctx.rank, ctx.world
    peaks, which = measured_peaks()
    if args.config == 1:
        out = demo_workload(ctx)
        line = {'metric': METRIC, 'value': out['value'], 'unit': 'env-steps/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': out['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': out['workload']},
                'gpu_launches': int(ctx.lib.rloa_launch_count())}
        if rank == 0:
            print(json.dumps(line), flush=True)
        return finish(ctx, 0)
    if args.config in (4, 5):
        out = rollout_workload(ctx) if args.config == 4 else panda_workload(ctx, steps=max(args.steps, 10))
        line = {'metric': METRIC, 'value': out['value'], 'unit': 'env-steps/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': out.get('us_per_step', out.get('ms_total', 0.0) * 1e3) / 1e3,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': out['workload'], 'parallelism': f'replicas x{world}'}, 'detail': out,
                'gpu_launches': int(ctx.lib.rloa_launch_count())}
        if rank == 0:
            print(json.dumps(line), flush=True)
        return finish(ctx, 0)

    c = train_cfg(args)
    rec = train_workload(ctx, args, c, args.steps, full=True)
    agent = rec['_agent']
    envs, batch = c['envs'], c['batch']
    naf = cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            naf = naf_update_leg(agent, ctx.dev, batch)
        cpu_value, cpu_ms, cores = run_cpu(40, 3, envs, batch, 1, c['var'])
        cpu = {'value': cpu_value, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port',
               'sample': '40 steps of the full per-GPU workload (%d envs stepped + one batch-%d NAF update) after 3 warm-up '
                         'steps; fp64 C restatement of the Bullet step (OpenMP, all cores, one process) + torch-CPU NAF'
                         % (envs, batch)}
    release(rec)

    # ---- the other BASELINE configurations, measured by the same processes (short) ----
    others = {}
    if not args.no_extras:
        if args.config == 2:
            try:
                c3 = dict(TRAIN_CONFIGS[3])
                r3 = train_workload(ctx, args, c3, min(args.steps, 100), full=False)
                others['config3'] = {'workload': workload_config(args, c3, 'gpu')['workload'], 'value': r3['value'],
                                     'unit': 'env-steps/s', 'ms_per_step': r3['ms_per_step'], 'valid_fraction': r3['valid_fraction'],
                                     'steps': min(args.steps, 100), 'ranks_identical': r3.get('ranks_identical'),
                                     'xchg_timed_out': r3.get('xchg_timed_out'), 'graphed': r3['graphed']}
                if world > 1 and (not r3.get('ranks_identical', True) or r3.get('xchg_timed_out')):
                    rec['ranks_identical'] = False
                release(r3)
            except Exception as err:           # a sub-record must not lose the headline
                others['config3'] = {'error': f'{type(err).__name__}: {err}'}
        for name, fn in (('config4', rollout_workload), ('config5', panda_workload)):
            try:
                others[name] = fn(ctx)
            except Exception as err:
                others[name] = {'error': f'{type(err).__name__}: {err}'}

    rc = 0
    if world > 1 and (not rec.get('ranks_identical', True) or rec.get('xchg_timed_out')):
        rc = 3
    if rank == 0:
        hbm_peak, sm_max = float(peaks['hbm_gbs']), float(peaks.get('sm_max_mhz', 1965.0))
        tensor_peak = float(peaks.get('bf16_tflops_sustained', 1373.4))
        clocks = rec['clocks']
        sim_avg_ms = rec['sim_avg_ms']
        sm_mhz = (clocks or {}).get('sm_mhz') or sm_max
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        tflops_nominal = FLOP_PER_ENV_STEP * envs / (sim_avg_ms * 1e-3) / 1e12
        sweeps = rec['mean_pgs_sweeps']
        flop_measured = FLOP_PER_ENV_STEP - (50.0 - sweeps) * FLOP_PER_SWEEP
        tflops_measured = flop_measured * envs / (sim_avg_ms * 1e-3) / 1e12
        gbs = BYTES_PER_ENV_STEP * envs / (sim_avg_ms * 1e-3) / 1e9
        learn_us = (naf or {}).get('ours_graph_us_per_update') or rec['learn_avg_ms'] * 1e3
        naf_tf = NAF_FLOP_PER_SAMPLE * batch / (learn_us * 1e-6) / 1e12
        naf_bytes = NAF_BYTES_FUSED(batch) if args.trunk == 'tc' else NAF_BYTES(batch)
        naf_gbs = naf_bytes / (learn_us * 1e-6) / 1e9
        line = {
            'metric': METRIC, 'value': rec['value'], 'unit': 'env-steps/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': rec['ms_per_step'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, c, 'gpu'),
            'naf_updates_per_s': args.steps / (rec['total_ms'] * 1e-3),
            'sim_substeps_per_s': rec['sim_substeps_per_s'],
            'valid_fraction': rec['valid_fraction'],
            'wall_ms_per_step_incl_flush': rec['wall_ms_per_step_incl_flush'],
            'roofline': {'kernel': 'rloa_sim_step = sim_dynamics_kernel + sim_minv_kernel + sim_contacts_kernel + sim_solve_kernel', 'bound': 'fp32',
                         'achieved': tflops_nominal, 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': tflops_nominal / fp32_peak,
                         'flop_per_env_step': FLOP_PER_ENV_STEP,
                         'peak_source': f'148 SMs x 128 FMA lanes x 2 flop x {sm_mhz:.0f} MHz (median SM clock sampled during the timed region)',
                         'at_measured_sweeps': {'mean_pgs_sweeps': sweeps, 'flop_per_env_step': flop_measured,
                                                'achieved': tflops_measured, 'frac': tflops_measured / fp32_peak},
                         'avg_launch_ms': sim_avg_ms, 'share_of_step': sim_avg_ms / max(rec['phase_sum_ms'], 1e-9),
                         'traffic': 8.47e6, 'traffic_source': 'ncu --set full at 4096 arms, dram read+write of the four kernels per '
                                                              'launch, caches flushed by ncu (profiles/r3i_sim4096_ncu_full.md)',
                         'hbm': {'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak, 'peak_source': which,
                                 'bytes_per_env_step': BYTES_PER_ENV_STEP}},
            'roofline_naf_update': {'kernel': 'NAFAgent.learn (replay batch %d, both networks, backward, clip + Adam + soft update)' % batch,
                                    'bound': 'tensor', 'achieved': naf_tf, 'peak': tensor_peak, 'unit': 'TFLOP/s',
                                    'frac': naf_tf / tensor_peak, 'flop_per_update': NAF_FLOP_PER_SAMPLE * batch,
                                    'us_per_update': learn_us, 'hbm_gbs': naf_gbs, 'hbm_frac': naf_gbs / hbm_peak,
                                    'bytes_per_update': naf_bytes, 'peak_source': which + ' bf16_tflops_sustained',
                                    'traffic': 2.77e6 if args.trunk == 'tc' else None,
                                    'traffic_source': 'ncu --set full, dram read+write of learn_pack_kernel + naf_learn_cluster_kernel per update '
                                                      '(profiles/r3c_learn_cluster_ncu_full.md; the writes stay in L2)',
                                    'note': 'latency-bound at this batch: both floors (0.45 us tensor, 0.5 - 1.8 us HBM) are far below the '
                                            'barrier / epilogue latency of the chain'},
            'cpu_baseline': cpu,
            'e2e': rec['e2e'],
            'fp32_trunk': rec.get('fp32_trunk'),
            'phases_note': 'second pass over the same loop with event nodes inside it (each costs ~4 us) and WITHOUT the software pipelining of the '
                           'simulator step, so its three kernels run back to back and can be timed: phases sum to more than ms_per_step',
            'phases_ms': rec['phases_ms'],
            'iteration_ms_after_flush_then_warm': rec['iteration_ms_after_flush_then_warm'],
            'warm_l2': rec.get('warm_l2'),
            'naf_update': naf,
            'gpu_launches': rec['launches'],
            'graphed': rec['graphed'], 'graph_error': rec['graph_error'],
            'clocks': clocks,
            'other_configs': others,
        }
        if world > 1:
            line['ranks_identical'] = rec.get('ranks_identical')
            line['xchg_timed_out'] = rec.get('xchg_timed_out')
            line['grad_exchange'] = rec.get('grad_exchange')
        print(json.dumps(line), flush=True)
        if rc:
            sys.stderr.write('bench.py: ranks diverged or the gradient exchange timed out\n')
    finish(ctx, rc)


def finish(ctx, rc):
    if ctx.world > 1:
        # a communicator that refuses to finalise must not turn a finished measurement into a hung job
        ctx.torch.cuda.synchronize()
        ctx.dist.barrier()
        import threading
        threading.Timer(20.0, lambda: os._exit(rc)).start()
        ctx.dist.destroy_process_group()
        sys.stdout.flush()
        os._exit(rc)
    if rc:
        sys.exit(rc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help='SURVEY.md section 8(d) config number = BASELINE.json configs[k-1]; 2 is the headline')
    ap.add_argument('--envs', type=int, default=0, help='override envs per GPU of the training configs')
    ap.add_argument('--batch', type=int, default=0, help='override the replay batch per GPU')
    ap.add_argument('--trunk', default='tc', choices=['fp32', 'tc'],
                    help='tc: tcgen05 policy kernel + trunk (bf16 operands, fp32 accumulate); fp32: CUDA-core reference-exact path')
    ap.add_argument('--no-flush', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel eagerly instead of replaying the CUDA graph')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-extras', action='store_true', help='skip the other_configs sub-records')
    ap.add_argument('--preroll', type=int, default=PREROLL_PAIRS,
                    help='untimed iteration pairs before the timed region (steady state needs >= 225; profiling runs pass a small number)')
    args = ap.parse_args()
    if args.impl == 'reference':
        reference_arm(args)
    else:
        ours(args)


if __name__ == '__main__':
    main()
