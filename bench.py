#!/usr/bin/env python
"""
bench.py — env-steps/s of the vectorised rollout + NAF training hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E] [--batch B]

Workload (config.workload): BASELINE.json configs[1] — KUKA IIWA, fixed target / obstacle, E = 4096 envs per
GPU, 400-step episodes with lock-step asynchronous resets, NAF replay batch 1024, one NAF update per
vectorised step, actions from the live policy (seed-0 init, reference exploration noise).
One "step" = act (4096 states) -> Environment.step (4096 arms) -> replay append -> episode bookkeeping / reset
scheduling -> replay sample -> NAFAgent.learn.  `value` = policy-visible transitions (reset sub-steps excluded)
per second over all GPUs; inputs are resident in HBM; per-step CUDA events, L2 flushed between timed steps.
`e2e` = the same loop with every step's states / actions / results crossing pinned HOST buffers.

--impl reference times the CPU restatement of the same loop (oracle/: fp64 Bullet restatement with OpenMP over
all host cores + torch-CPU NAF), because the reference's own PyBullet path cannot run in this image.
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
FRAMES = 400
FLOP_PER_ENV_STEP = 1.0e5        # KUKA 14 links / 12 dof / 34 rows / 50 iterations
BYTES_PER_ENV_STEP = 329         # sim_step kernel only: q,qd r/w 192 + action 24 + task 24 + obs 84 + reward/done 5
METRIC = 'env-steps/sec KUKA IIWA (4096 envs/GPU, 400-step episodes, NAF batch 1024, 1 update/step)'


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), float(p.get('sm_max_mhz', 1965.0)), 'measured'
    except Exception:
        return 6650.0, 1965.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
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
    def __init__(self, n_envs: int, batch: int, update_every: int, cores: int, seed: int = 0):
        import numpy as np
        import torch
        from helpers import make_oracle, step_motors
        from oracle.naf_restatement import NAFRef
        self.np, self.torch = np, torch
        torch.set_num_threads(cores)
        self.cores, self.n, self.batch, self.update_every = cores, n_envs, batch, update_every
        self.cfg = dict(KUKA)
        self.model, self.orc = make_oracle(dict(file=KUKA['file'], ee=KUKA['ee'], involved=KUKA['involved']))
        self.rng = np.random.default_rng(seed)
        nl = self.model.nl
        self.q, self.qd = np.zeros((n_envs, nl)), np.zeros((n_envs, nl))
        init = np.tile(np.asarray(KUKA['start'], dtype=np.float64), (n_envs, 1))
        self.orc.batch_reset(self.q, self.qd, init, 50, nthreads=cores)
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

    def step(self):
        np, torch = self.np, self.torch
        from oracle.naf_restatement import learn_ref
        self.main.eval()
        with torch.no_grad():
            mu, P, _, _ = self.main.heads(torch.from_numpy(self.obs))
            std = torch.rsqrt(torch.diagonal(P, dim1=1, dim2=2))
            act = torch.clamp(mu + std * torch.randn_like(mu), -1, 1).numpy()
        obs2, rew, done, _ = self.orc.batch_step(self.q, self.qd, act.astype(np.float64), KUKA['involved'], 200.0,
                                                 KUKA['obstacle'], KUKA['target'], nthreads=self.cores)
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
            init = np.tile(np.asarray(KUKA['start'], dtype=np.float64), (len(k), 1))
            self.orc.batch_reset(q, qd, init, 50, nthreads=self.cores)
            self.q[k], self.qd[k] = q, qd
            for i, e in enumerate(k):
                self.obs[e] = self.orc.observe(q[i], qd[i], KUKA['obstacle'], KUKA['target'])[0]
            self.frame[k] = 0


CPU_MIN_WARMUP = 3       # three NAF updates: the first torch-CPU backward / OpenMP team start-up are not steady state


def run_cpu(steps: int, warmup: int, n_envs: int, batch: int, update_every: int):
    cores = os.cpu_count() or 1
    loop = CpuLoop(n_envs, batch, update_every, cores)
    for _ in range(max(warmup, CPU_MIN_WARMUP)):
        loop.step()
    loop.transitions = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        loop.step()
    dt = time.perf_counter() - t0
    return loop.transitions / dt, dt / steps * 1e3, cores


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # one CPU step = the whole per-GPU workload of one GPU step: args.envs arms stepped + one batch-B NAF update
    value, ms, cores = run_cpu(args.steps, args.warmup, args.envs, args.batch, 1)
    sample = (f'the full per-GPU workload per step ({args.envs} envs stepped, one batch-{args.batch} NAF update), {args.steps} timed '
              f'steps after {max(args.warmup, CPU_MIN_WARMUP)} warm-up steps')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, 'cpu'),
        'cpu_baseline': {'value': value, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'PyBullet is absent from this image: the CPU arm is the fp64 C restatement of the Bullet step '
                '(oracle/, OpenMP over all host cores) + the torch-CPU NAF restatement',
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


def workload_config(args, where):
    return {'workload': 'BASELINE.json configs[1]: KUKA IIWA kuka_with_gripper2 (stand-in asset), fixed target/obstacle, '
                        f'{args.envs} envs per GPU, {FRAMES}-step episodes, NAF batch {args.batch}, 1 update per '
                        'vectorised step, live-policy actions',
            'envs_per_gpu': args.envs, 'replay_batch': args.batch, 'frames': FRAMES, 'parallelism': f'env-dp{args.gpus}',
            'l2': 'flushed between timed step pairs (256 MiB fill, untimed)' if where == 'gpu' else 'n/a',
            'naf_trunk': ('tcgen05: bf16 operands (tf32 for the S-wide input layer), fp32 TMEM accumulate (fused policy kernel + trunk); heads, BatchNorm, '
                          'backward and optimiser fp32' if getattr(args, 'trunk', 'tc') == 'tc' and where == 'gpu' else 'fp32'),
            'launch': 'eager' if getattr(args, 'no_graph', False) or where == 'cpu' else 'cuda-graph of 2 loop iterations'}


# ------------------------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist
    from robotic_manipulator_rloa_b200 import _native
    from robotic_manipulator_rloa_b200.environment.environment import Environment, EnvironmentConfiguration
    from robotic_manipulator_rloa_b200.naf_components.naf_algorithm import NAFAgent

    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    lib = _native.lib()

    import logging
    logging.getLogger().setLevel(logging.ERROR)
    cfg = EnvironmentConfiguration(endeffector_index=KUKA['ee'], fixed_joints=KUKA['fixed'],
                                   involved_joints=KUKA['involved'], target_position=KUKA['target'],
                                   obstacle_position=KUKA['obstacle'], initial_joint_positions=KUKA['start'],
                                   initial_positions_variation_range=KUKA['var'], visualize=False)
    env = Environment(KUKA['file'], cfg, n_envs=args.envs, device=dev, seed=rank)
    agent = NAFAgent(env, 21, 6, 256, args.batch, 100000, 1e-3, 1e-3, 0.99, 1, 1, 500, dev, seed=0)
    agent.seed = 1000 + rank                      # exploration noise differs per rank; weights start identical
    agent.memory.seed = 2000 + rank
    if args.trunk == 'tc':
        agent.set_trunk_mode(1)
    loop = agent.make_loop(FRAMES, 1 << 22)
    loop.reset_all()
    # steady-state episode phases: spread the 400-step timeouts uniformly
    loop.frame.copy_(torch.randint(0, FRAMES, (args.envs,), device=dev, dtype=torch.int32))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if not args.no_flush else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # phase-boundary events recorded inside the loop body (act | Environment.step | append | learn | tail);
    # `external` lets them sit inside a CUDA graph as event-record nodes
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
    if graphed:
        loop.replay_pair()                        # one untimed replay
    barrier()

    def run_pairs(count, odd, collect, use_g):
        """`count` pairs (+ `odd` single iteration) of the loop, one CUDA-event pair and one L2 flush per pair."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(count + odd)]
        for i in range(count + odd):
            if flush is not None:
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
                # the phase events are re-recorded by every pair: read them before the next pair overwrites them
                evs[i][1].synchronize()
                for par in range(2 if i < count else 1):
                    for k in range(5):
                        collect[0][k].append(phase_ev[par][k].elapsed_time(phase_ev[par][k + 1]))
                    collect[1][par].append(phase_ev[par][0].elapsed_time(phase_ev[par][5]))
        return evs

    # ---- timed region: K steps in pairs (the state buffers ping-pong), CUDA events per pair, L2 flushed between pairs.
    # No event nodes inside the loop here: six of them per iteration cost ~25 us, 11 % of the step ----
    npairs, odd = args.steps // 2, args.steps % 2
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    loop.transitions.zero_()
    launches0 = lib.rloa_launch_count()
    barrier()
    t_wall = time.perf_counter()
    ev = run_pairs(npairs, odd, None, graphed)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = lib.rloa_launch_count() - launches0
    if graphed:      # a graph replay re-launches the kernels recorded at capture: count them per replay
        launches = loop.graph_kernels * npairs + launches
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    trans = loop.transitions.clone().to(torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(trans, op=dist.ReduceOp.SUM)
    total_ms, trans = float(total_ms.item()), float(trans.item())
    value = trans / (total_ms * 1e-3)

    # ---- second pass, same loop and state, WITH the phase-boundary events recorded inside it (re-captured graph):
    # where the step goes, and the live launch time of the simulator step for the roofline ----
    phase_ms = [[] for _ in range(5)]
    parity_ms = [[], []]                   # whole iteration right after the L2 flush / the one after it
    loop.phase_events = phase_ev
    loop.pipeline_sim = False              # the three kernels of the simulator step back to back, so that they can be timed
    loop.phase_events = None
    loop._body(True, True, 0)              # consumes the half-step the pipelined loop left prepared
    loop._body(True, True, 1)
    loop.phase_events = phase_ev
    graphed_phases = False
    if graphed:
        loop._graph = None
        graphed_phases = loop.capture()
    run_pairs(max(8, min(npairs, 50)), 0, (phase_ms, parity_ms), graphed_phases)
    barrier()
    loop.phase_events = None
    loop.pipeline_sim = True
    loop._graph = None
    sim_ms = phase_ms[1]

    # ---- e2e: same loop, every step's states/actions/results cross pinned host buffers -----------------------
    n, S, A = args.envs, 21, 6
    h_state = torch.zeros(n, S).pin_memory(); h_act = torch.zeros(n, A).pin_memory()
    h_rew = torch.zeros(n).pin_memory(); h_done = torch.zeros(n, dtype=torch.uint8).pin_memory()
    h_state.copy_(loop.state)
    loop.bind_host_buffers(h_state, h_act, h_rew, h_done)
    e2e_steps = max(10, args.steps // 4)
    for _ in range(3):                     # warm-up of the host-facing path (captures its two graphs)
        loop.step_host(use_graph=use_graph)
    loop.transitions.zero_()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(e2e_steps):
        loop.step_host(use_graph=use_graph)        # the public call: host states in, host actions / results out
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    e2e_tr = loop.transitions.clone().to(torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_tr, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_tr.item()) / (float(e2e_ms.item()) * 1e-3)

    if rank == 0:
        hbm_peak, sm_max, which = measured_peaks()
        sim_avg_ms = sum(sim_ms) / len(sim_ms)
        # env sub-steps executed per launch (action steps + reset sub-steps): every env runs one
        gbs = BYTES_PER_ENV_STEP * args.envs / (sim_avg_ms * 1e-3) / 1e9
        sm_mhz = (clocks or {}).get('sm_mhz') or sm_max
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        tflops = FLOP_PER_ENV_STEP * args.envs / (sim_avg_ms * 1e-3) / 1e12
        cpu_value, cpu_ms, cores = (None, None, os.cpu_count())
        cpu = None
        naf = None
        if world == 1 and not args.no_cpu:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                naf = naf_update_leg(agent, dev, args.batch)
            cpu_value, cpu_ms, cores = run_cpu(40, 3, args.envs, args.batch, 1)
            cpu = {'value': cpu_value, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port',
                   'sample': '40 steps of the full per-GPU workload (%d envs stepped + one batch-%d NAF update) after 3 warm-up '
                             'steps; fp64 C restatement of the Bullet step (OpenMP, all cores) + torch-CPU NAF'
                             % (args.envs, args.batch)}
        line = {
            'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, 'gpu'),
            'naf_updates_per_s': args.steps / (total_ms * 1e-3),
            'sim_substeps_per_s': world * args.envs * args.steps / (total_ms * 1e-3),
            'wall_ms_per_step_incl_flush': t_wall / args.steps * 1e3,
            'roofline': {'kernel': 'rloa_sim_step = sim_dynamics_kernel + sim_minv_kernel + sim_solve_kernel', 'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                         'frac': gbs / hbm_peak, 'traffic': 8.07e6, 'traffic_source': 'ncu --set full at 4096 arms, dram read+write of '
                         'the three kernels per launch, caches flushed by ncu (profiles/r1m_sim4096_ncu_full.md)', 'peak_source': which,
                         'avg_launch_ms': sim_avg_ms,
                         'share_of_step': sim_avg_ms / max(sum(sum(x) / len(x) for x in phase_ms), 1e-9),
                         'note': 'the kernel is FP32-issue / latency bound, not HBM bound (SURVEY 8d): see fp32',
                         'fp32': {'achieved': tflops, 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': tflops / fp32_peak,
                                  'flop_per_env_step': FLOP_PER_ENV_STEP, 'sm_mhz': sm_mhz}},
            'cpu_baseline': cpu,
            'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': n * (S + A) * 4,
                    'd2h_bytes_per_step': n * (A * 4 + S * 4 + 4 + 1), 'steps': e2e_steps},
            'phases_note': 'second pass over the same loop with event nodes inside it (each costs ~4 us) and WITHOUT the software pipelining of the '
                           'simulator step, so its three kernels run back to back and can be timed: phases sum to more than ms_per_step',
            'phases_ms': dict(zip(('act', 'env_step', 'replay_append', 'sample_learn', 'bookkeeping_reset'),
                                  [round(sum(x) / len(x), 5) for x in phase_ms])),
            'iteration_ms_after_flush_then_warm': [round(sum(x) / max(len(x), 1), 5) for x in parity_ms],
            'naf_update': naf,
            'gpu_launches': int(launches),
            'graphed': graphed, 'graph_error': loop.graph_error,
            'clocks': clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # tear down in dependency order: the captured graph holds NCCL work, so it goes first; a communicator that
        # still refuses to finalise must not turn a finished measurement into a hung job
        loop._graph = None
        torch.cuda.synchronize()
        dist.barrier()
        import threading
        threading.Timer(20.0, lambda: os._exit(0)).start()
        dist.destroy_process_group()
        sys.stdout.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--envs', type=int, default=4096, help='envs per GPU')
    ap.add_argument('--batch', type=int, default=1024, help='replay batch per GPU')
    ap.add_argument('--trunk', default='tc', choices=['fp32', 'tc'],
                    help='tc: tcgen05 policy kernel + hidden layer (bf16 operands, fp32 accumulate); fp32: CUDA-core reference-exact path')
    ap.add_argument('--no-flush', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel eagerly instead of replaying the CUDA graph')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        reference_arm(args)
    else:
        ours(args)


if __name__ == '__main__':
    main()
