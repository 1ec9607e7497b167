"""
GPU tests of the callers either side of the kernels: the drop-in
ManipulatorFramework flow of the reference's README, the vectorised loop (eager launches vs the captured CUDA
graph must be the SAME computation), episode bookkeeping with lock-step resets, and checkpoint compatibility.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle.naf_restatement import NAFRef

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
KUKA = dict(manipulator_file='kuka_iiwa/kuka_with_gripper2.sdf', endeffector_index=13,
            fixed_joints=[6, 7, 8, 9, 10, 11, 12, 13], involved_joints=[0, 1, 2, 3, 4, 5],
            target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
            initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0], initial_positions_variation_range=[0, 0, .5, .5, .5, .5],
            visualize=False)


def test_readme_flow_single_env(tmp_path):
    """initialize_environment -> initialize_naf_agent -> run_training -> test_trained_model with one env, exactly as
    a user of the reference calls them (README.md:20-90); checkpoint files load into the reference's layout."""
    from robotic_manipulator_rloa_b200 import ManipulatorFramework
    mf = ManipulatorFramework()
    mf.set_hyperparameter('batch_size', 16)
    mf.initialize_environment(**KUKA)
    mf.initialize_naf_agent(checkpoint_frequency=2, seed=0)
    env = mf.env
    s = env.reset(verbose=False)
    assert isinstance(s, np.ndarray) and s.shape == (21,) and s.dtype == np.float64
    s2, r, d = env.step(np.zeros(6))
    assert s2.shape == (21,) and d in (0, 1) and isinstance(r, (int, float))
    assert env.observation_space.shape == (21,) and env.action_space.shape == (6,)
    scores = mf.run_training(3, 12, verbose=False)
    assert sorted(scores.keys()) == [1, 2, 3] and all(len(v) == 2 for v in scores.values())
    assert all(1 <= v[1] <= 12 for v in scores.values())
    assert os.path.isfile('model.p') and os.path.isfile('checkpoints/2/weights.p') and os.path.isfile('checkpoints/2/scores.txt')
    assert set(json.load(open('checkpoints/2/scores.txt')).keys()) == {'1', '2', '3'}
    sd = torch.load('model.p')
    ref = NAFRef(21, 6, 256, seed=0)
    assert list(sd.keys()) == list(ref.state_dict().keys()) and all(not v.is_cuda for v in sd.values())
    ref.load_state_dict(sd)                                   # the .p file is the reference's 20-entry state_dict
    mf.load_pretrained_parameters_from_episode(2)
    mf.test_trained_model(2, 10)
    assert len(mf.last_test_results) == 2


def _twin(n_envs, seed=0):
    from robotic_manipulator_rloa_b200.environment.environment import Environment, EnvironmentConfiguration
    from robotic_manipulator_rloa_b200.naf_components.naf_algorithm import NAFAgent
    cfg = EnvironmentConfiguration(**{k: v for k, v in KUKA.items() if k != 'manipulator_file'})
    env = Environment(KUKA['manipulator_file'], cfg, n_envs=n_envs, device=DEV, seed=seed)
    agent = NAFAgent(env, 21, 6, 256, 64, 4096, 1e-3, 1e-3, 0.99, 1, 1, 500, DEV, seed=0)
    loop = agent.make_loop(20, 1 << 14)
    loop.reset_all()
    return env, agent, loop


@pytest.mark.parametrize('trunk', [0, 1], ids=['fp32', 'tcgen05'])
def test_graph_replay_equals_eager_launches(trunk):
    """The captured CUDA graph of two loop iterations is the same kernels with the same arguments: after the same
    number of iterations the env states, the replay cursor and every network parameter are bit-identical."""
    n = 256
    runs = []
    for use_graph in (False, True):
        env, agent, loop = _twin(n)
        agent.set_trunk_mode(trunk)
        for _ in range(4):                                    # warm-up: workspaces exist, replay holds > batch
            loop.step()
        if use_graph:
            assert loop.capture(), loop.graph_error
            for _ in range(15):
                loop.replay_pair()
        else:
            for _ in range(30):
                loop.step()
        torch.cuda.synchronize()
        q, qd = env.sim.get_state()
        runs.append(dict(state=loop.state.clone(), q=q.clone(), qd=qd.clone(), cursor=int(agent.memory.cursor.item()),
                         params=torch.cat([p.detach().reshape(-1) for p in agent.qnetwork_main.parameters()]).clone(),
                         episodes=int(loop.log_count.item()), transitions=int(loop.transitions.item()),
                         tick=int(loop.tick.item()), mem_len=len(agent.memory)))
    a, b = runs
    assert a['tick'] == b['tick'] == 34
    assert a['cursor'] == b['cursor'] and a['episodes'] == b['episodes'] and a['transitions'] == b['transitions']
    assert a['mem_len'] == b['mem_len']
    assert torch.equal(a['q'], b['q']) and torch.equal(a['qd'], b['qd']) and torch.equal(a['state'], b['state'])
    assert torch.equal(a['params'], b['params'])


def test_episode_bookkeeping_and_lockstep_reset():
    """frames = 20: every env times out at step 20 (unless it finished earlier), is logged once, spends the next
    50 launches on its reset sub-steps (valid = 0, no transition stored) and then acts again."""
    n = 128
    env, agent, loop = _twin(n)
    loop.learn = False
    for _ in range(20):
        loop.step()
    torch.cuda.synchronize()
    assert int(loop.log_count.item()) >= n                    # everyone finished at least once (done or 20 frames)
    frames = loop.log_frame[:int(loop.log_count.item())].cpu().numpy()
    assert frames.max() == 20 and frames.min() >= 1
    trans_at_20 = int(loop.transitions.item())
    assert trans_at_20 <= 20 * n
    # the envs that timed out at step 20 are all resetting now: the next launch emits no transition for them
    loop.step()
    assert int(loop.valid.sum().item()) < n
    for _ in range(49):
        loop.step()
    loop.step()
    torch.cuda.synchronize()
    assert int(loop.valid.sum().item()) > 0                   # the resets are over, envs act again
    assert int(agent.memory.cursor.item()) == int(loop.transitions.item())   # only valid transitions were stored
    # start poses of the new episodes: reset drives joints 2..5 towards pos + var * U(-1, 1), var = 0.5
    q, _ = env.sim.get_state()
    q = q.cpu().numpy()
    assert np.abs(q[:, 0] - 0.9).max() < 0.2 and np.abs(q[:, 2:6]).max() < 0.75
    assert q[:, 2:6].std() > 0.1                              # the draw differs per env


def test_short_training_run_learns_something():
    """1024 arms x 150 vectorised steps with one update per step: losses stay finite, parameters move, episodes
    complete; the greedy policy's mean end-effector distance to the target is reported before / after."""
    from robotic_manipulator_rloa_b200.environment.environment import Environment, EnvironmentConfiguration
    from robotic_manipulator_rloa_b200.naf_components.naf_algorithm import NAFAgent
    cfg = EnvironmentConfiguration(**{k: v for k, v in KUKA.items() if k != 'manipulator_file'})
    env = Environment(KUKA['manipulator_file'], cfg, n_envs=1024, device=DEV, seed=1)
    agent = NAFAgent(env, 21, 6, 256, 512, 200000, 1e-3, 1e-2, 0.99, 1, 1, 10 ** 9, DEV, seed=0)
    before = torch.cat([p.detach().reshape(-1) for p in agent.qnetwork_main.parameters()]).clone()

    def greedy_distance():
        probe = agent.make_loop(60, 1 << 12, learn=False)
        probe.reset_all()
        old = agent.noise_scale
        agent.noise_scale = 0.0
        for _ in range(40):
            a = agent.act_batch(probe.state)
            env.sim.step(a, out=(probe.next_state, probe.reward, probe.done))
            probe.state, probe.next_state = probe.next_state, probe.state
        agent.noise_scale = old
        ee, tg = probe.state[:, 12:15], probe.state[:, 15:18]
        return float((ee - tg).norm(dim=1).mean().item())

    d0 = greedy_distance()
    loop = agent.make_loop(100, 1 << 16)
    loop.reset_all()
    loop.run_steps(150)
    torch.cuda.synchronize()
    after = torch.cat([p.detach().reshape(-1) for p in agent.qnetwork_main.parameters()])
    assert torch.isfinite(after).all() and torch.isfinite(agent.last_loss).all()
    assert float((after - before).abs().max()) > 1e-3
    assert int(loop.log_count.item()) >= 1024
    d1 = greedy_distance()
    print(f'greedy mean |ee - target| after 40 steps: untrained {d0:.3f} m -> after 150 updates {d1:.3f} m; '
          f'last loss {float(agent.last_loss.item()):.4f}')
    assert np.isfinite(d1)


def test_store_beside_the_update_is_bit_identical():
    """With the one-kernel update the step's rows are copied into the replay ring on another stream WHILE the update runs (it
    reads drawn slots of the pending range from the loop's own buffers) and the cursor moves afterwards.  Against append ->
    sample + learn in sequence: same samples, same parameters, same ring - through wrap-around (ring of 4096 rows, 256 arms),
    partial valid masks (20-frame episodes, 50 reset sub-steps), eager launches and the captured graph."""
    runs = []
    for overlap in (False, True):
        env, agent, loop = _twin(256)
        agent.set_trunk_mode(1)
        loop.overlap_store = overlap
        for _ in range(24):
            loop.step()
        assert loop.capture(), loop.graph_error
        for _ in range(20):
            loop.replay_pair()
        loop.step(); loop.step()
        torch.cuda.synchronize()
        assert (loop._store_stream is not None) == overlap
        m = agent.memory
        runs.append(dict(cursor=int(m.cursor.item()), ring=[t.clone() for t in (m.states, m.actions, m.rewards, m.next_states, m.dones)],
                         params=torch.cat([p.detach().reshape(-1) for net in (agent.qnetwork_main, agent.qnetwork_target)
                                           for p in net.parameters()]).clone(),
                         loss=float(agent.last_loss.item()), transitions=int(loop.transitions.item())))
    a, b = runs
    assert a['cursor'] == b['cursor'] > 4096 and a['transitions'] == b['transitions']
    for x, y in zip(a['ring'], b['ring']):
        assert torch.equal(x, y)
    assert torch.equal(a['params'], b['params']) and a['loss'] == b['loss']


def test_step_host_moves_every_step_through_pinned_buffers():
    """step_host: the caller's pinned buffers carry states in and actions / states / rewards / dones out; the graph
    replays and the eager launches of the same path produce the same trajectory."""
    n = 256
    outs = []
    for use_graph in (False, True):
        env, agent, loop = _twin(n)
        hs, ha = torch.zeros(n, 21).pin_memory(), torch.zeros(n, 6).pin_memory()
        hr, hd = torch.zeros(n).pin_memory(), torch.zeros(n, dtype=torch.uint8).pin_memory()
        hs.copy_(loop.state)
        loop.bind_host_buffers(hs, ha, hr, hd)
        for k in range(12):
            loop.step_host(use_graph=use_graph)
            if k >= 6:      # the call returns once the caller's buffers are complete, possibly before the update ends:
                early = (hs.clone(), hr.clone(), hd.clone())      # what the host sees at return ...
                torch.cuda.synchronize()
                assert torch.equal(early[0], loop.next_state.cpu()) and torch.equal(early[1], loop.reward.cpu())
                assert torch.equal(early[2], loop.done.cpu())     # ... is this step's device result
        torch.cuda.synchronize()
        assert (loop._host_graphs not in (None, False)) == use_graph
        assert torch.equal(hs, loop.next_state.cpu()) and torch.equal(ha, loop.actions.cpu())
        assert torch.equal(hr, loop.reward.cpu()) and torch.equal(hd, loop.done.cpu())
        assert ha.abs().max() <= 1.0 and int(loop.tick.item()) == 12
        outs.append((hs.clone(), ha.clone(), hr.clone(), int(agent.memory.cursor.item())))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][2], outs[1][2]) and outs[0][3] == outs[1][3]
    with pytest.raises(ValueError):
        loop.bind_host_buffers(torch.zeros(n, 21), ha, hr, hd)


def test_config4_rollouts_with_per_env_targets():
    """BASELINE configs[3] at test scale: test_trained_model on the demo KUKA weights with per-env randomised target /
    obstacle positions, several waves of envs (n_episodes > n_envs), results in the reference's (completed, frame) form."""
    from robotic_manipulator_rloa_b200 import ManipulatorFramework
    n = 512
    mf = ManipulatorFramework()
    mf.initialize_environment(**KUKA, n_envs=n)
    mf.initialize_naf_agent()
    mf.load_pretrained_parameters_from_weights_file(os.path.join(
        os.path.dirname(os.path.abspath(__file__)), '..', 'robotic_manipulator_rloa_b200', 'naf_components', 'demo_weights',
        'weights_kuka.p'))
    g = torch.Generator().manual_seed(4321)
    tgt = torch.tensor([0.4, 0.85, 0.71]) + 0.2 * (torch.rand(n, 3, generator=g) - 0.5)
    obs = torch.tensor([0.45, 0.55, 0.55]) + 0.2 * (torch.rand(n, 3, generator=g) - 0.5)
    mf.env.set_task_positions(tgt.to(DEV), obs.to(DEV))
    state = mf.env.get_state()
    assert torch.allclose(state[:, 15:18].cpu(), tgt, atol=1e-6) and torch.allclose(state[:, 18:21].cpu(), obs, atol=1e-6)
    mf.test_trained_model(n + 100, 60)                        # two waves
    res = mf.last_test_results
    assert len(res) == n + 100 and all(isinstance(r[0], bool) and 0 <= r[1] <= 59 for r in res)


def test_training_state_resume_is_bit_identical(tmp_path):
    """save_training_state / load_training_state (an addition over the reference's weights-only checkpoints): a fresh
    agent restored from the file continues appends, samples, updates and noisy actions bit-for-bit."""
    from robotic_manipulator_rloa_b200.naf_components.naf_algorithm import NAFAgent

    def make():
        return NAFAgent(None, 21, 6, 256, 64, 1024, 1e-3, 1e-3, 0.99, 1, 1, 500, DEV, seed=0)

    g = torch.Generator().manual_seed(5)
    chunks = [(torch.randn(300, 21, generator=g).to(DEV), torch.rand(300, 6, generator=g).to(DEV) * 2 - 1,
               torch.randn(300, generator=g).to(DEV), torch.randn(300, 21, generator=g).to(DEV),
               (torch.rand(300, generator=g) < 0.05).to(torch.uint8).to(DEV)) for _ in range(6)]
    probe = torch.randn(128, 21, generator=g).to(DEV)

    def advance(agent, lo, hi):
        for k in range(lo, hi):
            agent.memory.add_batch(*chunks[k])                 # 6 x 300 rows wrap the 1024-row ring
            tick = torch.tensor([k], dtype=torch.int64, device=DEV)
            agent.learn_from_memory(tick=tick)
        tick = torch.tensor([hi], dtype=torch.int64, device=DEV)
        return agent.act_batch(probe, tick=tick).clone()

    straight = make()
    act_a = advance(straight, 0, 6)
    first = make()
    advance(first, 0, 3)
    path = str(tmp_path / 'resume.p')
    first.save_training_state(path)
    resumed = make()
    resumed.load_training_state(path)
    act_b = advance(resumed, 3, 6)
    torch.cuda.synchronize()
    assert torch.equal(act_a, act_b)
    for net in ('qnetwork_main', 'qnetwork_target'):
        sa, sb = getattr(straight, net).state_dict(), getattr(resumed, net).state_dict()
        assert all(torch.equal(sa[k], sb[k]) for k in sa), net
    assert torch.equal(straight.optimizer.exp_avg, resumed.optimizer.exp_avg)
    assert torch.equal(straight.optimizer.exp_avg_sq, resumed.optimizer.exp_avg_sq)
    assert int(straight.optimizer.step_count.item()) == int(resumed.optimizer.step_count.item()) == 6
    for name in ('states', 'next_states', 'actions', 'rewards', 'dones', 'cursor'):
        assert torch.equal(getattr(straight.memory, name), getattr(resumed.memory, name)), name
    assert len(straight.memory) == len(resumed.memory) == 1024
    with pytest.raises(ValueError):
        NAFAgent(None, 23, 7, 256, 64, 1024, 1e-3, 1e-3, 0.99, 1, 1, 500, DEV, seed=0).load_training_state(path)


def test_run_steps_realigns_graph_after_odd_eager_steps():
    """run_steps mixes graph replays (pairs) with eager iterations; after an odd number of eager iterations the state
    buffers and the ping-pong loop counter are out of phase with the captured graph and one eager iteration realigns
    them.  The mixed schedule must equal 24 eager iterations bit-for-bit."""
    n = 256

    def snapshot(env, agent, loop):
        torch.cuda.synchronize()
        q, qd = env.sim.get_state()
        return dict(state=loop.state.clone(), q=q.clone(), qd=qd.clone(), tick=int(loop.tick.item()),
                    cursor=int(agent.memory.cursor.item()), episodes=int(loop.log_count.item()),
                    params=torch.cat([p.detach().reshape(-1) for p in agent.qnetwork_main.parameters()]).clone())

    env, agent, loop = _twin(n)
    for _ in range(24):
        loop.step()
    want = snapshot(env, agent, loop)
    env, agent, loop = _twin(n)
    for _ in range(4):
        loop.step()
    loop.run_steps(6)                    # capture + 3 replays
    assert loop._graph is not None, loop.graph_error
    loop.step()                          # odd
    with pytest.raises(RuntimeError):
        loop.replay_pair()
    loop.run_steps(7)                    # 1 eager (realign) + 3 replays
    loop.run_steps(5)                    # 2 replays + 1 eager
    loop.run_steps(1)                    # eager (realign only)
    got = snapshot(env, agent, loop)
    assert got['tick'] == want['tick'] == 24
    assert got['cursor'] == want['cursor'] and got['episodes'] == want['episodes']
    for k in ('state', 'q', 'qd', 'params'):
        assert torch.equal(got[k], want[k]), k
