"""
north_star's last sentence as a test: "a NAF agent that learns the obstacle-avoidance task end to end on device".

The reference's kuka_training task (rl_framework.py:547-555: target [0.4, 0.85, 0.71], obstacle [0.45, 0.55, 0.55], start
[0.9, 0.45, 0, 0, 0, 0]) with the reference's hyper-parameters (lr 1e-3, gamma 0.99, tau 1e-3, buffer 100,000, batch 128),
256 arms in the device-resident loop, one NAF update per vectorised step, the tcgen05 trunk bench.py runs.  Training goes on
in chunks of 20,000 updates (about 3 s each) until the GREEDY (noise-free) policy reaches the target in >= 90 % of 512
fresh episodes with no obstacle hit; the test fails when that has not happened after 240,000 updates.  The chunk at which
it happened is printed (tools/train_curve.py and profiles/r1n_train_curve.md hold the full curve).
"""
import logging

import pytest
import torch

pytestmark = pytest.mark.gpu

CHUNK, MAX_UPDATES, N_ENVS, FRAMES = 20000, 240000, 256, 400


def greedy_eval(agent, frames):
    old = agent.noise_scale
    agent.noise_scale = 0.0
    ev = agent.make_loop(frames, 8 * N_ENVS, learn=False, store=False)
    ev.reset_all()
    ev.run_steps(2 * (frames + 60))            # two episodes per env, resets included
    torch.cuda.synchronize()
    k = min(int(ev.log_count.item()), ev.cap)
    last, fr = ev.log_last[:k].cpu().numpy(), ev.log_frame[:k].cpu().numpy()
    agent.noise_scale = old
    return k, float((last == 250).mean()), float((last == -1000).mean()), float(fr.mean())


def test_the_agent_learns_the_obstacle_avoidance_task_on_device():
    from robotic_manipulator_rloa_b200 import ManipulatorFramework
    mf = ManipulatorFramework()
    mf.set_log_level(logging.ERROR)
    try:
        mf.set_hyperparameter('batch_size', 128)
        mf.initialize_environment(manipulator_file='kuka_iiwa/kuka_with_gripper2.sdf', endeffector_index=13,
                                  fixed_joints=[6, 7, 8, 9, 10, 11, 12, 13], involved_joints=[0, 1, 2, 3, 4, 5],
                                  target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
                                  initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0],
                                  initial_positions_variation_range=[0, 0, 0, 0, 0, 0], visualize=False, n_envs=N_ENVS)
        mf.initialize_naf_agent(seed=0)
        agent = mf.naf_agent
        agent.set_trunk_mode(1)
        k0, hit0, obs0, _ = greedy_eval(agent, FRAMES)
        assert k0 >= N_ENVS and hit0 <= 0.05                 # the untrained policy does not solve the task
        loop = agent.make_loop(FRAMES, 1 << 20)
        loop.reset_all()
        done, solved_at, curve = 0, None, []
        while done < MAX_UPDATES:
            loop.log_count.zero_()
            loop.run_steps(CHUNK)
            done += CHUNK
            k, hit, obs, fr = greedy_eval(agent, FRAMES)
            loop.reset_all()                   # the evaluation moved the arms: the training loop starts fresh episodes
            curve.append((done, k, round(100 * hit, 1), round(100 * obs, 1), round(fr, 1)))
            if k >= 2 * N_ENVS - 8 and hit >= 0.90 and obs == 0.0:
                solved_at = done
                break
        print('greedy policy (updates, episodes, target %, obstacle %, mean frames):', curve)
        assert torch.isfinite(agent.last_loss).all()
        assert solved_at is not None, f'greedy success >= 90 % with 0 % obstacle hits not reached in {MAX_UPDATES} updates: {curve}'
    finally:
        mf.set_log_level(logging.INFO)
        mf.delete_environment()
