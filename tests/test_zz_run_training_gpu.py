"""
GPU test of ManipulatorFramework.run_training with n_envs > 1: the vectorised loop behind the reference's public API
(naf_algorithm.py:228-292 for every env at once) — eager warm-up, then CUDA-graph replays, episode accounting in
completion order, checkpoint cadence and file layout.  (Named to run last: it exercises everything at once.)
"""
import json
import logging
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_run_training_vectorised_through_the_public_api():
    from robotic_manipulator_rloa_b200 import ManipulatorFramework
    mf = ManipulatorFramework()
    mf.set_log_level(logging.ERROR)               # four log lines per finished episode otherwise
    try:
        mf.set_hyperparameter('batch_size', 64)
        mf.initialize_environment(manipulator_file='kuka_iiwa/kuka_with_gripper2.sdf', endeffector_index=13,
                                  fixed_joints=[6, 7, 8, 9, 10, 11, 12, 13], involved_joints=[0, 1, 2, 3, 4, 5],
                                  target_position=[0.4, 0.85, 0.71], obstacle_position=[0.45, 0.55, 0.55],
                                  initial_joint_positions=[0.9, 0.45, 0, 0, 0, 0],
                                  initial_positions_variation_range=[0, 0, .5, .5, .5, .5], visualize=False, n_envs=256)
        mf.initialize_naf_agent(checkpoint_frequency=400, seed=0)
        before = torch.cat([p.detach().reshape(-1) for p in mf.naf_agent.qnetwork_main.parameters()]).clone()
        scores = mf.run_training(900, 20, verbose=False)
        torch.cuda.synchronize()
        assert sorted(scores) == list(range(1, 901))
        frames = np.array([v[1] for v in scores.values()])
        returns = np.array([v[0] for v in scores.values()])
        assert frames.min() >= 1 and frames.max() == 20 and np.isfinite(returns).all()
        after = torch.cat([p.detach().reshape(-1) for p in mf.naf_agent.qnetwork_main.parameters()])
        assert torch.isfinite(after).all() and float((after - before).abs().max()) > 1e-4      # it did learn
        for ep in (400, 800):
            assert os.path.isfile(f'checkpoints/{ep}/weights.p') and os.path.isfile(f'checkpoints/{ep}/scores.txt')
        saved = json.load(open('checkpoints/800/scores.txt'))
        assert len(saved) == 900 and saved['800'][1] >= 1 and saved['900'] == [0, 0]    # later episodes still placeholders
        sd = torch.load('model.p')
        assert len(sd) == 20 and all(not v.is_cuda for v in sd.values())
        assert len(mf.naf_agent.memory) > 64 and mf.naf_agent._tick_base > 0
    finally:
        mf.set_log_level(logging.INFO)
