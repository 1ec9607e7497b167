"""
CPU tests pinning the NAF oracle (oracle/naf_restatement.py): against the reference's own golden vector,
against the fixtures generated from the unmodified reference (tests/golden/make_naf_golden.py), and — in the
container that has /root/reference — against the live reference classes.
"""
import os

import numpy as np
import pytest
import torch

from oracle.naf_restatement import NAFRef, learn_ref, load_reference_classes

from helpers import assert_params_close

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
DEMO = os.path.join(os.path.dirname(GOLD), '..', 'robotic_manipulator_rloa_b200', 'naf_components', 'demo_weights')


def test_reference_golden_vector():
    """tests/robotic_manipulator_rloa/naf_components/test_naf_neural_network.py:53-67 (literals from there)."""
    net = NAFRef(10, 5, 256, seed=0)
    net.train()
    states = torch.tensor([[0., 1, 2, 3, 4, 5, 6, 7, 8, 9], [10, 11, 12, 13, 14, 15, 16, 17, 18, 19]])
    actions = torch.tensor([[0, 1, 2, 3, 4], [10, 11, 12, 13, 14]])
    with torch.no_grad():
        _, _, q, v = net.heads(states, actions)
    np.testing.assert_allclose(q.numpy(), [[-35.50931930541992], [-638.494873046875]], rtol=2e-5)
    np.testing.assert_allclose(v.numpy(), [[0.5665180683135986], [-0.08311141282320023]], rtol=5e-5)
    g = np.load(os.path.join(GOLD, 'naf_golden_vector.npz'))
    np.testing.assert_allclose(q.numpy(), g['q'], rtol=1e-6)
    np.testing.assert_allclose(v.numpy(), g['v'], rtol=1e-5)


def test_demo_weights_layout_and_forward_fixture():
    sd = torch.load(os.path.join(DEMO, 'weights_kuka.p'))
    assert len(sd) == 20 and list(sd)[0] == 'input_layer.weight' and sd['bn1.num_batches_tracked'].item() == 775017
    g = np.load(os.path.join(GOLD, 'naf_forward_kuka.npz'))
    net = NAFRef(21, 6)
    net.load_state_dict(sd)
    s, a = torch.tensor(g['states']), torch.tensor(g['actions'])
    net.eval()
    with torch.no_grad():
        mu, P, q, v = net.heads(s, a)
    np.testing.assert_allclose(mu.numpy(), g['eval_mu'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(torch.diagonal(P, dim1=1, dim2=2).numpy(), g['eval_pdiag'], rtol=1e-5)
    np.testing.assert_allclose(q.numpy(), g['eval_q'], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(v.numpy(), g['eval_v'], rtol=1e-5, atol=1e-4)
    assert np.abs(P.numpy() - np.diag(np.ones(6))[None] * P.numpy()).max() == 0.0     # P is diagonal (B.1)
    net.train()
    with torch.no_grad():
        mu, P, q, v = net.heads(s, a)
    np.testing.assert_allclose(q.numpy(), g['train_q'], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(net.bn2.running_var.numpy(), g['after.bn2.running_var'], rtol=1e-6)


# the running means absorb the (noise-driven) linear biases, so they inherit lr-sized differences
LOOSE = {'bn1.running_mean': dict(atol=1e-3, max_frac=0.0), 'bn2.running_mean': dict(atol=1e-3, max_frac=0.0)}


def test_learn_fixture():
    g = np.load(os.path.join(GOLD, 'naf_learn_seed0.npz'))
    main, target = NAFRef(21, 6), NAFRef(21, 6)
    main.load_state_dict({k[len('main.before.'):]: torch.tensor(g[k]) for k in g.files if k.startswith('main.before.')})
    target.load_state_dict({k[len('target.before.'):]: torch.tensor(g[k]) for k in g.files if k.startswith('target.before.')})
    opt = torch.optim.Adam(main.parameters(), lr=1e-3)
    batch = (torch.tensor(g['states']), torch.tensor(g['actions']).long(), torch.tensor(g['rewards']),
             torch.tensor(g['next_states']), torch.tensor(g['dones']))
    for step in range(3):
        learn_ref(main, target, opt, batch, gamma=0.99, tau=1e-3)
        for k, v in main.state_dict().items():
            if k in ('input_layer.bias', 'hidden_layer.bias'):
                continue      # gradient is rounding noise (BatchNorm removes the bias): Adam amplifies it, not comparable
            assert_params_close(v.numpy(), g[f'main.after{step + 1}.{k}'], k, **LOOSE.get(k, {}))
        for k, v in target.state_dict().items():
            if k in ('input_layer.bias', 'hidden_layer.bias'):
                continue
            assert_params_close(v.numpy(), g[f'target.after{step + 1}.{k}'], k, **LOOSE.get(k, {}))


@pytest.mark.skipif(not os.path.isdir('/root/reference/robotic_manipulator_rloa'), reason='reference not mounted')
def test_against_live_reference():
    NAF, NAFAgent, ReplayBuffer = load_reference_classes()
    ref = NAF(21, 6, 256, 0, torch.device('cpu'))
    mine = NAFRef(21, 6, 256, seed=0)
    for (k1, v1), (k2, v2) in zip(ref.state_dict().items(), mine.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)          # same names, same seeded initialisation
    g = torch.Generator().manual_seed(5)
    s = torch.randn(32, 21, generator=g)
    a = torch.randint(-1, 2, (32, 6), generator=g)
    ref.train(); mine.train()
    with torch.no_grad():
        _, q_ref, v_ref = ref(s, a)
        _, _, q, v = mine.heads(s, a)
    assert torch.allclose(q, q_ref, rtol=1e-6, atol=1e-6) and torch.allclose(v, v_ref, rtol=1e-6, atol=1e-6)
    # the reference's sampler: int64 actions in {-1, 0, 1}
    rb = ReplayBuffer(100, 8, torch.device('cpu'), 0)
    for i in range(20):
        rb.add(np.zeros(21), np.array([0.99, -1.0, 1.0, 0.3, -0.7, 1.0]), -1.0, np.zeros(21), 0)
    acts = rb.sample()[1]
    assert acts.dtype == torch.int64 and set(acts.unique().tolist()) <= {-1, 0, 1}
