"""
Generates tests/golden/naf_*.npz by importing the UNMODIFIED reference from /root/reference
(pybullet / pybullet_data / matplotlib stubbed in sys.modules) — run in this container only:

    cd /tmp && python /root/repo/tests/golden/make_naf_golden.py

Fixtures (inputs AND reference outputs, fp32):
  naf_forward_kuka.npz  : weights_kuka.p demo weights, eval- and train-mode forward on 64 seeded states
  naf_forward_seed0.npz : NAF(21, 6, 256, seed=0) init, train-mode forward, B = 128
  naf_learn_seed0.npz   : one NAFAgent.learn() on a seeded batch (B = 128): every parameter / BN buffer of main and
                          target nets before and after, loss
  naf_golden_vector.npz : the inputs of tests/naf_components/test_naf_neural_network.py:53-67 with the outputs of
                          this torch build (the test's own literals are kept in tests/test_naf_oracle.py)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
from oracle.naf_restatement import load_reference_classes   # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
NAF, NAFAgent, ReplayBuffer = load_reference_classes()
dev = torch.device('cpu')
torch.set_num_threads(1)


def sd_np(net, prefix):
    return {f'{prefix}{k}': v.detach().numpy().copy() for k, v in net.state_dict().items()}


def heads(net, s, a=None):
    """mu, diag(P), Q, V exactly as NAF.forward computes them, with the sampling skipped."""
    with torch.no_grad():
        x = torch.relu(net.bn1(net.input_layer(s)))
        x = torch.relu(net.bn2(net.hidden_layer(x)))
        mu = torch.tanh(net.action_values(x))
        ent = torch.tanh(net.matrix_entries(x))
        V = net.value(x)
        A = net.action_size
        L = torch.zeros(s.shape[0], A, A)
        idx = torch.tril_indices(A, A)
        L[:, idx[0], idx[1]] = ent
        L.diagonal(dim1=1, dim2=2).exp_()
        P = L * L.transpose(2, 1)
    return mu.numpy(), torch.diagonal(P, dim1=1, dim2=2).numpy().copy(), V.numpy()


# --- demo weights, eval + train forward ------------------------------------------------------------------
g = torch.Generator().manual_seed(123)
net = NAF(21, 6, 256, 0, dev)
net.load_state_dict(torch.load('/root/reference/robotic_manipulator_rloa/naf_components/demo_weights/weights_kuka.p'))
s = torch.randn(64, 21, generator=g) * 0.7
a = torch.clamp(torch.randn(64, 6, generator=g), -1, 1)
out = {'states': s.numpy(), 'actions': a.numpy()}
net.eval()
mu, pd, V = heads(net, s)
with torch.no_grad():
    _, Q, V2 = net(s, a)
out.update(eval_mu=mu, eval_pdiag=pd, eval_v=V, eval_q=Q.numpy())
assert np.allclose(V, V2.numpy())
net.train()
before = sd_np(net, 'before.')
with torch.no_grad():
    _, Q, V2 = net(s, a)
after = sd_np(net, 'after.')
net2 = NAF(21, 6, 256, 0, dev); net2.load_state_dict({k[7:]: torch.tensor(v) for k, v in before.items()}); net2.train()
mu, pd, V = heads(net2, s)
out.update(train_mu=mu, train_pdiag=pd, train_v=V, train_q=Q.numpy())
out.update({k: v for k, v in after.items() if 'running' in k or 'num_batches' in k})
np.savez_compressed(os.path.join(OUT, 'naf_forward_kuka.npz'), **out)

# --- seed-0 init, train-mode forward with long-cast actions ----------------------------------------------
net = NAF(21, 6, 256, 0, dev); net.train()
s = torch.randn(128, 21, generator=g)
a = torch.clamp(torch.randn(128, 6, generator=g) * 1.5, -1, 1)
with torch.no_grad():
    _, Q, V = net(s, a.long())
net3 = NAF(21, 6, 256, 0, dev); net3.train()
mu, pd, _ = heads(net3, s)
np.savez_compressed(os.path.join(OUT, 'naf_forward_seed0.npz'), states=s.numpy(), actions=a.numpy(), mu=mu, pdiag=pd,
                    q=Q.numpy(), v=V.numpy())

# --- one learn() ------------------------------------------------------------------------------------------
os.chdir('/tmp')
agent = NAFAgent(environment=None, state_size=21, action_size=6, layer_size=256, batch_size=128, buffer_size=1000,
                 learning_rate=1e-3, tau=1e-3, gamma=0.99, update_freq=1, num_updates=1, checkpoint_frequency=500,
                 device=dev, seed=0)
# make main and target differ, like mid-training
with torch.no_grad():
    for p_ in agent.qnetwork_target.parameters():
        p_.add_(0.01 * torch.randn(p_.shape, generator=g))
B = 128
s = torch.randn(B, 21, generator=g)
s2 = s + 0.05 * torch.randn(B, 21, generator=g)
a = torch.clamp(torch.randn(B, 6, generator=g) * 1.5, -1, 1)
r = -torch.rand(B, 1, generator=g)
d = (torch.rand(B, 1, generator=g) < 0.05).float()
fix = {'states': s.numpy(), 'actions': a.numpy(), 'rewards': r.numpy(), 'next_states': s2.numpy(), 'dones': d.numpy()}
fix.update(sd_np(agent.qnetwork_main, 'main.before.'))
fix.update(sd_np(agent.qnetwork_target, 'target.before.'))
for step in range(3):      # three consecutive updates on the same batch: exercises Adam's moment/bias-correction path
    agent.learn((s, a.long(), r, s2, d))
    fix.update(sd_np(agent.qnetwork_main, f'main.after{step + 1}.'))
    fix.update(sd_np(agent.qnetwork_target, f'target.after{step + 1}.'))
np.savez_compressed(os.path.join(OUT, 'naf_learn_seed0.npz'), **fix)

# --- the reference's own golden inputs --------------------------------------------------------------------
net = NAF(10, 5, 256, 0, dev); net.train()
gs = torch.tensor([[0., 1, 2, 3, 4, 5, 6, 7, 8, 9], [10, 11, 12, 13, 14, 15, 16, 17, 18, 19]])
ga = torch.tensor([[0, 1, 2, 3, 4], [10, 11, 12, 13, 14]])
with torch.no_grad():
    _, Q, V = net(gs, ga)
np.savez_compressed(os.path.join(OUT, 'naf_golden_vector.npz'), states=gs.numpy(), actions=ga.numpy().astype(np.float32),
                    q=Q.numpy(), v=V.numpy())
print('golden q', Q.numpy().ravel(), 'v', V.numpy().ravel())
