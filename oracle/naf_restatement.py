"""
ORACLE — TEST INFRASTRUCTURE ONLY.  CPU fp32 torch restatement of the reference's NAF network and
NAFAgent.learn; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it.  The product package never does.

Pinned (tests/test_naf_oracle.py) against
  * the reference's own golden vector, tests/robotic_manipulator_rloa/naf_components/test_naf_neural_network.py:53-67
    (rtol 2e-5: the literal was produced by an older torch/MKL);
  * tests/golden/naf_*.npz, produced by importing the UNMODIFIED reference classes from /root/reference
    (tests/golden/make_naf_golden.py; pybullet / matplotlib stubbed in sys.modules);
  * the live reference import whenever /root/reference exists (this container only).

Follows /root/reference/robotic_manipulator_rloa/naf_components/naf_neural_network.py:41-123 (network,
L / P / advantage construction) and naf_algorithm.py:180-226 (learn, soft_update), including the
behaviour-defining quirks listed in SURVEY.md Appendix B.
"""
from collections import OrderedDict

import torch
from torch import nn

PARAM_NAMES = ['input_layer.weight', 'input_layer.bias', 'bn1.weight', 'bn1.bias', 'hidden_layer.weight',
               'hidden_layer.bias', 'bn2.weight', 'bn2.bias', 'action_values.weight', 'action_values.bias',
               'value.weight', 'value.bias', 'matrix_entries.weight', 'matrix_entries.bias']


class NAFRef(nn.Module):
    """Same submodule names / creation order as the reference so state_dicts and seeded init coincide."""

    def __init__(self, state_size, action_size, layer_size=256, seed=0):
        super().__init__()
        torch.manual_seed(seed)                                   # naf_neural_network.py:33
        self.action_size = action_size
        self.input_layer = nn.Linear(state_size, layer_size)      # :41
        self.bn1 = nn.BatchNorm1d(layer_size)                     # :42
        self.hidden_layer = nn.Linear(layer_size, layer_size)     # :45
        self.bn2 = nn.BatchNorm1d(layer_size)                     # :46
        self.action_values = nn.Linear(layer_size, action_size)   # :49
        self.value = nn.Linear(layer_size, 1)                     # :51
        self.matrix_entries = nn.Linear(layer_size, action_size * (action_size + 1) // 2)   # :53-54

    def trunk(self, x):
        x = torch.relu(self.bn1(self.input_layer(x)))             # :76
        return torch.relu(self.bn2(self.hidden_layer(x)))         # :78

    def heads(self, states, action=None):
        """(mu [B,A], P [B,A,A], Q [B,1] | None, V [B,1]) — everything of forward() except the sampling."""
        x = self.trunk(states)
        mu = torch.tanh(self.action_values(x))                    # :81
        entries = torch.tanh(self.matrix_entries(x))              # :84
        V = self.value(x)                                         # :87
        B, A = states.shape[0], self.action_size
        L = torch.zeros(B, A, A, dtype=states.dtype).to(states.device)   # :95 host allocation + copy, as the reference
        rows, cols = torch.tril_indices(A, A)                     # :98 row-major lower-triangular order
        L[:, rows, cols] = entries
        diag = torch.arange(A)
        L[:, diag, diag] = torch.exp(L[:, diag, diag])            # :102
        P = L * L.transpose(1, 2)                                 # :104 elementwise, NOT a matrix product
        Q = None
        if action is not None:
            d = (action.to(states.dtype) - mu).unsqueeze(-1)      # :111-113
            Q = -0.5 * (d.transpose(1, 2) @ P @ d).squeeze(-1) + V
        return mu, P, Q, V

    def forward(self, states, action=None):
        mu, P, Q, V = self.heads(states, action)
        dist = torch.distributions.MultivariateNormal(mu, torch.inverse(P))   # :119
        return torch.clamp(dist.sample(), -1, 1), Q, V           # :120-121


def learn_ref(main: NAFRef, target: NAFRef, opt: torch.optim.Adam, batch, gamma: float, tau: float):
    """One NAFAgent.learn (naf_algorithm.py:180-213).  batch = (states, actions, rewards [B,1], next_states,
    dones) with actions ALREADY cast the way ReplayBuffer.sample does (int64, replay_buffer.py:60).
    Returns (loss, grad-norm before clipping, flat gradient before clipping)."""
    states, actions, rewards, next_states, _dones = batch
    opt.zero_grad()
    main.train(); target.train()                                  # both nets: train-mode BatchNorm
    with torch.no_grad():
        _, _, _, v_next = target.heads(next_states)               # :194-195
    y = rewards + gamma * v_next                                  # :199 (no done mask)
    _, _, q, _ = main.heads(states, actions)                      # :202
    loss = torch.nn.functional.mse_loss(q, y)                     # :205
    loss.backward()                                               # :208
    flat = torch.cat([p.grad.reshape(-1) for p in main.parameters()]).clone()
    norm = torch.nn.utils.clip_grad_norm_(main.parameters(), 1)   # :209
    opt.step()                                                    # :210
    with torch.no_grad():                                         # :225-226 parameters only, no BN buffers
        for pt, pm in zip(target.parameters(), main.parameters()):
            pt.copy_(tau * pm + (1.0 - tau) * pt)
    return float(loss.detach()), float(norm), flat


def load_reference_classes():
    """The UNMODIFIED reference NAF / NAFAgent / ReplayBuffer (only where /root/reference exists)."""
    import os
    import sys
    from unittest.mock import MagicMock
    if not os.path.isdir('/root/reference/robotic_manipulator_rloa'):
        return None
    for m in ('pybullet', 'pybullet_data', 'matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(m, MagicMock())
    if '/root/reference' not in sys.path:
        sys.path.insert(0, '/root/reference')
    from robotic_manipulator_rloa.naf_components.naf_algorithm import NAFAgent
    from robotic_manipulator_rloa.naf_components.naf_neural_network import NAF
    from robotic_manipulator_rloa.utils.replay_buffer import ReplayBuffer
    return NAF, NAFAgent, ReplayBuffer


def state_dict_like_reference(net: NAFRef) -> OrderedDict:
    return OrderedDict((k, v.detach().clone()) for k, v in net.state_dict().items())
