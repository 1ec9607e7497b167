"""
GPU parity of the tcgen05 tensor-core trunk (csrc/naf_trunk_tc.cu) through the C ABI.

Two bounds, both written here (north_star: "stated looser bound for bf16 tensor-core GEMMs"):
  * against a torch fp32 evaluation of the SAME bf16-rounded operands the kernel must agree to 2e-5 of the
    output scale — only the fp32 accumulation order differs, so this pins the shared-memory swizzle, the UMMA
    descriptors and the TMEM read-back exactly;
  * against the all-fp32 hidden layer (the reference arithmetic) the bound is the bf16 operand rounding:
    |err| <= 8e-3 x output scale for K = 256 (2^-9 relative per operand, random-sign accumulation).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle.naf_restatement import NAFRef, learn_ref

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
H = 256


def _hidden_layer(ws, z1, scale, shift, w2, b2):
    from robotic_manipulator_rloa_b200 import _native as N
    z2 = torch.full((z1.shape[0], H), float('nan'), device=DEV)
    N.check(ws.lib.rloa_naf_hidden_layer(ws.handle, z1.data_ptr(), scale.data_ptr(), shift.data_ptr(), w2.data_ptr(),
                                         b2.data_ptr(), z2.data_ptr(), z1.shape[0],
                                         torch.cuda.current_stream().cuda_stream), 'rloa_naf_hidden_layer')
    torch.cuda.synchronize()
    return z2


@pytest.mark.parametrize('B', [1, 77, 128, 129, 1024, 4096])
def test_hidden_layer_tcgen05_vs_same_operands(B):
    from robotic_manipulator_rloa_b200.naf_components.naf_neural_network import NafWorkspace
    g = torch.Generator().manual_seed(B)
    z1 = torch.randn(B, H, generator=g).to(DEV)
    scale = (0.5 + torch.rand(H, generator=g)).to(DEV)
    shift = (0.2 * torch.randn(H, generator=g)).to(DEV)
    w2 = (torch.randn(H, H, generator=g) / 16).to(DEV)
    b2 = torch.randn(H, generator=g).to(DEV)
    ws = NafWorkspace(21, 6, H, max(B, 256), DEV)
    z2_fp32 = _hidden_layer(ws, z1, scale, shift, w2, b2)
    # the kernel's prologue is one fmaf per element: a double product + sum rounded once to fp32 reproduces it,
    # so the bf16 rounding below sees bit-identical inputs (a separate fp32 mul + add flips ~3e-5 of them)
    a = torch.relu((z1.double() * scale.double() + shift.double()).float())
    want_fp32 = a.double() @ w2.double().t() + b2.double()
    assert (z2_fp32.double() - want_fp32).abs().max() <= 1e-5 * want_fp32.abs().max()
    ws.set_trunk(1)
    z2 = _hidden_layer(ws, z1, scale, shift, w2, b2)
    assert torch.isfinite(z2).all()
    want_same = a.bfloat16().double() @ w2.bfloat16().double().t() + b2.double()
    s = float(want_fp32.abs().max())
    err_same = float((z2.double() - want_same).abs().max())
    err_fp32 = float((z2.double() - want_fp32).abs().max())
    assert err_same <= 2e-5 * s, f'vs bf16-rounded operands: {err_same:.3e} (scale {s:.3g})'
    assert err_fp32 <= 8e-3 * s, f'vs fp32: {err_fp32:.3e} (scale {s:.3g})'
    ws.close()


def test_forward_and_learn_with_tensor_core_trunk():
    """NAF.forward / NAFAgent.learn with trunk mode 1 against the fp32 restatement, bf16 bound."""
    from robotic_manipulator_rloa_b200.naf_components.naf_algorithm import NAFAgent
    S, A, B = 21, 6, 1024
    ref_main, ref_target = NAFRef(S, A, H, seed=1), NAFRef(S, A, H, seed=2)
    agent = NAFAgent(None, S, A, H, B, 1000, 1e-3, 1e-3, 0.99, 1, 1, 500, DEV, 0)
    agent.qnetwork_main.load_state_dict(ref_main.state_dict())
    agent.qnetwork_target.load_state_dict(ref_target.state_dict())
    agent.set_trunk_mode(1)
    g = torch.Generator().manual_seed(7)
    s = torch.randn(B, S, generator=g); s2 = s + 0.1 * torch.randn(B, S, generator=g)
    a = torch.clamp(torch.randn(B, A, generator=g) * 1.5, -1, 1)
    r = -torch.rand(B, 1, generator=g); d = torch.zeros(B, 1)
    ref_main.eval(); agent.qnetwork_main.eval()
    with torch.no_grad():
        mu_r, P_r, q_r, v_r = ref_main.heads(s, a.long().float())
    mu, pd, q, v = agent.qnetwork_main.heads(s, a, trunc_action=True)
    assert (mu.cpu() - mu_r).abs().max() <= 1e-2
    assert (v.cpu() - v_r).abs().max() <= 1e-2 * max(1.0, float(v_r.abs().max()))
    assert (q.cpu() - q_r).abs().max() <= 2e-2 * max(1.0, float(q_r.abs().max()))
    ref_main.train(); agent.qnetwork_main.train()
    opt = torch.optim.Adam(ref_main.parameters(), lr=1e-3)
    loss_r, norm_r, flat_r = learn_ref(ref_main, ref_target, opt, (s, a.long(), r, s2, d), 0.99, 1e-3)
    agent.learn((s, a.long(), r, s2, d))
    b = agent._learn_buffers()
    assert abs(float(b['loss'].item()) - loss_r) <= 2e-2 * abs(loss_r)
    got, want = b['grad'].cpu().numpy(), flat_r.numpy()
    cos = float(np.dot(got, want) / (np.linalg.norm(got) * np.linalg.norm(want)))
    assert cos >= 0.999, cos


@pytest.mark.parametrize('S,A,n', [(21, 6, 1), (21, 6, 100), (21, 6, 128), (23, 7, 4096), (21, 6, 5000)])
def test_fused_policy_kernel_matches_fp32_act(S, A, n):
    """NAFAgent.act through the fused tcgen05 policy kernel (trunk mode 1: one launch, W2 / heads in bf16, layer 1
    in tf32 with the observations split hi + lo) against the fp32 kernels (mode 0) on the same states, weights, running BatchNorm statistics and Philox
    keys.  Bound: bf16 operand rounding through two 256-long contractions ahead of a tanh — |d mu| <= 1e-2; with
    noise the same eps is drawn, so clamped actions agree to 2e-2."""
    from robotic_manipulator_rloa_b200.naf_components.naf_algorithm import NAFAgent
    ref = NAFRef(S, A, H, seed=5)
    g = torch.Generator().manual_seed(n)
    with torch.no_grad():                                   # non-trivial eval-mode statistics and affine terms
        for bn in (ref.bn1, ref.bn2):
            bn.running_mean.copy_(0.3 * torch.randn(H, generator=g))
            bn.running_var.copy_(0.5 + torch.rand(H, generator=g))
            bn.weight.copy_(0.8 + 0.4 * torch.rand(H, generator=g))
            bn.bias.copy_(0.1 * torch.randn(H, generator=g))
    agent = NAFAgent(None, S, A, H, 128, 1000, 1e-3, 1e-3, 0.99, 1, 1, 500, DEV, 3)
    agent.qnetwork_main.load_state_dict(ref.state_dict())
    states = torch.randn(n, S, generator=g).to(DEV)
    out = {}
    for mode in (0, 1):
        agent.set_trunk_mode(mode)
        agent.noise_scale = 0.0
        agent._act_calls = 7
        mu = agent.act_batch(states).clone()
        agent.noise_scale = 1.0
        agent._act_calls = 7
        noisy = agent.act_batch(states).clone()
        out[mode] = (mu, noisy)
    torch.cuda.synchronize()
    ref.eval()
    with torch.no_grad():
        mu_ref = ref.heads(states.cpu())[0]
    assert (out[0][0].cpu() - mu_ref).abs().max() <= 1e-5
    assert torch.isfinite(out[1][0]).all() and torch.isfinite(out[1][1]).all()
    d_mu = float((out[1][0] - out[0][0]).abs().max())
    d_act = float((out[1][1] - out[0][1]).abs().max())
    assert d_mu <= 1e-2, d_mu
    assert d_act <= 2e-2, d_act
    assert out[1][1].abs().max() <= 1.0
