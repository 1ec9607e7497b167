"""
tcgen05 operand layouts (csrc/umma_probe.cu): a bf16 tile staged once in the K-major SWIZZLE_128B layout, consumed K-major
and — untransposed — MN-major, against a plain matrix product of the same bf16-rounded operands.  These are the views the
fused learn kernel uses for the backward contractions (da = dz W, dW = dz^T a).
"""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('mode', [0, 1, 2, 3])
def test_umma_operand_views(mode):
    from robotic_manipulator_rloa_b200 import _native as N
    lib = N.lib()
    g = torch.Generator().manual_seed(100 + mode)
    a = torch.randn(128, 256, generator=g)
    b = torch.randn(*{0: (256, 256), 1: (256, 256), 2: (128, 256), 3: (128, 64)}[mode], generator=g)
    ab, bb = a.bfloat16().float(), b.bfloat16().float()
    want = {0: lambda: ab @ bb.t(), 1: lambda: ab @ bb, 2: lambda: ab.t() @ bb, 3: lambda: ab.t() @ bb}[mode]()
    da, db = a.cuda().contiguous(), b.cuda().contiguous()
    out = torch.full(tuple(want.shape), float('nan'), device='cuda')
    N.check(lib.rloa_umma_probe(mode, da.data_ptr(), db.data_ptr(), out.data_ptr(),
                                torch.cuda.current_stream().cuda_stream), 'rloa_umma_probe')
    torch.cuda.synchronize()
    err = (out.cpu() - want).abs().max().item()
    print(f'mode {mode}: max |err| = {err:.3e} (|C| max {want.abs().max().item():.1f})')
    assert err <= 2e-3 * want.abs().max().item()
