"""
NAF network — drop-in for /root/reference/robotic_manipulator_rloa/naf_components/naf_neural_network.py.

The class is still an ``nn.Module`` with the reference's submodule names (``input_layer, bn1,
hidden_layer, bn2, action_values, value, matrix_entries``), so ``state_dict()`` / ``load_state_dict()``
round-trip the 20-entry checkpoint layout (``weights_kuka.p`` loads unchanged) and seeded initialisation
draws the same numbers as the reference.  The arithmetic of ``forward`` is NOT torch: the parameters
are handed to librloa_b200.so by pointer (rloa_naf_forward) — trunk GEMMs, BatchNorm, the fused head
(mu, V, diag L, advantage).  There is no CPU fallback: calling ``forward`` without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Optional, Tuple

import torch
from torch import nn

from .. import _native as N


class NafWorkspace:
    """Owns an rloa_naf_ws handle (activations / gradient scratch) sized for ``max_batch`` rows."""

    def __init__(self, state_size: int, action_size: int, hidden: int, max_batch: int, device: torch.device):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise N.NativeLibraryError('the NAF network runs in librloa_b200.so (sm_100a): a CUDA device is required, '
                                       'there is no CPU fallback')
        self.lib = N.lib()
        self.shape = (state_size, action_size, hidden)
        self.max_batch = int(max_batch)
        self.trunk_mode = 0
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(self.lib.rloa_naf_ws_create(state_size, action_size, hidden, self.max_batch, C.byref(self._h)),
                    'rloa_naf_ws_create')

    @property
    def handle(self):
        return self._h

    def set_trunk(self, mode: int) -> None:
        """0 = fp32 CUDA-core trunk, 1 = tcgen05 tensor-core trunk (bf16 operands)."""
        N.check(self.lib.rloa_naf_ws_set_trunk(self._h, int(mode)), 'rloa_naf_ws_set_trunk')
        self.trunk_mode = int(mode)

    def close(self) -> None:
        if self._h:
            torch.cuda.synchronize(self.device)
            self.lib.rloa_naf_ws_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NAF(nn.Module):

    def __init__(self, state_size: int, action_size: int, layer_size: int, seed: int, device: torch.device) -> None:
        super().__init__()
        self.seed = torch.manual_seed(seed)          # same RNG protocol as the reference (:33)
        self.state_size = state_size
        self.action_size = action_size
        self.layer_size = layer_size
        self.device = device
        self.input_layer = nn.Linear(in_features=state_size, out_features=layer_size)
        self.bn1 = nn.BatchNorm1d(layer_size)
        self.hidden_layer = nn.Linear(in_features=layer_size, out_features=layer_size)
        self.bn2 = nn.BatchNorm1d(layer_size)
        self.action_values = nn.Linear(in_features=layer_size, out_features=action_size)
        self.value = nn.Linear(in_features=layer_size, out_features=1)
        self.matrix_entries = nn.Linear(in_features=layer_size, out_features=int(action_size * (action_size + 1) / 2))
        self._ws: Optional[NafWorkspace] = None
        # reference-exact by default; True replays actions as stored instead of truncating them (B.2)
        self.trunk_mode = 0

    # ------------------------------------------------------------------------------------------
    def native_params(self) -> N.NafParams:
        """rloa_naf_params view of this module's tensors (pointers only; the module keeps ownership)."""
        p = N.NafParams()
        p.state_size, p.action_size, p.hidden = self.state_size, self.action_size, self.layer_size
        tensors = [self.input_layer.weight, self.input_layer.bias, self.bn1.weight, self.bn1.bias,
                   self.bn1.running_mean, self.bn1.running_var, self.bn1.num_batches_tracked,
                   self.hidden_layer.weight, self.hidden_layer.bias, self.bn2.weight, self.bn2.bias,
                   self.bn2.running_mean, self.bn2.running_var, self.bn2.num_batches_tracked,
                   self.action_values.weight, self.action_values.bias, self.value.weight, self.value.bias,
                   self.matrix_entries.weight, self.matrix_entries.bias]
        names = ['w1', 'b1', 'bn1_w', 'bn1_b', 'bn1_mean', 'bn1_var', 'bn1_batches', 'w2', 'b2', 'bn2_w', 'bn2_b',
                 'bn2_mean', 'bn2_var', 'bn2_batches', 'w_mu', 'b_mu', 'w_v', 'b_v', 'w_l', 'b_l']
        for name, t in zip(names, tensors):
            if not t.is_cuda:
                raise N.NativeLibraryError('NAF parameters must live on a CUDA device: the network runs in '
                                           'librloa_b200.so (sm_100a) and has no CPU fallback')
            if not t.is_contiguous():
                raise N.NativeLibraryError(f'NAF parameter {name} is not contiguous')
            setattr(p, name, t.data_ptr())
        return p

    def workspace(self, batch: int) -> NafWorkspace:
        dev = self.input_layer.weight.device
        if self._ws is None or self._ws.max_batch < batch or self._ws.device != dev:
            if self._ws is not None:
                self._ws.close()
            size = max(256, 1 << (int(batch) - 1).bit_length())
            self._ws = NafWorkspace(self.state_size, self.action_size, self.layer_size, size, dev)
            if self.trunk_mode:
                self._ws.set_trunk(self.trunk_mode)
        return self._ws

    def set_trunk_mode(self, mode: int) -> None:
        self.trunk_mode = int(mode)
        if self._ws is not None:
            self._ws.set_trunk(self.trunk_mode)

    def heads(self, input_: torch.Tensor, action: Optional[torch.Tensor] = None, trunc_action: bool = False):
        """(mu [B,A], diag P [B,A], Q [B,1] | None, V [B,1]) through rloa_naf_forward; BN mode = self.training."""
        dev = self.input_layer.weight.device
        x = input_.to(device=dev, dtype=torch.float32).contiguous()
        B = x.shape[0]
        a = None
        if action is not None:
            a = action.to(device=dev, dtype=torch.float32).contiguous()
        ws = self.workspace(B)
        f32 = dict(dtype=torch.float32, device=dev)
        mu = torch.empty(B, self.action_size, **f32)
        pd = torch.empty(B, self.action_size, **f32)
        v = torch.empty(B, 1, **f32)
        q = torch.empty(B, 1, **f32) if a is not None else None
        p = self.native_params()
        N.check(ws.lib.rloa_naf_forward(ws.handle, C.byref(p), x.data_ptr(), N.ptr(a), B, int(self.training),
                                        int(trunc_action), mu.data_ptr(), pd.data_ptr(), N.ptr(q), v.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream), 'rloa_naf_forward')
        return mu, pd, q, v

    def forward(self, input_: torch.Tensor, action: Optional[torch.Tensor] = None
                ) -> Tuple[torch.Tensor, Optional[Any], Any]:
        """Same contract as the reference forward (:56-123): (sampled action clamped to [-1,1], Q | None, V).
        The sample is mu + eps / sqrt(P_kk): P is diagonal (elementwise L o L^T), so N(mu, P^-1) factorises."""
        mu, pd, q, v = self.heads(input_, action)
        noise = torch.randn_like(mu) * torch.rsqrt(pd)
        return torch.clamp(mu + noise, min=-1, max=1), q, v
