"""kld-net inference on the CUDA path: the line-detection U-Net whose output becomes the movement-group
masks of the IM-MoCo fit (src/models/kld_net.py:4-11 ``get_unet`` -> fastmri.models.Unet, the network of
src/models/unet.py:17-187 with InstanceNorm2d; used at src/test/test_immoco.py:17-20,50-61).

``get_unet(in_chans, out_chans, chans, num_pool_layers, drop_prob)`` keeps the reference signature and
returns a module whose ``state_dict()`` has fastmri's keys and shapes, so
``net.load_state_dict(torch.load("kLDNet.pth"))`` works unchanged.  ``forward`` is inference only and runs
hand-written kernels (csrc/unet_tc.cu: the 3x3 convolutions as tcgen05 implicit GEMMs with the 3xTF32 split;
csrc/unet.cu: the 2-channel input layer, conv-transpose, norm, head) with fused instance statistics, one
normalise + LeakyReLU (+ 2x2 average pool) pass per layer, the channel concat folded into the next
convolution's loads.  No cuDNN, no eager fallback.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _native as nat
from .motion_utils import extract_movement_groups, movement_group_labels  # noqa: F401
from .ops import IFFT, _need_cuda, _stream

_EPS, _SLOPE = 1e-5, 0.2


class _Scope(nn.Module):
    """Name-space node: parameters are registered under fastmri's dotted state-dict keys."""


def _block_keys(prefix: str) -> Tuple[str, str]:
    return f"{prefix}.layers.0.weight", f"{prefix}.layers.4.weight"


class Unet(nn.Module):
    def __init__(self, in_chans: int, out_chans: int, chans: int = 32, num_pool_layers: int = 4,
                 drop_prob: float = 0.0):
        super().__init__()
        if num_pool_layers < 1:
            raise ValueError("num_pool_layers must be >= 1")
        self.in_chans, self.out_chans, self.chans = in_chans, out_chans, chans
        self.num_pool_layers, self.drop_prob = num_pool_layers, drop_prob
        self.tensor_cores = True      # 3x3 convolutions on tcgen05 (False: the fp32 SIMT kernels, A/B checks)
        self._pack_cache: dict = {}
        self._keys: List[str] = []
        # encoder / bottleneck / decoder convolution pairs, in fastmri's registration order
        ch = chans
        self._down = [self._add_block("down_sample_layers.0", in_chans, chans)]
        for i in range(1, num_pool_layers):
            self._down.append(self._add_block(f"down_sample_layers.{i}", ch, ch * 2))
            ch *= 2
        self._mid = self._add_block("conv", ch, ch * 2)
        self._up, up_t = [], []
        for i in range(num_pool_layers):
            prefix = f"up_conv.{i}" if i < num_pool_layers - 1 else f"up_conv.{i}.0"
            self._up.append(self._add_block(prefix, ch * 2, ch))
            up_t.append((f"up_transpose_conv.{i}.layers.0.weight", (ch * 2, ch, 2, 2)))
            if i < num_pool_layers - 1:
                ch //= 2
        last = num_pool_layers - 1
        self._head = (self._add(f"up_conv.{last}.1.weight", (out_chans, ch, 1, 1), ch),
                      self._add(f"up_conv.{last}.1.bias", (out_chans,), ch))
        self._up_t = [self._add(k, s, s[0]) for k, s in up_t]

    # ---- parameter registration under dotted names ----------------------------------------------------
    def _add(self, key: str, shape, fan_in: int) -> str:
        node = self
        parts = key.split(".")
        for name in parts[:-1]:
            if name not in node._modules:
                node.add_module(name, _Scope())
            node = node._modules[name]
        bound = 1.0 / math.sqrt(fan_in)
        node.register_parameter(parts[-1], nn.Parameter(torch.empty(shape).uniform_(-bound, bound)))
        self._keys.append(key)
        return key

    def _add_block(self, prefix: str, cin: int, cout: int):
        k0, k1 = _block_keys(prefix)
        return (self._add(k0, (cout, cin, 3, 3), cin * 9), self._add(k1, (cout, cout, 3, 3), cout * 9), cout)

    def _p(self, key: str) -> torch.Tensor:
        return self.get_parameter(key).detach()

    # ---- kernels ------------------------------------------------------------------------------------------
    def _packed(self, key: str):
        """tf32 hi / lo parts of a 3x3 weight in the tensor-core kernel's [channel quad][tap][cout][4] layout,
        packed once per parameter version (inference weights are constant)."""
        p = self.get_parameter(key)
        tag = (p.data_ptr(), p._version, str(p.device))
        hit = self._pack_cache.get(key)
        if hit is None or hit[0] != tag:
            cout, cin = p.shape[0], p.shape[1]
            w_hi = torch.empty((cin // 4) * 9 * cout * 4, dtype=torch.float32, device=p.device)
            w_lo = torch.empty_like(w_hi)
            nat.check(nat.lib().immoco_unet_pack_conv3x3(p.detach().data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), cout,
                                                         cin, _stream()), "unet_pack_conv3x3")
            hit = (tag, w_hi, w_lo)
            self._pack_cache[key] = hit
        return hit[1], hit[2]

    def _conv3x3(self, in0, in1, key, cout):
        n, c0, h, w = in0.shape
        c1 = 0 if in1 is None else in1.shape[1]
        out = torch.empty((n, cout, h, w), dtype=torch.float32, device=in0.device)
        stats = torch.zeros((n, cout, 2), dtype=torch.float64, device=in0.device)
        in1_ptr = 0 if in1 is None else in1.data_ptr()
        if self.tensor_cores and (c0 + c1) % 8 == 0 and c0 % 4 == 0 and cout % 32 == 0:
            # tcgen05 implicit GEMM (csrc/unet_tc.cu); the 2-channel input layer below stays SIMT
            w_hi, w_lo = self._packed(key)
            nat.check(nat.lib().immoco_unet_conv3x3_tc(in0.data_ptr(), c0, in1_ptr, c1, w_hi.data_ptr(), w_lo.data_ptr(),
                                                       out.data_ptr(), stats.data_ptr(), n, cout, h, w, _stream()),
                      "unet_conv3x3_tc")
        else:
            nat.check(nat.lib().immoco_unet_conv3x3(in0.data_ptr(), c0, in1_ptr, c1, self._p(key).data_ptr(),
                                                    out.data_ptr(), stats.data_ptr(), n, cout, h, w, _stream()),
                      "unet_conv3x3")
        return out, stats

    @staticmethod
    def _norm_act(x, stats, pool: bool):
        n, c, h, w = x.shape
        pooled = torch.empty((n, c, h // 2, w // 2), dtype=torch.float32, device=x.device) if pool else None
        nat.check(nat.lib().immoco_unet_instnorm_lrelu(x.data_ptr(), stats.data_ptr(),
                                                       0 if pooled is None else pooled.data_ptr(), n * c, h, w,
                                                       _EPS, _SLOPE, _stream()), "unet_instnorm_lrelu")
        return pooled

    def _block(self, in0, in1, keys, pool: bool):
        k0, k1, cout = keys
        a, st = self._conv3x3(in0, in1, k0, cout)
        self._norm_act(a, st, False)
        b, st = self._conv3x3(a, None, k1, cout)
        pooled = self._norm_act(b, st, pool)
        return b, pooled

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        _need_cuda(image, "Unet.forward")
        if image.dim() != 4 or image.shape[1] != self.in_chans:
            raise ValueError(f"expected an (N, {self.in_chans}, H, W) tensor")
        n, _, h, w = image.shape
        div = 1 << self.num_pool_layers
        if h % div or w % div:
            raise NotImplementedError(f"H and W must be multiples of {div} (the reference reflect-pads odd "
                                      "levels; not on the IM-MoCo path: 320, 640 and 368 divide)")
        for k in self._keys:
            _need_cuda(self.get_parameter(k), "Unet parameters (call .cuda())")
        with torch.no_grad():
            cur = image.detach().float().contiguous()
            skips = []
            for keys in self._down:
                act, cur = self._block(cur, None, keys, True)
                skips.append(act)
            cur, _ = self._block(cur, None, self._mid, False)
            for keys, kt in zip(self._up, self._up_t):
                skip = skips.pop()
                wt = self._p(kt)
                cin, cout = wt.shape[0], wt.shape[1]
                hh, ww = cur.shape[-2:]
                up = torch.empty((n, cout, 2 * hh, 2 * ww), dtype=torch.float32, device=cur.device)
                st = torch.zeros((n, cout, 2), dtype=torch.float64, device=cur.device)
                nat.check(nat.lib().immoco_unet_convt2x2(cur.data_ptr(), wt.data_ptr(), up.data_ptr(), st.data_ptr(),
                                                         n, cin, cout, hh, ww, _stream()), "unet_convt2x2")
                self._norm_act(up, st, False)
                cur, _ = self._block(up, skip, keys, False)       # concat [up, skip] folded into the loads
            wk, bk = self._head
            out = torch.empty((n, self.out_chans, h, w), dtype=torch.float32, device=cur.device)
            nat.check(nat.lib().immoco_unet_conv1x1(cur.data_ptr(), self._p(wk).data_ptr(), self._p(bk).data_ptr(),
                                                    out.data_ptr(), n, cur.shape[1], self.out_chans, h * w, _stream()),
                      "unet_conv1x1")
        return out


def get_unet(in_chans: int, out_chans: int, chans: int, num_pool_layers: int, drop_prob: float, **kwargs):
    """src/models/kld_net.py:4-11."""
    return Unet(in_chans=in_chans, out_chans=out_chans, chans=chans, num_pool_layers=num_pool_layers,
                drop_prob=drop_prob, **kwargs)


def kld_net_input(kspace: torch.Tensor) -> torch.Tensor:
    """(B, 2, H, W) network input from (B, H, W) / (H, W) complex k-space, per slice
    ``k / IFFT(k).abs().std()`` as (re, im) channels (test_immoco.py:47-56)."""
    k = kspace.to(torch.complex64)
    if k.dim() == 2:
        k = k.unsqueeze(0)
    img = IFFT(k).abs()
    k = k / img.flatten(1).std(dim=1).view(-1, 1, 1)
    return torch.view_as_real(k).permute(0, 3, 1, 2).contiguous()


def detect_motion_lines(net: Unet, kspace: torch.Tensor) -> torch.Tensor:
    """(B, W) bool: sigmoid(net) > 0.5, then the per-column vote > 0.2 (test_immoco.py:50-61)."""
    logits = net(kld_net_input(kspace))
    mask = logits.sigmoid() > 0.5                       # (B, 1, H, W)
    return mask[:, 0].sum(1).div(mask.shape[-2]) > 0.2


def movement_masks_from_kspace(net: Unet, kspace: torch.Tensor) -> List[torch.Tensor]:
    """kld-net -> movement groups for every slice of a (B, H, W) stack: the (M_b, H, W) int64 mask lists
    ``IMMoCo`` / ``reconstruct_batch`` take (what ``extract_movement_groups(lines[b], make_list=True)`` returns
    per slice).  The run labelling is one batched pass and ONE host synchronisation for the whole stack (the
    group counts size the outputs)."""
    lines = detect_motion_lines(net, kspace)
    labels = movement_group_labels(lines)                       # (B, W)
    counts = labels.max(dim=1).values.cpu().tolist() if labels.shape[1] else [0] * labels.shape[0]
    h, w = kspace.shape[-2], labels.shape[1]
    out = []
    for b, m in enumerate(counts):
        ids = torch.arange(1, m + 1, device=labels.device, dtype=torch.long).view(m, 1, 1)
        out.append((labels[b].view(1, 1, w).expand(1, h, w) == ids).to(torch.long))
    return out
