"""ORACLE -- test infrastructure, NOT product code.

Plain torch (fp32) restatement of the kld-net inference path that feeds the IM-MoCo fit:
``get_unet(in_chans=2, out_chans=1, chans=32, num_pool_layers=4, drop_prob=0.0)``
(src/models/kld_net.py:4-11 -> fastmri.models.Unet 0.3.0, absent here) followed by the mask /
column-vote pre- and post-processing of src/test/test_immoco.py:47-61.

Pinning status: fastmri is absent, but the reference vendors the same network as
``src/models/unet.py`` (Unet :17-118, ConvBlock :121-163, TransposeConvBlock :166-187); with
``batchnorm=nn.InstanceNorm2d`` it has fastmri's layer structure and state-dict keys.
``oracle/gen_golden_unet.py`` imports that file UNCHANGED, loads the seeded state below into it and
checks this functional restatement bit-for-bit, then stores ``tests/golden/unet_small.npz``.  PINNED
against the reference's own module; "fastmri == vendored copy" itself rests on SURVEY's probe.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict

import torch
import torch.nn.functional as F


def unet_state_spec(in_chans: int, out_chans: int, chans: int, num_pool_layers: int):
    """[(state-dict key, shape)] in fastmri.models.Unet / src/models/unet.py order (:44-70)."""
    spec = []

    def block(prefix, cin, cout):
        spec.append((f"{prefix}.layers.0.weight", (cout, cin, 3, 3)))
        spec.append((f"{prefix}.layers.4.weight", (cout, cout, 3, 3)))

    block("down_sample_layers.0", in_chans, chans)
    ch = chans
    for i in range(1, num_pool_layers):
        block(f"down_sample_layers.{i}", ch, ch * 2)
        ch *= 2
    block("conv", ch, ch * 2)
    ups, tcs = [], []
    for i in range(num_pool_layers - 1):
        tcs.append((f"up_transpose_conv.{i}.layers.0.weight", (ch * 2, ch, 2, 2)))
        ups.append((f"up_conv.{i}", ch * 2, ch))
        ch //= 2
    last = num_pool_layers - 1
    tcs.append((f"up_transpose_conv.{last}.layers.0.weight", (ch * 2, ch, 2, 2)))
    # nn.Module registration order: up_conv (all) is created before up_transpose_conv (unet.py:57-58)
    for prefix, cin, cout in ups:
        block(prefix, cin, cout)
    block(f"up_conv.{last}.0", ch * 2, ch)
    spec.append((f"up_conv.{last}.1.weight", (out_chans, ch, 1, 1)))
    spec.append((f"up_conv.{last}.1.bias", (out_chans,)))
    spec.extend(tcs)
    return spec


def init_unet_state(seed: int, in_chans: int = 2, out_chans: int = 1, chans: int = 32,
                    num_pool_layers: int = 4, gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded stand-in weights (the trained kLDNet.pth cannot be downloaded here): every tensor
    U(-b, b), b = gain / sqrt(fan_in), drawn in state-dict order from one CPU generator."""
    g = torch.Generator().manual_seed(seed)
    state = OrderedDict()
    for key, shape in unet_state_spec(in_chans, out_chans, chans, num_pool_layers):
        if key.endswith("bias"):
            fan_in = chans
        elif "up_transpose_conv" in key:
            fan_in = shape[0] * 4 // 4            # one input pixel feeds each output: Cin taps
        else:
            fan_in = shape[1] * shape[2] * shape[3]
        b = gain / math.sqrt(fan_in)
        state[key] = (torch.rand(shape, generator=g) * 2 - 1) * b
    return state


def _conv_block(x, w0, w1):
    # ConvBlock (unet.py:121-163): conv3x3(no bias) -> InstanceNorm2d -> LeakyReLU(0.2) -> Dropout2d, twice
    for w in (w0, w1):
        x = F.conv2d(x, w, padding=1)
        x = F.instance_norm(x, eps=1e-5)
        x = F.leaky_relu(x, 0.2)
    return x


def unet_forward(state: Dict[str, torch.Tensor], image: torch.Tensor, num_pool_layers: int = 4) -> torch.Tensor:
    """Unet.forward (unet.py:72-118), eval mode (dropout is the identity, drop_prob = 0 in kld-net)."""
    stack = []
    out = image
    for i in range(num_pool_layers):
        p = f"down_sample_layers.{i}"
        out = _conv_block(out, state[f"{p}.layers.0.weight"], state[f"{p}.layers.4.weight"])
        stack.append(out)
        out = F.avg_pool2d(out, kernel_size=2, stride=2, padding=0)
    out = _conv_block(out, state["conv.layers.0.weight"], state["conv.layers.4.weight"])
    for i in range(num_pool_layers):
        skip = stack.pop()
        # TransposeConvBlock (unet.py:166-187)
        out = F.conv_transpose2d(out, state[f"up_transpose_conv.{i}.layers.0.weight"], stride=2)
        out = F.leaky_relu(F.instance_norm(out, eps=1e-5), 0.2)
        pad = [0, 0, 0, 0]
        if out.shape[-1] != skip.shape[-1]:
            pad[1] = 1
        if out.shape[-2] != skip.shape[-2]:
            pad[3] = 1
        if sum(pad) != 0:
            out = F.pad(out, pad, "reflect")
        out = torch.cat([out, skip], dim=1)
        p = f"up_conv.{i}" if i < num_pool_layers - 1 else f"up_conv.{i}.0"
        out = _conv_block(out, state[f"{p}.layers.0.weight"], state[f"{p}.layers.4.weight"])
    last = num_pool_layers - 1
    return F.conv2d(out, state[f"up_conv.{last}.1.weight"], state[f"up_conv.{last}.1.bias"])


def kld_net_input(kspace: torch.Tensor, FFT_inverse) -> torch.Tensor:
    """(1, 2, H, W) network input of test_immoco.py:47-56: k / IFFT(k).abs().std(), (re, im) channels."""
    k = kspace.reshape(1, 1, *kspace.shape[-2:])
    img = FFT_inverse(k).abs()
    return torch.view_as_real(k / img.std()).squeeze(1).permute(0, 3, 1, 2).contiguous()


def motion_lines_from_logits(logits: torch.Tensor) -> torch.Tensor:
    """test_immoco.py:50-61: sigmoid > 0.5, column vote > 0.2 -> (W,) bool."""
    mask = logits.sigmoid() > 0.5
    m = mask.squeeze()
    return m.sum(0).div(m.shape[0]) > 0.2
