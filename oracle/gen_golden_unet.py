"""ORACLE tooling -- pins oracle/kld_net_oracle.py against the reference's own ``src/models/unet.py``
(imported UNCHANGED, ``batchnorm=nn.InstanceNorm2d`` = fastmri's Unet structure) and writes
tests/golden/unet_small.npz.  Run in the build container only (needs /root/reference):

    python oracle/gen_golden_unet.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from oracle import kld_net_oracle as ko  # noqa: E402


def main():
    from models.unet import Unet            # the reference's file
    torch.set_num_threads(4)
    out = {}
    for tag, (in_c, out_c, chans, pools, n, h, w, seed) in {
        "a": (2, 1, 8, 4, 2, 32, 48, 5),      # kld-net topology, narrow
        "b": (2, 1, 32, 4, 1, 48, 32, 6),     # kld-net itself (7,756,385 parameters)
        "c": (1, 2, 4, 2, 1, 20, 12, 7),      # other channel counts / depth
    }.items():
        state = ko.init_unet_state(seed, in_c, out_c, chans, pools)
        ref = Unet(in_c, out_c, chans, pools, 0.0, batchnorm=nn.InstanceNorm2d).eval()
        assert list(ref.state_dict().keys()) == list(state.keys()), "state-dict keys / order differ"
        ref.load_state_dict(state)
        n_par = sum(p.numel() for p in ref.parameters())
        g = torch.Generator().manual_seed(seed + 100)
        x = torch.randn(n, in_c, h, w, generator=g) * 3.0
        with torch.no_grad():
            want = ref(x)
            got = ko.unet_forward(state, x, pools)
        assert torch.equal(got, want), f"restatement differs from the reference module ({tag})"
        print(f"case {tag}: {n_par} parameters, output {tuple(want.shape)}, restatement bit-identical")
        out[f"{tag}_cfg"] = np.asarray([in_c, out_c, chans, pools, n, h, w, seed])
        out[f"{tag}_x"] = x.numpy()
        out[f"{tag}_y"] = want.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "unet_small.npz"), **out)
    print("wrote tests/golden/unet_small.npz")


if __name__ == "__main__":
    main()
