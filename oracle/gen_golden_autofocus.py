"""ORACLE tooling -- pins oracle/autofocus_oracle.py against the reference's own
``src/models/autofocusing.py`` (imported UNCHANGED; h5py stand-in for utils/data_utils.py's import) and
writes tests/golden/autofocus_small.npz.  Build container only (needs /root/reference):

    python oracle/gen_golden_autofocus.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, "/root/reference/src")

from oracle import autofocus_oracle as ao  # noqa: E402
from oracle import immoco_oracle as orc  # noqa: E402


def main():
    from models.autofocusing import Autofocusing          # the reference's file
    from utils.losses import GradientEntropyLoss
    from utils.data_utils import IFFT
    torch.set_num_threads(4)
    out = {}
    for tag, (h, w, n_mov, seed, iters) in {"a": (48, 48, 2, 3, 6), "b": (64, 40, 3, 5, 4)}.items():
        case = orc.make_case(h, w, n_mov, seed)
        k = case["kspace_motion"]
        k = k / IFFT(k).abs().max()
        masks = case["masks"]
        # forward at non-trivial parameters
        g = torch.Generator().manual_seed(seed)
        p0 = [(torch.rand(masks.shape[0], generator=g) - 0.5) * s for s in (8.0, 6.0, 6.0)]
        ref = Autofocusing(masks)
        with torch.no_grad():
            ref.motion_parameters["rot_vector"].copy_(p0[0])
            ref.motion_parameters["x_shifts"].copy_(p0[1])
            ref.motion_parameters["y_shifts"].copy_(p0[2])
            want = ref(k)
            got = ao.autofocus_forward(k, masks, *p0)
        assert torch.equal(got, want), f"forward restatement differs ({tag})"
        # optimisation trajectory from zero parameters (test_autofocusing.py:58-72)
        ref = Autofocusing(masks)
        opt = torch.optim.Adam(ref.parameters(), lr=1.0)
        trace = []
        for _ in range(iters):
            opt.zero_grad()
            k_ref = ref(k)
            loss = GradientEntropyLoss()(IFFT(k_ref)) * 1e-4
            loss.backward()
            opt.step()
            trace.append(float(loss))
        _, k_o, trace_o, params_o = ao.autofocus_loop(case["kspace_motion"], masks, iters)
        assert trace == trace_o and torch.equal(k_o, k_ref.detach()), f"loop restatement differs ({tag})"
        for name, po in zip(("rot_vector", "x_shifts", "y_shifts"), params_o):
            assert torch.equal(po, ref.motion_parameters[name].detach())
        print(f"case {tag}: forward + {iters}-step Adam trajectory bit-identical; loss {trace[0]:.6f} -> {trace[-1]:.6f}")
        out[f"{tag}_cfg"] = np.asarray([h, w, n_mov, seed, iters])
        out[f"{tag}_p0"] = torch.stack(p0).numpy()
        out[f"{tag}_k_fwd"] = want.numpy()
        out[f"{tag}_trace"] = np.asarray(trace)
        out[f"{tag}_params"] = torch.stack(params_o).numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "autofocus_small.npz"), **out)
    print("wrote tests/golden/autofocus_small.npz")


if __name__ == "__main__":
    main()
