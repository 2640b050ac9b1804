"""CPU: the oracle restatement against the reference-produced golden vectors (tests/golden/*.npz,
written by oracle/gen_golden.py from the reference's own Python files)."""
import os

import numpy as np
import pytest
import torch

from oracle import immoco_oracle as orc


@pytest.fixture(scope="module")
def ops(golden_dir):
    return np.load(os.path.join(golden_dir, "ops_small.npz"))


@pytest.mark.parametrize("hw", ["32x32", "48x20", "64x46"])
def test_fft_matches_reference(ops, hw):
    x = torch.from_numpy(ops[f"fft_in_{hw}"])
    ref = torch.from_numpy(ops[f"fft_out_{hw}"])
    assert torch.equal(orc.FFT(x), ref)
    # IFFT(FFT(x)) == x and the transform is un-normalised: ||F x||^2 = HW ||x||^2   
    assert torch.allclose(orc.IFFT(ref), x, atol=1e-5)
    h, w = x.shape
    assert abs(float((ref.abs() ** 2).sum() / (x.abs() ** 2).sum()) / (h * w) - 1) < 1e-5


def test_gradient_entropy_matches_reference(ops):
    x = torch.from_numpy(ops["ge_in"]).clone().requires_grad_(True)
    val = orc.gradient_entropy(x)
    val.backward()
    assert torch.equal(val.detach(), torch.from_numpy(ops["ge_val"]))
    assert torch.equal(x.grad, torch.from_numpy(ops["ge_grad"]))


@pytest.mark.parametrize("name", ["empty", "all", "runs", "edges", "single_last"])
@pytest.mark.parametrize("make_list", [0, 1])
def test_movement_groups_match_reference(ops, name, make_list):
    lines = torch.from_numpy(ops[f"groups_{name}_in"])
    want = torch.from_numpy(ops[f"groups_{name}_{make_list}"])
    got = orc.extract_movement_groups(lines, make_list=bool(make_list))
    assert got.shape == want.shape and torch.equal(got, want)


def test_motion_simulation_matches_reference(ops):
    img = torch.from_numpy(ops["sim_image"])
    torch.manual_seed(11)
    k, mask, rot, trans = orc.motion_simulation2D(img.clone(), 3)
    assert torch.equal(k, torch.from_numpy(ops["sim_kspace"]))
    assert torch.equal(mask, torch.from_numpy(ops["sim_mask"]))
    assert torch.equal(rot, torch.from_numpy(ops["sim_rot"]))
    assert torch.equal(trans, torch.from_numpy(ops["sim_trans"]))


def _case_params(seed):
    lv2 = orc.make_grid_levels(2, orc.ENCODING_CONFIG)
    lv3 = orc.make_grid_levels(3, orc.ENCODING_CONFIG)
    return (orc.init_params(lv2, orc.IMAGE_NETWORK_CONFIG, 100 + seed),
            orc.init_params(lv3, orc.MOTION_NETWORK_CONFIG, 200 + seed))


def test_forward_and_loop_match_reference_small(golden_dir):
    g = np.load(os.path.join(golden_dir, "loop_s32_m1.npz"))
    h, n_mov, seed, iters = int(g["h"]), int(g["n_mov"]), int(g["seed"]), int(g["iters"])
    case = orc.make_case(h, h, n_mov, seed)
    assert np.array_equal(case["masks"][:, 0, :].numpy().astype(np.uint8), g["masks_lines"])
    assert np.array_equal(case["kspace_motion"].numpy(), g["kspace_motion"])
    p_img, p_mot = _case_params(seed)
    model = orc.IMMoCo(case["masks"], image_params=p_img, motion_params=p_mot)
    with torch.no_grad():
        k0, im0 = model()
    assert np.allclose(k0.numpy(), g["k_fwd0"], rtol=1e-6, atol=1e-6 * np.abs(g["k_fwd0"]).max())
    assert np.allclose(im0.numpy(), g["image0"], rtol=1e-6, atol=1e-9)
    im, k, trace = orc.imcoco_motion_correction(case["kspace_motion"], case["masks"], iters=iters,
                                                image_params=p_img, motion_params=p_mot, return_trace=True)
    want = g["loss_trace"]
    rel = np.abs(np.asarray(trace) - want) / np.abs(want)
    # bit-identical in the build container; another CPU may round differently -> drift band
    band = np.maximum.accumulate(np.abs(g["loss_trace_perturbed"] - want) / np.abs(want))
    assert np.all(rel <= np.maximum(1e-5, 20 * band)), rel.max()


def test_lambda_schedule_quirk():
    # halved on every iteration that is NOT a multiple of iters//10 after the midpoint
    lams = orc.lambda_schedule(200, 1e-2)
    halvings = sum(1 for a, b in zip(lams[:-1], lams[1:]) if b != a)
    assert halvings == 94 and lams[101] == 1e-2 and lams[102] == 5e-3   # j=101 is the first halving
    with pytest.raises(ZeroDivisionError):
        orc.lambda_schedule(9, 1e-2)
    d = orc.lambda_schedule(200, 1e-2, variant="downstream")
    assert d[90] == 1e-2 and d[91] == 5e-3


def test_level_tables_match_survey():
    lv2 = orc.make_grid_levels(2, orc.ENCODING_CONFIG)
    lv3 = orc.make_grid_levels(3, orc.ENCODING_CONFIG)
    assert lv2.offsets[:8] == (0, 256, 1280, 5376, 21760, 87296, 349440, 873728) and lv2.offsets[-1] == 5592320
    assert lv3.offsets[:5] == (0, 4096, 36864, 299008, 823296) and lv3.offsets[-1] == 7114752
    assert lv2.hashed == (False,) * 6 + (True,) * 10
    assert lv3.hashed == (False,) * 3 + (True,) * 13
    assert lv2.n_table_params + orc.mlp_param_count(32, orc.IMAGE_NETWORK_CONFIG) == 11196928
    assert lv3.n_table_params + orc.mlp_param_count(32, orc.MOTION_NETWORK_CONFIG) == 14232576


def test_metrics_sane():
    a = torch.rand(64, 64)
    m = orc.crop_metrics(a, a)
    assert m["ssim"] > 0.999999 and m["rmse"] == 0.0
    assert abs(m["haarpsi"] - 1.0) < 1e-6


def _haarpsi_loops(x, y, c=30.0, alpha=4.2):
    """HaarPSI written out pixel by pixel in numpy (an independent statement of the published algorithm that
    piq.haarpsi implements; piq itself is absent here, the seam stays 'parity unpinned')."""
    def pool(a):
        d = max(a.shape[0] % 2, a.shape[1] % 2)
        a = np.pad(a * 255.0, ((0, d), (0, d)))
        a = a[:a.shape[0] // 2 * 2, :a.shape[1] // 2 * 2]            # avg_pool2d floors: a last odd row / column drops
        return 0.25 * (a[0::2, 0::2] + a[1::2, 0::2] + a[0::2, 1::2] + a[1::2, 1::2])

    def coeff(a, k, transpose):
        h, w = a.shape
        out = np.zeros((h, w))
        off = k // 2 - 1
        for i in range(h):
            for j in range(w):
                s = 0.0
                for u in range(k):
                    for v in range(k):
                        yy, xx = i + u - off, j + v - off
                        if 0 <= yy < h and 0 <= xx < w:
                            neg = (v >= k // 2) if transpose else (u >= k // 2)
                            s += -a[yy, xx] if neg else a[yy, xx]
                out[i, j] = s / k
        return out

    px, py = pool(x), pool(y)
    num = den = 0.0
    for o in (False, True):
        cx = [np.abs(coeff(px, k, o)) for k in (2, 4, 8)]
        cy = [np.abs(coeff(py, k, o)) for k in (2, 4, 8)]
        sim = sum((2 * a * b + c) / (a * a + b * b + c) for a, b in zip(cx[:2], cy[:2])) / 2
        wgt = np.maximum(cx[2], cy[2])
        num += (wgt / (1 + np.exp(-alpha * sim))).sum()
        den += wgt.sum()
    eps = float(np.finfo(np.float32).eps)
    s = (num + eps) / (den + eps)
    return (np.log(s / (1 - s)) / alpha) ** 2


@pytest.mark.parametrize("hw", [(16, 16), (21, 18), (24, 33)])
def test_haarpsi_oracle_equals_pixel_loops(hw):
    g = torch.Generator().manual_seed(hw[0] * 100 + hw[1])
    x = torch.rand(hw, generator=g, dtype=torch.float64)
    y = (x + 0.2 * torch.rand(hw, generator=g, dtype=torch.float64)).clamp(0, 1)
    want = _haarpsi_loops(x.numpy(), y.numpy())
    got = orc.haarpsi01(x, y)
    assert 0.0 < got < 1.0 and abs(got - want) < 1e-10
    assert orc.haarpsi01(x, (x + 0.6 * torch.rand(hw, generator=g, dtype=torch.float64)).clamp(0, 1)) < got


def test_kld_net_oracle_against_reference_unet_golden(golden_dir):
    """oracle/kld_net_oracle.py vs outputs of the reference's own src/models/unet.py
    (oracle/gen_golden_unet.py checked bit-equality in the build container)."""
    from oracle import kld_net_oracle as ko
    g = np.load(os.path.join(golden_dir, "unet_small.npz"))
    for tag in ("a", "c"):
        in_c, out_c, chans, pools, n, h, w, seed = (int(v) for v in g[f"{tag}_cfg"])
        state = ko.init_unet_state(seed, in_c, out_c, chans, pools)
        y = ko.unet_forward(state, torch.from_numpy(g[f"{tag}_x"]), pools)
        want = torch.from_numpy(g[f"{tag}_y"])
        assert y.shape == want.shape
        assert float((y - want).abs().max()) <= 1e-5 * float(want.abs().max())
    spec = ko.unet_state_spec(2, 1, 32, 4)
    assert len(spec) == 24 and sum(int(np.prod(s)) for _, s in spec) == 7756385


def test_get_unet_has_fastmri_state_dict_layout():
    """kLDNet.pth (fastmri.models.Unet(2,1,32,4)) must load into the CUDA-path module unchanged."""
    import miccai24_immoco_b200 as mb
    from oracle import kld_net_oracle as ko
    net = mb.get_unet(in_chans=2, out_chans=1, chans=32, num_pool_layers=4, drop_prob=0.0)
    spec = ko.unet_state_spec(2, 1, 32, 4)
    sd = net.state_dict()
    assert list(sd.keys()) == [k for k, _ in spec]
    assert all(tuple(sd[k].shape) == tuple(s) for k, s in spec)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 2, 32, 32))          # CPU tensors: no fallback


def test_autofocus_oracle_against_reference_golden(golden_dir):
    """oracle/autofocus_oracle.py vs the reference's own src/models/autofocusing.py outputs
    (oracle/gen_golden_autofocus.py checked bit-equality in the build container)."""
    from oracle import autofocus_oracle as ao
    g = np.load(os.path.join(golden_dir, "autofocus_small.npz"))
    h, w, n_mov, seed, iters = (int(v) for v in g["a_cfg"])
    case = orc.make_case(h, w, n_mov, seed)
    k = case["kspace_motion"]
    k = k / orc.IFFT(k).abs().max()
    p0 = [torch.from_numpy(v) for v in g["a_p0"]]
    got = ao.autofocus_forward(k, case["masks"], *p0)
    want = torch.from_numpy(g["a_k_fwd"])
    assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())
    _, _, trace, params = ao.autofocus_loop(case["kspace_motion"], case["masks"], iters)
    assert np.allclose(trace, g["a_trace"], rtol=1e-4)
    assert np.allclose(torch.stack(params).numpy(), g["a_params"], atol=1e-3)
