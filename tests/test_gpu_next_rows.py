"""GPU: the rows either side of the hot path: synthetic rigid motion (f2) and the
evaluation metrics (f4), against the oracle and the reference-generated golden vectors."""
import numpy as np
import pytest
import torch

import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _built():
    mb.build()
    yield


# ---------------------------------------------------------------------------------------------------
# f2: motion_simulation2D
# ---------------------------------------------------------------------------------------------------
def test_motion_simulation_matches_reference_golden(golden_dir):
    """Same seed as oracle/gen_golden.py (the reference's own motion_simulation2D produced these)."""
    ops = np.load(f"{golden_dir}/ops_small.npz")
    img = torch.from_numpy(ops["sim_image"])
    torch.manual_seed(11)
    k, mask, rot, trans = mb.motion_simulation2D(img.to(DEV), 3)
    want = torch.from_numpy(ops["sim_kspace"])
    assert k.shape == want.shape and k.dtype == torch.complex64
    assert rel_l2(k, want) < 1e-5
    assert torch.equal(mask.cpu(), torch.from_numpy(ops["sim_mask"]))
    assert torch.equal(rot, torch.from_numpy(ops["sim_rot"]))
    assert torch.equal(trans, torch.from_numpy(ops["sim_trans"]))


@pytest.mark.parametrize("h,w,n_mov,seed", [(320, 320, 4, 1000), (640, 368, 5, 7), (64, 46, 2, 3), (320, 320, None, 5)])
def test_motion_simulation_matches_oracle(h, w, n_mov, seed):
    img = orc.make_phantom(h, w, seed)
    torch.manual_seed(seed)
    ko, mo, ro, to = orc.motion_simulation2D(img, n_mov)
    torch.manual_seed(seed)
    k, mask, rot, trans = mb.motion_simulation2D(img.to(DEV), n_mov)
    assert torch.equal(mask.cpu(), mo) and torch.equal(rot, ro) and torch.equal(trans, to)
    assert rel_l2(k, ko) < 1e-5
    # the corrupted lines are exactly the mask's columns; the others are the clean k-space
    clean = orc.FFT(img)
    static = mo[0] == 0
    assert rel_l2(k.cpu()[:, static], clean[:, static]) < 2e-6
    # feeding the result through the group interface gives n_mov groups (windows never merge for n <= 8)
    masks = mb.extract_movement_groups(mb.lines_from_mask(mask), make_list=True, height=h)
    assert masks.shape[0] == (n_mov if n_mov is not None else ro.shape[0]) or n_mov is None


def test_motion_simulation_rejects_cpu_and_bad_rank():
    with pytest.raises(RuntimeError):
        mb.motion_simulation2D(torch.zeros(8, 8, dtype=torch.complex64), 2)
    with pytest.raises(ValueError):
        mb.motion_simulation2D(torch.zeros(2, 8, 8, dtype=torch.complex64, device=DEV), 2)


# ---------------------------------------------------------------------------------------------------
# f4: calmetric2D / crop_metrics
# ---------------------------------------------------------------------------------------------------
def _pair(h, w, seed, noise=0.05):
    g = torch.Generator().manual_seed(seed)
    gt = orc.make_phantom(h, w, seed).abs()
    pred = gt * (1 + noise * torch.randn(h, w, generator=g)) + 0.02 * torch.rand(h, w, generator=g)
    return pred, gt


@pytest.mark.parametrize("h,w", [(160, 160), (320, 184), (33, 47), (640, 640), (300, 520)])
def test_calmetric2d_matches_oracle(h, w):
    pred, gt = _pair(h, w, h + w)
    p, g = orc.normalize01(pred), orc.normalize01(gt)
    want_psnr, want_ssim = orc.psnr01(p, g), orc.ssim01(p, g)
    want_rmse = float(torch.sqrt(torch.mean((p - g) ** 2)))
    psnr, ssim, haar, rmse = mb.calmetric2D(pred[None, None].to(DEV), gt[None, None].to(DEV))
    assert psnr.is_cuda and psnr.dim() == 0
    assert abs(float(psnr) - want_psnr) < 1e-3          # dB
    assert abs(float(ssim) - want_ssim) < 2e-5
    assert abs(float(rmse) - want_rmse) < 1e-6
    # HaarPSI (piq.haarpsi restated, like SSIM parity unpinned): odd sizes exercise the zero-padded 2x2 pooling
    assert abs(float(haar) - orc.haarpsi01(p, g)) < 2e-5


def test_calmetric2d_batch_and_helpers():
    pairs = [_pair(96, 80, s) for s in (1, 2, 3)]
    pred = torch.stack([p for p, _ in pairs])[:, None].to(DEV)
    gt = torch.stack([g for _, g in pairs])[:, None].to(DEV)
    psnr, ssim, haar, rmse = mb.calmetric2D(pred, gt)
    pn, gn = mb.normalize(pred), mb.normalize(gt)               # batch-wise, like evaluate.py:19-29
    ws = [orc.ssim01(pn[i, 0].cpu(), gn[i, 0].cpu()) for i in range(3)]
    assert abs(float(ssim) - float(np.mean(ws))) < 2e-5
    hs = [orc.haarpsi01(pn[i, 0].cpu(), gn[i, 0].cpu()) for i in range(3)]
    assert abs(float(haar) - float(np.mean(hs))) < 2e-5
    same = mb.calmetric2D(gt, gt)
    assert abs(float(same[2]) - 1.0) < 1e-5 and float(same[1]) > 0.99999          # identical images: HaarPSI = SSIM = 1
    tiny = mb.calmetric2D(pred[:, :, :12, :12], gt[:, :, :12, :12])                 # below HaarPSI's 16-pixel kernel
    assert np.isnan(float(tiny[2])) and np.isfinite(float(tiny[0]))
    assert abs(float(psnr) - float(mb.my_psnr(pn, gn, data_range=1.0))) < 1e-3
    assert abs(float(rmse) - float(mb.rmse(pn, gn))) < 1e-6
    with pytest.raises(ValueError):
        mb.calmetric2D(pred[0], gt[0])
    with pytest.raises(RuntimeError):
        mb.calmetric2D(pred.cpu(), gt.cpu())


@pytest.mark.parametrize("h,w", [(320, 320), (640, 368)])
def test_crop_metrics_on_complex_reconstruction(h, w):
    """test_immoco.py:74-85: complex refined image -> magnitude -> central half -> metrics, no copies."""
    img = orc.make_phantom(h, w, 9)
    g = torch.Generator().manual_seed(1)
    refined = img * 16000 * (1 + 0.03 * torch.randn(h, w, generator=g))
    want = orc.crop_metrics(refined.abs(), img.abs())
    psnr, ssim, _, rmse = mb.crop_metrics(refined.to(DEV), img.abs().to(DEV))
    assert abs(float(psnr) - want["psnr"]) < 1e-3
    assert abs(float(ssim) - want["ssim"]) < 2e-5
    assert abs(float(rmse) - want["rmse"]) < 1e-6


# ---------------------------------------------------------------------------------------------------
# f1: kld-net (U-Net) inference -> movement groups
# ---------------------------------------------------------------------------------------------------
from oracle import kld_net_oracle as ko  # noqa: E402


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_unet_matches_reference_golden(golden_dir, tag):
    """Outputs of the reference's own src/models/unet.py (oracle/gen_golden_unet.py) on seeded weights."""
    g = np.load(f"{golden_dir}/unet_small.npz")
    in_c, out_c, chans, pools, n, h, w, seed = (int(v) for v in g[f"{tag}_cfg"])
    net = mb.get_unet(in_c, out_c, chans, pools, 0.0)
    net.load_state_dict(ko.init_unet_state(seed, in_c, out_c, chans, pools))
    net = net.cuda()
    y = net(torch.from_numpy(g[f"{tag}_x"]).to(DEV))
    want = torch.from_numpy(g[f"{tag}_y"])
    assert y.shape == want.shape
    assert rel_l2(y, want) < 2e-5, rel_l2(y, want)


@pytest.mark.parametrize("n,h,w", [(2, 320, 320), (1, 640, 368), (3, 64, 48)])
def test_unet_full_size_matches_oracle(n, h, w):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    state = ko.init_unet_state(3)
    net = mb.get_unet(2, 1, 32, 4, 0.0)
    net.load_state_dict(state)
    net = net.cuda()
    x = torch.randn(n, 2, h, w, generator=torch.Generator().manual_seed(h)).to(DEV) * 2.0
    y = net(x)
    want = ko.unet_forward({k: v.to(DEV) for k, v in state.items()}, x)
    print(f"unet {n}x{h}x{w}: rel-L2 vs torch fp32 {rel_l2(y, want):.2e}")
    assert rel_l2(y, want) < 5e-5


@pytest.mark.parametrize("n,c0,c1,cout,h,w", [(2, 32, 0, 32, 48, 40), (1, 64, 64, 64, 40, 40), (3, 32, 0, 64, 20, 20),
                                              (1, 256, 256, 256, 24, 16), (2, 512, 0, 512, 20, 20), (1, 32, 32, 32, 320, 320),
                                              (1, 8, 0, 96, 17, 9)])
def test_conv3x3_tensor_core_kernel_matches_torch(native_lib, n, c0, c1, cout, h, w):
    """immoco_unet_conv3x3_tc (tcgen05 implicit GEMM, 3xTF32) vs F.conv2d fp32 and vs the fp32 SIMT kernel: raw
    output, instance statistics, channel concat, partial tiles, long channel loops (accumulator flushes)."""
    import ctypes as C
    import torch.nn.functional as F
    from miccai24_immoco_b200 import _native as nat
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(c0 + cout + h)
    x0 = torch.randn(n, c0, h, w, generator=g).to(DEV)
    x1 = torch.randn(n, c1, h, w, generator=g).to(DEV) if c1 else None
    cin = c0 + c1
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (3.0 * cin ** 0.5)).to(DEV)
    s = torch.cuda.current_stream().cuda_stream
    w_hi = torch.empty((cin // 4) * 9 * cout * 4, device=DEV)
    w_lo = torch.empty_like(w_hi)
    nat.check(native_lib.immoco_unet_pack_conv3x3(wt.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), cout, cin, s), "pack")
    out = torch.full((n, cout, h, w), float("nan"), device=DEV)
    stats = torch.zeros(n, cout, 2, dtype=torch.float64, device=DEV)
    nat.check(native_lib.immoco_unet_conv3x3_tc(x0.data_ptr(), c0, 0 if x1 is None else x1.data_ptr(), c1, w_hi.data_ptr(),
                                                w_lo.data_ptr(), out.data_ptr(), stats.data_ptr(), n, cout, h, w, s), "conv_tc")
    xin = x0 if x1 is None else torch.cat([x0, x1], 1)
    want = F.conv2d(xin.double(), wt.double(), padding=1)
    ref32 = F.conv2d(xin, wt, padding=1)
    err, floor = rel_l2(out, want), rel_l2(ref32, want)
    print(f"conv3x3_tc {n}x{cin}->{cout} {h}x{w}: rel-L2 vs fp64 {err:.2e} (cuDNN fp32: {floor:.2e})")
    assert bool(torch.isfinite(out).all()) and err < max(5e-6, 3.0 * floor)
    # sums of the (rounded) fp32 outputs, accumulated in fp64: errors relative to sum |x| (the plain sum cancels)
    assert float(((stats[..., 0] - want.sum((2, 3))).abs() / want.abs().sum((2, 3))).max()) < 1e-6
    assert rel_l2(stats[..., 1], (want * want).sum((2, 3))) < 1e-5
    # the SIMT kernel on the same inputs
    out_s = torch.empty_like(out)
    stats_s = torch.zeros_like(stats)
    nat.check(native_lib.immoco_unet_conv3x3(x0.data_ptr(), c0, 0 if x1 is None else x1.data_ptr(), c1, wt.data_ptr(),
                                             out_s.data_ptr(), stats_s.data_ptr(), n, cout, h, w, s), "conv_simt")
    assert rel_l2(out, out_s) < 5e-6


def test_conv3x3_tensor_core_rejects_unsupported(native_lib):
    from miccai24_immoco_b200 import _native as nat
    z = torch.zeros(16, device=DEV)
    p = z.data_ptr()
    s = torch.cuda.current_stream().cuda_stream
    assert native_lib.immoco_unet_conv3x3_tc(p, 2, 0, 0, p, p, p, p, 1, 32, 8, 8, s) == nat.ERR_UNSUPPORTED     # cin = 2
    assert native_lib.immoco_unet_conv3x3_tc(p, 8, 0, 0, p, p, p, p, 1, 24, 8, 8, s) == nat.ERR_UNSUPPORTED     # cout % 32
    assert native_lib.immoco_unet_conv3x3_tc(p, 8, 0, 0, p, p, p, p, 0, 32, 8, 8, s) == 0                       # empty batch


def test_unet_tensor_core_path_equals_simt_path():
    """The whole network with the 3x3 convolutions on tcgen05 vs on the fp32 SIMT kernels (A/B switch)."""
    state = ko.init_unet_state(5)
    net = mb.get_unet(2, 1, 32, 4, 0.0)
    net.load_state_dict(state)
    net = net.cuda()
    x = torch.randn(2, 2, 96, 80, generator=torch.Generator().manual_seed(1)).to(DEV)
    y_tc = net(x)
    net.tensor_cores = False
    y_simt = net(x)
    assert rel_l2(y_tc, y_simt) < 2e-5


def test_kld_net_to_movement_groups_pipeline():
    """test_immoco.py:47-61 end to end: k-space -> network input -> logits -> mask -> column vote ->
    groups, against the oracle's restatement on the same seeded weights; then the fit accepts them."""
    h = w = 64
    case = orc.make_case(h, w, 2, 4)
    k = case["kspace_motion"]
    state = ko.init_unet_state(11, gain=3.0)
    net = mb.get_unet(2, 1, 32, 4, 0.0)
    net.load_state_dict(state)
    net = net.cuda()
    x_o = ko.kld_net_input(k, orc.IFFT)
    x = mb.kld_net_input(k.to(DEV))
    assert rel_l2(x, x_o) < 1e-5
    logits_o = ko.unet_forward(state, x_o)
    logits = net(x)
    assert rel_l2(logits, logits_o) < 1e-4
    lines_o = ko.motion_lines_from_logits(logits_o)
    lines = mb.detect_motion_lines(net, k.to(DEV))
    # pixels whose logit sits within rounding of 0 may flip; the 20 % column vote must not
    assert lines.shape == (1, w) and torch.equal(lines[0].cpu(), lines_o)
    masks = mb.movement_masks_from_kspace(net, k.to(DEV)[None])
    want = orc.extract_movement_groups(lines_o, make_list=True, height=h)
    assert torch.equal(masks[0].cpu(), want)
    if masks[0].shape[0] > 0:
        im, _ = mb.imcoco_motion_correction(k.to(DEV), masks[0], iters=10)
        assert im.shape == (h, w) and bool(torch.isfinite(torch.view_as_real(im)).all())


def test_unet_rejects_unsupported():
    net = mb.get_unet(2, 1, 8, 4, 0.0).cuda()
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 2, 40, 40, device=DEV))          # 40 is not a multiple of 16
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 32, 32, device=DEV))
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 2, 32, 32))


# ---------------------------------------------------------------------------------------------------
# f3: autofocusing baseline on the same kernels
# ---------------------------------------------------------------------------------------------------
from oracle import autofocus_oracle as ao  # noqa: E402


def _af_case(golden, tag):
    h, w, n_mov, seed, iters = (int(v) for v in golden[f"{tag}_cfg"])
    case = orc.make_case(h, w, n_mov, seed)
    return case, h, w, iters


@pytest.mark.parametrize("tag", ["a", "b"])
def test_autofocusing_forward_matches_reference_golden(golden_dir, tag):
    g = np.load(f"{golden_dir}/autofocus_small.npz")
    case, h, w, _ = _af_case(g, tag)
    k = case["kspace_motion"]
    k = (k / orc.IFFT(k).abs().max()).to(DEV)
    model = mb.Autofocusing(case["masks"].to(DEV))
    p0 = torch.from_numpy(g[f"{tag}_p0"]).to(DEV)
    with torch.no_grad():
        for name, v in zip(("rot_vector", "x_shifts", "y_shifts"), p0):
            model.motion_parameters[name].copy_(v)
    got = model(k)
    assert rel_l2(got, torch.from_numpy(g[f"{tag}_k_fwd"])) < 2e-5


@pytest.mark.parametrize("h,w,n_mov", [(320, 320, 4), (64, 40, 3)])
def test_autofocusing_gradients_match_oracle(h, w, n_mov):
    case = orc.make_case(h, w, n_mov, 8)
    k = case["kspace_motion"]
    k = k / orc.IFFT(k).abs().max()
    masks = case["masks"]
    g = torch.Generator().manual_seed(2)
    p0 = [((torch.rand(n_mov, generator=g) - 0.5) * s) for s in (6.0, 5.0, 5.0)]
    po = [p.clone().requires_grad_(True) for p in p0]
    loss_o = orc.gradient_entropy(orc.IFFT(ao.autofocus_forward(k, masks, *po))) * 1e-4
    loss_o.backward()
    model = mb.Autofocusing(masks.to(DEV))
    with torch.no_grad():
        for name, v in zip(("rot_vector", "x_shifts", "y_shifts"), p0):
            model.motion_parameters[name].copy_(v.to(DEV))
    loss = mb.GradientEntropyLoss()(mb.IFFT(model(k.to(DEV)))) * 1e-4
    loss.backward()
    assert abs(float(loss) - float(loss_o)) < 1e-5 * abs(float(loss_o))
    for name, ref in zip(("rot_vector", "x_shifts", "y_shifts"), po):
        got = model.motion_parameters[name].grad
        assert rel_l2(got, ref.grad) < 2e-3, (name, got, ref.grad)


def test_autofocusing_loop_follows_reference_trajectory(golden_dir):
    """Adam(lr=1) is sign-like in the first steps, so the parameter trajectory is checked to the
    reference's golden one with a tolerance that allows rounding-level gradient differences."""
    g = np.load(f"{golden_dir}/autofocus_small.npz")
    case, h, w, iters = _af_case(g, "a")
    img, k_ref, trace = mb.autofocus_motion_correction(case["kspace_motion"], case["masks"], iters=iters,
                                                       return_trace=True)
    assert img.shape == (h, w) and k_ref.dtype == torch.complex64
    want = g["a_trace"]
    assert np.allclose(trace[:2], want[:2], rtol=1e-4)
    assert np.allclose(trace, want, rtol=2e-2)
    with pytest.raises(RuntimeError):
        mb.Autofocusing(case["masks"])                     # CPU masks: no fallback


# ---------------------------------------------------------------------------------------------------
# the whole of src/test/test_immoco.py on synthetic data: simulate -> detect lines -> groups -> fit -> metrics
# ---------------------------------------------------------------------------------------------------
def test_pipeline_simulate_groups_fit_metrics():
    """Every stage on the CUDA path (no oracle in the loop except the phantom): the corrected image must be
    closer to the ground truth than the motion-corrupted one, like the reference's evaluation reports."""
    h = w = 320                                   # the C2 golden case: seed 1004, n_M = 4 (29.6 dB corrupted)
    gt = orc.make_phantom(h, w, 1004).to(DEV)
    torch.manual_seed(1004)
    k_motion, mask, rot, trans = mb.motion_simulation2D(gt, 4)
    # movement groups from the (ground-truth) line mask through the reference's column vote (test_immoco.py:59-61)
    masks = mb.extract_movement_groups(mb.lines_from_mask(mask), make_list=True, height=h)
    assert masks.shape == (4, h, w)
    corrupted = mb.IFFT(k_motion)
    refined = mb.reconstruct_batch([k_motion], [masks], 200, 1e-2, 1e-2)[0]
    psnr0, ssim0, _, rmse0 = mb.crop_metrics(corrupted, gt.abs())
    psnr1, ssim1, _, rmse1 = mb.crop_metrics(refined, gt.abs())
    print(f"corrupted: {float(psnr0):.2f} dB / {float(ssim0):.4f};  IM-MoCo 200 its: {float(psnr1):.2f} dB / {float(ssim1):.4f}")
    assert float(psnr1) > float(psnr0) + 3.0 and float(ssim1) > float(ssim0)
    assert float(rmse1) < float(rmse0)
