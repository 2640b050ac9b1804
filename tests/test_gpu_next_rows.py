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
    assert np.isnan(float(haar))


def test_calmetric2d_batch_and_helpers():
    pairs = [_pair(96, 80, s) for s in (1, 2, 3)]
    pred = torch.stack([p for p, _ in pairs])[:, None].to(DEV)
    gt = torch.stack([g for _, g in pairs])[:, None].to(DEV)
    psnr, ssim, _, rmse = mb.calmetric2D(pred, gt)
    pn, gn = mb.normalize(pred), mb.normalize(gt)               # batch-wise, like evaluate.py:19-29
    ws = [orc.ssim01(pn[i, 0].cpu(), gn[i, 0].cpu()) for i in range(3)]
    assert abs(float(ssim) - float(np.mean(ws))) < 2e-5
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
