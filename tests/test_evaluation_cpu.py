"""CPU: the data format and the evaluation loop either side of the fit (miccai24_immoco_b200/evaluation.py, mirroring
src/utils/prepareData.py:144-216 and src/test/test_immoco.py:27-130).  The compute steps are injected (the package's own
are CUDA-only and tested in tests/test_gpu_next_rows.py); what is checked here is the file layout, the control flow, the
chunking and the statistics."""
import os

import numpy as np
import pytest
import torch

import miccai24_immoco_b200 as mb
from miccai24_immoco_b200 import evaluation as ev
from oracle import immoco_oracle as orc


def _fake_simulate(img, n_mov):
    """Stand-in with motion_simulation2D's return contract (k-space, int64 line mask, rotations, translations)."""
    k = orc.FFT(img)
    mask = torch.zeros(img.shape, dtype=torch.long)
    mask[:, : n_mov] = 1
    return k, mask, torch.arange(n_mov, dtype=torch.float32), torch.ones((n_mov, 2))


def _fake_metrics(pred, gt):
    d = float((pred - gt).abs().mean())
    return (torch.tensor(40.0 - d), torch.tensor(0.9), torch.tensor(0.8), torch.tensor(d))     # psnr, ssim, haarpsi, rmse


def _images(n, h=16, w=12, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.complex(torch.rand(h, w, generator=g), torch.rand(h, w, generator=g)) for _ in range(n)]


def test_make_save_load_test_set_round_trip(tmp_path):
    np.random.seed(3)
    data = ev.make_test_set(_images(5), movements=(6, 10), simulate=_fake_simulate, metrics=_fake_metrics, ifft=orc.IFFT,
                            device="cpu")
    assert tuple(data) == ev.TEST_SET_KEYS                                   # the keys prepareData.py:199-206 writes
    assert data["kspace_motion"].shape == (5, 16, 12) and data["kspace_motion"].dtype == torch.complex64
    assert data["mask"].dtype == torch.int64 and data["image_rss"].shape == (5, 16, 12)
    counts = [int(r.numel()) for r in data["rotation"]]
    assert all(6 <= c < 10 for c in counts) and [t.shape for t in data["translation"]] == [(c, 2) for c in counts]
    assert all(tuple(m) == ev.METRIC_KEYS for m in data["metrics"])          # key order of test_immoco.py:87-94
    # IFFT(FFT(img)) == img: the corrupted-image metrics of the stand-in are those of the ground truth itself
    assert all(abs(m["rmse"]) < 1e-6 and abs(m["psnr"] - 40.0) < 1e-5 for m in data["metrics"])
    path = os.path.join(tmp_path, "_test_data_light.pth")
    mb.save_test_set(data, path)
    back = mb.load_test_set(path)
    assert torch.equal(back["kspace_motion"], data["kspace_motion"]) and torch.equal(back["mask"], data["mask"])
    assert back["metrics"] == data["metrics"]
    # a file written the way the reference writes it (plain torch.save of the dict) loads as well
    torch.save(dict(data), path)
    assert torch.equal(mb.load_test_set(path)["image_rss"], data["image_rss"])


def test_test_set_layout_is_checked():
    data = ev.make_test_set(_images(2), simulate=_fake_simulate, metrics=_fake_metrics, ifft=orc.IFFT, device="cpu")
    for broken in ({k: v for k, v in data.items() if k != "mask"},
                   {**data, "mask": data["mask"].float()},
                   {**data, "kspace_motion": data["kspace_motion"].real},
                   {**data, "rotation": data["rotation"][:1]},
                   {**data, "metrics": [{"psnr": 1.0}] * 2},
                   {**data, "image_rss": data["image_rss"][:, :8]}):
        with pytest.raises((KeyError, ValueError)):
            mb.validate_test_set(broken)
    with pytest.raises(ValueError):
        ev.make_test_set([], simulate=_fake_simulate, metrics=_fake_metrics, ifft=orc.IFFT, device="cpu")
    with pytest.raises(ValueError):
        ev.make_test_set(_images(1) + _images(1, h=8), simulate=_fake_simulate, metrics=_fake_metrics, ifft=orc.IFFT,
                         device="cpu")


def test_run_test_immoco_control_flow_and_chunking(tmp_path):
    data = ev.make_test_set(_images(7, seed=1), simulate=_fake_simulate, metrics=_fake_metrics, ifft=orc.IFFT, device="cpu")
    calls = {"detect": [], "fit": []}

    def detect(k):                      # one call per chunk; one (M, H, W) int64 mask stack per slice
        calls["detect"].append(k.shape[0])
        return [torch.ones((1 + i % 3, k.shape[1], k.shape[2]), dtype=torch.long) for i in range(k.shape[0])]

    def fit(ks, masks):                 # "corrected image" = the inverse transform of what came in
        calls["fit"].append((len(ks), [int(m.shape[0]) for m in masks]))
        return [orc.IFFT(k) for k in ks]

    out, images = mb.run_test_immoco(data, None, device="cpu", chunk=3, detect=detect, fit=fit, metrics=_fake_metrics,
                                     return_images=True)
    assert calls["detect"] == [3, 3, 1] and [c[0] for c in calls["fit"]] == [3, 3, 1]
    assert calls["fit"][0][1] == [1, 2, 3]
    assert len(out) == 7 and all(tuple(m) == ev.METRIC_KEYS for m in out) and len(images) == 7
    assert all(m["rmse"] < 1e-6 for m in out)               # slice order kept: every result met ITS ground truth
    # from a file, and the whole-script form: one file per scenario -> list of lists + table
    p_light, p_heavy = os.path.join(tmp_path, "light.pth"), os.path.join(tmp_path, "heavy.pth")
    mb.save_test_set(data, p_light)
    mb.save_test_set(ev.make_test_set(_images(4, seed=2), movements=ev.SCENARIO_MOVEMENTS["heavy"], simulate=_fake_simulate,
                                      metrics=_fake_metrics, ifft=orc.IFFT, device="cpu"), p_heavy)
    res = os.path.join(tmp_path, "immoco_metrics.pth")
    metrics_all, table = mb.evaluate_scenarios({"light": p_light, "heavy": p_heavy}, None, out_path=res, device="cpu",
                                               detect=detect, fit=fit, metrics=_fake_metrics)
    assert [len(m) for m in metrics_all] == [7, 4] and set(table) == {"light", "heavy"}
    stored = torch.load(res, weights_only=False)             # the file test_immoco.py:110 writes and :113 re-reads
    assert stored == metrics_all


def test_summarize_metrics_is_mean_and_unbiased_std():
    rows = [{"ssim": 0.9, "psnr": 30.0, "haar_psi": 0.7, "rmse": 0.1}, {"ssim": 0.8, "psnr": 34.0, "haar_psi": 0.9, "rmse": 0.3},
            {"ssim": 0.7, "psnr": 32.0, "haar_psi": 0.8, "rmse": 0.2}]
    table = mb.summarize_metrics([rows, rows[:1]])
    assert abs(table["light"]["psnr"][0] - 32.0) < 1e-5 and abs(table["light"]["psnr"][1] - 2.0) < 1e-5
    assert abs(table["light"]["ssim"][1] - 0.1) < 1e-6
    assert table["heavy"]["rmse"][0] == pytest.approx(0.1) and np.isnan(table["heavy"]["rmse"][1])
