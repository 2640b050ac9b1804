"""GPU: the bit-reproducible fit (immoco_fit::deterministic, ``FitEngine(deterministic=True)``).

Two runs of the same call must agree bit for bit -- image, forward k-space, loss trace, every parameter
and both Adam moments -- whatever the two-stream schedule does; the fused gather + Adam kernel must equal
gather-then-Adam bit for bit; the batch driver must return exactly what the per-slice call returns; and
the reproducible path must agree with the float-atomic path to rounding on the first iterations."""
import numpy as np
import pytest
import torch

import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc
from tests.gpu_util import case_params, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _setup():
    mb.build()
    yield


def _engine_run(h, w, n_mov, seed, iters, deterministic, fuse_adam=None, chunks=None, kick=True):
    case = orc.make_case(h, w, max(n_mov, 1), seed)
    masks = case["masks"][:n_mov].to(DEV)
    p_img, p_mot = case_params(seed, DEV)
    if kick and n_mov > 0:       # a displacement field of ~0.1 from the first iteration on
        p_mot = p_mot.clone()
        p_mot[2048:3072] *= 10.0
        p_mot[3072:] *= 300.0
    model = mb.IMMoCo(masks)
    eng = mb.FitEngine(model, iters, deterministic=deterministic, fuse_adam=fuse_adam)
    k = case["kspace_motion"]
    eng.set_kspace((k / k.abs().max() * 16000).to(DEV))
    eng.reset(p_img, p_mot)
    lam = mb.lambda_schedule(max(iters, 10), 1e-2)[:iters]
    for a, b in (chunks or [(0, iters)]):
        eng.run(lam, 1e-2, a, b)
    torch.cuda.synchronize()
    return {"image": eng.image.clone(), "k": eng.k_out.clone(), "params": eng.params.clone(),
            "m": eng.state[1].clone(), "v": eng.state[2].clone(), "trace": eng.loss_trace(lam).copy(),
            "loss": eng.loss[:iters].clone(),
            # reference layout (the two paths store the tables in different row orders)
            "params_ref": torch.cat([eng.motion_params(), eng.image_params()])}


def _assert_identical(a, b, what):
    for key in ("image", "k", "params", "m", "v", "loss"):
        assert torch.equal(a[key], b[key]), f"{what}: {key} differs"
    assert np.array_equal(a["trace"], b["trace"]), f"{what}: loss trace differs"


@pytest.mark.parametrize("h,w,n_mov,iters", [(320, 320, 4, 40), (640, 368, 5, 12), (64, 48, 2, 30), (48, 40, 1, 30),
                                             (32, 32, 0, 20)])
def test_deterministic_fit_is_bit_reproducible(h, w, n_mov, iters):
    runs = [_engine_run(h, w, n_mov, 1000, iters, True) for _ in range(3)]
    _assert_identical(runs[0], runs[1], "run 2")
    _assert_identical(runs[0], runs[2], "run 3")
    assert np.isfinite(runs[0]["trace"]).all() and runs[0]["trace"][-1] < runs[0]["trace"][0]


def test_deterministic_fit_does_not_depend_on_the_schedule():
    """Serial (instrumented), chunked and two-stream issue orders, PDL on / off: same bits."""
    lib = mb.lib()
    h, w, n_mov, iters = 64, 48, 2, 24
    ref = _engine_run(h, w, n_mov, 7, iters, True)
    _assert_identical(ref, _engine_run(h, w, n_mov, 7, iters, True, chunks=[(0, 5), (5, 6), (6, iters)]), "chunked")
    try:
        lib.immoco_set_branch_overlap(0)
        _assert_identical(ref, _engine_run(h, w, n_mov, 7, iters, True), "single stream")
        lib.immoco_set_branch_overlap(1)
        lib.immoco_set_pdl(0)
        _assert_identical(ref, _engine_run(h, w, n_mov, 7, iters, True), "no PDL")
    finally:
        lib.immoco_set_branch_overlap(1)
        lib.immoco_set_pdl(1)


@pytest.mark.parametrize("h,w,n_mov", [(320, 320, 4), (64, 48, 2), (32, 32, 0)])
def test_fused_gather_adam_equals_gather_then_adam(h, w, n_mov):
    a = _engine_run(h, w, n_mov, 1001, 16, True, fuse_adam=True)
    b = _engine_run(h, w, n_mov, 1001, 16, True, fuse_adam=False)
    for key in ("image", "k", "params", "m", "v", "loss"):
        assert torch.equal(a[key], b[key]), key


@pytest.mark.parametrize("h,w,n_mov", [(320, 320, 4), (64, 48, 2)])
def test_deterministic_path_agrees_with_atomic_path(h, w, n_mov):
    """Same kernels up to summation order: first iteration's forward is identical, the first iterations' losses
    agree to rounding, the image cotangent (fixed point vs float atomics) to 1e-6."""
    det = _engine_run(h, w, n_mov, 1002, 6, True)
    atm = _engine_run(h, w, n_mov, 1002, 6, False)
    rel = np.abs(det["trace"] - atm["trace"]) / np.abs(atm["trace"])
    print(f"deterministic vs atomic loss: {rel}")
    assert rel[0] < 1e-6 and rel[:4].max() < 1e-4
    # one iteration only: gradients of both paths on identical parameters
    d1 = _engine_run(h, w, n_mov, 1002, 1, True, fuse_adam=False)
    a1 = _engine_run(h, w, n_mov, 1002, 1, False)
    assert torch.equal(d1["k"], a1["k"]) and torch.equal(d1["image"], a1["image"])
    moved = (d1["params_ref"] - a1["params_ref"]).abs() > 1e-3        # Adam's first step is lr * sign(g)
    assert float(moved.float().mean()) < 1e-3


def test_image_cotangent_fixed_point_matches_float_accumulation():
    """d_image of one iteration: 64-bit fixed-point accumulation vs the float-atomic kernels vs the oracle's
    autograd (same inputs)."""
    h, w, n_mov = 64, 48, 2
    outs = {}
    for det in (True, False):
        case = orc.make_case(h, w, n_mov, 5)
        p_img, p_mot = case_params(5, DEV)
        p_mot = p_mot.clone()
        p_mot[2048:3072] *= 10.0
        p_mot[3072:] *= 300.0
        model = mb.IMMoCo(case["masks"].to(DEV))
        eng = mb.FitEngine(model, 10, deterministic=det, fuse_adam=False)
        k = case["kspace_motion"]
        eng.set_kspace((k / k.abs().max() * 16000).to(DEV))
        eng.reset(p_img, p_mot)
        mb.lib().immoco_set_branch_overlap(0)
        try:
            eng.run(mb.lambda_schedule(10, 1e-2), 1e-2, 0, 1)
            torch.cuda.synchronize()
        finally:
            mb.lib().immoco_set_branch_overlap(1)
        outs[det] = (eng.d_image.clone(), eng.d_disp.clone())
        if det:
            assert int(eng.d_image_fx.abs().max()) == 0          # plane re-zeroed by the finalize kernel
            assert int(eng.dc_max_bits[0]) > 0
    assert rel_l2(outs[True][0], outs[False][0]) < 1e-6
    assert torch.equal(outs[True][1], outs[False][1])            # displacement cotangent: no accumulation involved


def test_reconstruct_batch_is_bit_identical_to_single_calls_when_deterministic():
    from miccai24_immoco_b200 import reconstruct_batch
    iters, ks, ms, pis, pms = 12, [], [], [], []
    for s, (h, w, m) in enumerate([(64, 48, 2), (48, 40, 1), (64, 48, 3), (32, 32, 0)]):
        case = orc.make_case(h, w, max(m, 1), 20 + s)
        ks.append(case["kspace_motion"])
        ms.append(case["masks"][:m])
        pi, pm = case_params(20 + s)
        pis.append(pi)
        pms.append(pm)
    imgs, ksp, traces = reconstruct_batch(ks, ms, iters, in_flight=3, chunk=5, image_params=pis, motion_params=pms,
                                          return_kspace=True, return_traces=True, deterministic=True)
    for i in range(len(ks)):
        im1, k1, tr1 = mb.imcoco_motion_correction(ks[i].to(DEV), ms[i].to(DEV), iters=iters, image_params=pis[i],
                                                   motion_params=pms[i], return_trace=True, deterministic=True)
        assert torch.equal(imgs[i], im1) and torch.equal(ksp[i], k1), i
        assert np.array_equal(traces[i], tr1), i


def _batch_engines(shape, seeds, iters, deterministic):
    h, w, n_mov = shape
    engines, lam = [], mb.lambda_schedule(max(iters, 10), 1e-2)[:iters]
    for seed in seeds:
        case = orc.make_case(h, w, max(n_mov, 1), seed)
        p_img, p_mot = case_params(seed, DEV)
        p_mot = p_mot.clone()
        p_mot[2048:3072] *= 10.0
        p_mot[3072:] *= 300.0
        model = mb.IMMoCo(case["masks"][:n_mov].to(DEV))
        eng = mb.FitEngine(model, iters, deterministic=deterministic)
        k = case["kspace_motion"]
        eng.set_kspace((k / k.abs().max() * 16000).to(DEV))
        eng.reset(p_img, p_mot)
        engines.append(eng)
    return engines, lam


def _state(eng, lam):
    return {"image": eng.image.clone(), "k": eng.k_out.clone(), "params": eng.params.clone(), "m": eng.state[1].clone(),
            "v": eng.state[2].clone(), "loss": eng.loss[:len(lam)].clone(), "trace": eng.loss_trace(lam).copy()}


@pytest.mark.parametrize("shape,n_batch", [((320, 320, 4), 4), ((64, 48, 2), 3), ((640, 368, 5), 2), ((32, 32, 0), 2),
                                           ((48, 40, 1), 8)])
def test_batched_fit_is_bit_identical_to_single_fits(shape, n_batch):
    """immoco_fit_run_batched (instance dimension on the latency-bound kernels): every instance of the batch ends
    exactly where its own immoco_fit_run puts it -- deterministic mode, so bit for bit -- also when the batch is
    advanced in several calls."""
    iters = 10 if shape[0] >= 320 else 16
    seeds = [30 + i for i in range(n_batch)]
    singles, lam = _batch_engines(shape, seeds, iters, True)
    for e in singles:
        e.run(lam, 1e-2)
    torch.cuda.synchronize()
    want = [_state(e, lam) for e in singles]
    del singles
    for chunks in ([(0, iters)], [(0, 3), (3, iters)]):
        batch, _ = _batch_engines(shape, seeds, iters, True)
        for a, b in chunks:
            mb.run_batched(batch, lam, 1e-2, a, b)
        torch.cuda.synchronize()
        for i, e in enumerate(batch):
            _assert_identical(_state(e, lam), want[i], f"instance {i} of {n_batch}, chunks {chunks}")
        del batch


def test_batched_fit_atomic_mode_agrees_with_single_fits():
    shape, iters, seeds = (64, 48, 2), 12, [41, 42, 43, 44]
    singles, lam = _batch_engines(shape, seeds, iters, False)
    for e in singles:
        e.run(lam, 1e-2)
    batch, _ = _batch_engines(shape, seeds, iters, False)
    mb.run_batched(batch, lam, 1e-2)
    torch.cuda.synchronize()
    for a, b in zip(batch, singles):
        ta, tb = a.loss_trace(lam), b.loss_trace(lam)
        rel = np.abs(ta - tb) / np.abs(tb)
        assert rel[:4].max() < 1e-5 and rel.max() < 5e-3, rel
    with pytest.raises(ValueError):          # different shapes cannot share a batch
        other, _ = _batch_engines((48, 40, 2), [50], iters, False)
        mb.run_batched([batch[0], other[0]], lam, 1e-2)


def test_reconstruct_batch_groups_equal_shapes_and_stays_bit_identical():
    from miccai24_immoco_b200 import reconstruct_batch
    iters, ks, ms, pis, pms = 12, [], [], [], []
    shapes = [(64, 48, 2), (48, 40, 1), (64, 48, 2), (64, 48, 2), (48, 40, 1), (64, 48, 3), (64, 48, 2), (64, 48, 2)]
    for s, (h, w, m) in enumerate(shapes):
        case = orc.make_case(h, w, m, 60 + s)
        ks.append(case["kspace_motion"])
        ms.append(case["masks"][:m])
        pi, pm = case_params(60 + s)
        pis.append(pi)
        pms.append(pm)
    imgs, ksp, traces = reconstruct_batch(ks, ms, iters, image_params=pis, motion_params=pms, return_kspace=True,
                                          return_traces=True, deterministic=True, batch=4)
    for i in range(len(ks)):
        im1, k1, tr1 = mb.imcoco_motion_correction(ks[i].to(DEV), ms[i].to(DEV), iters=iters, image_params=pis[i],
                                                   motion_params=pms[i], return_trace=True, deterministic=True)
        assert torch.equal(imgs[i], im1) and torch.equal(ksp[i], k1), i
        assert np.array_equal(traces[i], tr1), i


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process_gives_the_same_bits():
    """launch attributes (dynamic shared-memory opt-in) and the SM count are per device, and the
    native calls must run on the tensors' device whatever the current device is.  A fit on cuda:1, issued while
    cuda:0 is current and after cuda:0 has used every kernel, must equal the fit on cuda:0 bit for bit
    (deterministic mode); module-mode operators are exercised on cuda:1 too."""
    case = orc.make_case(64, 48, 2, 11)
    p_img, p_mot = case_params(11)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        torch.cuda.set_device(0)                      # current device stays cuda:0 throughout
        im, k, tr = mb.imcoco_motion_correction(case["kspace_motion"].to(dev), case["masks"].to(dev), iters=12,
                                                image_params=p_img.to(dev), motion_params=p_mot.to(dev),
                                                return_trace=True, deterministic=True)
        torch.cuda.synchronize(dev)
        outs.append((im.cpu(), k.cpu(), tr))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])
    x = torch.randn(64, 48, dtype=torch.complex64, device="cuda:1")
    assert rel_l2(mb.IFFT(mb.FFT(x)), x) < 2e-6
    model = mb.IMMoCo(case["masks"].to("cuda:1"))
    kf, img = model()
    assert kf.device.index == 1 and bool(torch.isfinite(torch.view_as_real(kf)).all())
