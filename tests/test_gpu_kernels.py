"""GPU: every CUDA entry point against the oracle (torch fp32) on the same seeded inputs.
All calls go through the C ABI (ctypes) via the reference-facing Python wrappers."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import miccai24_immoco_b200 as mb
from miccai24_immoco_b200 import _native as nat
from miccai24_immoco_b200.encoding import grid_spec, mlp_spec
from miccai24_immoco_b200.ops import twiddle_table
from oracle import immoco_oracle as orc
from tests.gpu_util import case_params, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _s():
    return torch.cuda.current_stream().cuda_stream


@pytest.fixture(scope="module", autouse=True)
def _fp32_reference():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    mb.build()
    yield


def _coords(dims, kind):
    if kind == "grid":
        return (orc.identity_grid(320, 320).view(-1, 2) if dims == 2 else orc.make_grids((4, 320, 320))).contiguous()
    g = torch.Generator().manual_seed(dims * 7 + len(kind))
    n = {"ragged": 1001, "tiny": 1}[kind]
    return (torch.rand(n, dims, generator=g) * 3.0 - 1.5).contiguous()       # incl. negative cells


@pytest.fixture
def hashgrid_impl(native_lib, request):
    native_lib.immoco_set_hashgrid_impl(request.param)
    yield request.param
    native_lib.immoco_set_hashgrid_impl(1)


@pytest.mark.parametrize("hashgrid_impl", [1, 0], ids=["lane-pair", "thread-per-point"], indirect=True)
@pytest.mark.parametrize("dims", [2, 3])
@pytest.mark.parametrize("kind", ["grid", "ragged", "tiny"])
def test_hashgrid_forward_and_backward(native_lib, dims, kind, hashgrid_impl):
    gs = grid_spec(dims, mb.encoding_config)
    lv = orc.make_grid_levels(dims, orc.ENCODING_CONFIG)
    x = _coords(dims, kind).to(DEV)
    n = x.shape[0]
    g = torch.Generator().manual_seed(3)
    table = ((torch.rand(gs.n_rows, 2, generator=g) * 2 - 1) * 1e-1).to(DEV).requires_grad_(True)
    enc = torch.empty((16, n, 2), device=DEV)
    d = gs.desc()
    nat.check(native_lib.immoco_hashgrid_fwd(C.byref(d), x.data_ptr(), table.data_ptr(), enc.data_ptr(), n, _s()), "fwd")
    ref = orc.hashgrid_encode(x, table, lv, cache=False)                      # (N, 32)
    got = enc.permute(1, 0, 2).reshape(n, 32)
    assert rel_l2(got, ref) < 1e-6
    assert float((got - ref).abs().max()) < 1e-6
    # backward: scatter-add of a random cotangent
    d_enc = torch.randn(16, n, 2, generator=torch.Generator().manual_seed(4)).to(DEV)
    grad = torch.zeros_like(table)
    nat.check(native_lib.immoco_hashgrid_bwd(C.byref(d), x.data_ptr(), d_enc.data_ptr(), grad.data_ptr(), n, _s()), "bwd")
    (ref * d_enc.permute(1, 0, 2).reshape(n, 32)).sum().backward()
    assert rel_l2(grad, table.grad) < 2e-6
    untouched = table.grad.abs().sum(1) == 0
    assert bool((grad[untouched] == 0).all())


class _SimtChecker:
    """tests/checkers/_mlp_simt.so: the fp32 SIMT MLP kernels, a test-side library with the product entry
    points' contracts (the product library itself only contains the tcgen05 kernels)."""

    def __init__(self):
        import __graft_entry__ as entry
        h = C.CDLL(entry.build_checkers())
        _P = C.c_void_p
        h.immoco_simt_mlp_fwd.restype = C.c_int
        h.immoco_simt_mlp_fwd.argtypes = [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P]
        h.immoco_simt_mlp_bwd.restype = C.c_int
        h.immoco_simt_mlp_bwd.argtypes = [_P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]
        self.immoco_mlp_fwd = h.immoco_simt_mlp_fwd
        self.immoco_mlp_bwd = h.immoco_simt_mlp_bwd


@pytest.mark.parametrize("impl", [1, 0], ids=["tcgen05", "simt-checker"])
@pytest.mark.parametrize("width,act", [(256, "relu"), (64, "tanh"), (64, "relu"), (256, "tanh")])
@pytest.mark.parametrize("n", [1, 127, 128, 1000, 40000])
def test_mlp_forward_and_backward(native_lib, width, act, n, impl):
    """impl 1 = tcgen05 kind::tf32 with the 3xTF32 split (the product library), 0 = the fp32 SIMT check kernels of
    the test-side library; both must meet the same fp32 tolerances against torch fp32."""
    assert not hasattr(native_lib, "immoco_set_mlp_impl")      # the product library has no SIMT MLP any more
    _mlp_case(native_lib if impl == 1 else _SimtChecker(), width, act, n)


def _mlp_case(native_lib, width, act, n):
    g = torch.Generator().manual_seed(width + n)
    enc = (torch.randn(16, n, 2, generator=g) * 0.5).to(DEV)
    w1 = (torch.randn(width, 32, generator=g) * 0.2).to(DEV).requires_grad_(True)
    w2 = (torch.randn(16, width, generator=g) * 0.2).to(DEV).requires_grad_(True)
    e = enc.permute(1, 0, 2).reshape(n, 32).clone().requires_grad_(True)
    f = torch.relu if act == "relu" else torch.tanh
    code = nat.ACT_RELU if act == "relu" else nat.ACT_TANH
    for out_tanh in (0, 1):
        out = torch.empty((n, 2), device=DEV)
        nat.check(native_lib.immoco_mlp_fwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), out.data_ptr(), n,
                                            width, code, out_tanh, _s()), "mlp_fwd")
        ref = (f(e @ w1.t()) @ w2.t())[:, :2]
        if out_tanh:
            ref = ref.tanh()
        assert rel_l2(out, ref) < 2e-6
    ref = (f(e @ w1.t()) @ w2.t())[:, :2]
    d_out = torch.randn(n, 2, generator=g).to(DEV)
    (ref * d_out).sum().backward()
    d_enc = torch.empty_like(enc)
    g1 = torch.zeros_like(w1)
    g2 = torch.zeros_like(w2)
    nat.check(native_lib.immoco_mlp_bwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), d_out.data_ptr(),
                                        d_enc.data_ptr(), g1.data_ptr(), g2.data_ptr(), n, width, code, _s()), "mlp_bwd")
    assert rel_l2(d_enc.permute(1, 0, 2).reshape(n, 32), e.grad) < 5e-6
    assert rel_l2(g1, w1.grad) < 5e-6
    assert rel_l2(g2[:2], w2.grad[:2]) < 5e-6
    assert bool((g2[2:] == 0).all())                                          # padded rows get no gradient


@pytest.mark.parametrize("dims,cfg", [(2, "image"), (3, "motion")])
def test_network_with_input_encoding_module(dims, cfg):
    net_cfg = mb.network_config if cfg == "image" else mb.mot_network_config
    o = orc.NetworkWithInputEncoding(dims, 2, orc.ENCODING_CONFIG, net_cfg, seed=11).to(DEV)
    m = mb.NetworkWithInputEncoding(dims, 2, mb.encoding_config, net_cfg, seed=99)
    assert m.params.shape == o.params.shape and m.params.is_cuda
    with torch.no_grad():
        m.params.copy_(o.params)
    x = (orc.identity_grid(96, 80).view(-1, 2) if dims == 2 else orc.make_grids((2, 48, 40))).to(DEV)
    ya, yb = m(x), o(x)
    assert ya.shape == yb.shape and ya.dtype == torch.float32
    assert rel_l2(ya, yb) < 1e-5
    d = torch.randn(ya.shape, generator=torch.Generator().manual_seed(1)).to(DEV)
    (ya * d).sum().backward()
    (yb * d).sum().backward()
    n_mlp = m.mlp.n_params
    assert rel_l2(m.params.grad[:n_mlp], o.params.grad[:n_mlp]) < 1e-4
    assert rel_l2(m.params.grad[n_mlp:], o.params.grad[n_mlp:]) < 1e-4
    with pytest.raises(RuntimeError):
        m(x.cpu())


@pytest.mark.parametrize("shape", [(320, 320), (640, 368), (64, 46), (3, 32, 20), (2, 2, 16, 24)])
def test_fft_ifft_and_adjoints(shape):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.complex(torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)).to(DEV)
    ref = orc.FFT(x)
    assert rel_l2(mb.FFT(x), ref) < 2e-6
    assert rel_l2(mb.IFFT(x), orc.IFFT(x)) < 2e-6
    assert rel_l2(mb.IFFT(mb.FFT(x)), x) < 2e-6
    # adjoint through autograd
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    c = torch.complex(torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)).to(DEV)
    (mb.FFT(xa) * c).real.sum().backward()
    (orc.FFT(xb) * c).real.sum().backward()
    assert rel_l2(xa.grad, xb.grad) < 2e-6
    xa.grad = None
    xb.grad = None
    (mb.IFFT(xa) * c).real.sum().backward()
    (orc.IFFT(xb) * c).real.sum().backward()
    assert rel_l2(xa.grad, xb.grad) < 2e-6


def test_fft_golden_from_reference(golden_dir):
    ops = np.load(os.path.join(golden_dir, "ops_small.npz"))
    for hw in ("32x32", "48x20", "64x46"):
        x = torch.from_numpy(ops[f"fft_in_{hw}"]).to(DEV)
        assert rel_l2(mb.FFT(x), torch.from_numpy(ops[f"fft_out_{hw}"])) < 2e-6
    with pytest.raises(NotImplementedError):
        mb.FFT(torch.zeros(7, 9, dtype=torch.complex64, device=DEV))          # odd sizes: not on the path


def test_gradient_entropy_value_and_grad(golden_dir):
    ops = np.load(os.path.join(golden_dir, "ops_small.npz"))
    x = torch.from_numpy(ops["ge_in"]).to(DEV).requires_grad_(True)
    val = mb.GradientEntropyLoss()(x)
    val.backward()
    assert abs(float(val) - float(ops["ge_val"])) <= 2e-6 * abs(float(ops["ge_val"]))
    assert rel_l2(x.grad, torch.from_numpy(ops["ge_grad"])) < 5e-6
    g = torch.Generator().manual_seed(5)
    y = torch.complex(torch.randn(320, 320, generator=g), torch.randn(320, 320, generator=g)).to(DEV)
    y[10, 10] = y[10, 11]
    ya = y.clone().requires_grad_(True)
    yb = y.clone().requires_grad_(True)
    a = mb.GradientEntropyLoss()(ya)
    b = orc.gradient_entropy(yb)
    (a * 0.37).backward()
    (b * 0.37).backward()
    assert abs(float(a) - float(b)) <= 2e-6 * abs(float(b))
    assert rel_l2(ya.grad, yb.grad) < 5e-6


def test_adam_step_matches_torch(native_lib):
    n = 1_000_002                                                            # exercises the scalar tail
    g = torch.Generator().manual_seed(9)
    p = torch.randn(n, generator=g).to(DEV)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-2)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 5):
        grad = (torch.randn(n, generator=g) * (10.0 ** float(torch.randint(-6, 3, (1,), generator=g)))).to(DEV)
        grad[::7] = 0.0
        gbuf = grad.clone()
        nat.check(native_lib.immoco_adam_step(p.data_ptr(), gbuf.data_ptr(), m.data_ptr(), v.data_ptr(), n,
                                              1e-2, 0.9, 0.999, 1e-8, step, 1, _s()), "adam")
        ref.grad = grad.clone()
        opt.step()
        assert bool((gbuf == 0).all())                                       # zero_grad fused
        assert float((p - ref.detach()).abs().max()) < 2e-6
    st = opt.state[ref]
    assert rel_l2(m, st["exp_avg"]) < 1e-6 and rel_l2(v, st["exp_avg_sq"]) < 1e-6
    # untouched (zero-gradient) entries never move: m = v = 0
    z, zm, zv = (torch.zeros(1024, device=DEV) for _ in range(3))
    p0 = torch.randn(1024, device=DEV)
    p1 = p0.clone()
    nat.check(native_lib.immoco_adam_step(p1.data_ptr(), z.data_ptr(), zm.data_ptr(), zv.data_ptr(),
                                          1024, 1e-2, 0.9, 0.999, 1e-8, 1, 1, _s()), "adam")
    assert torch.equal(p0, p1)


@pytest.mark.parametrize("width,act,n", [(64, "tanh", 409600), (256, "relu", 102400), (64, "tanh", 12345)])
def test_tensor_core_mlp_is_run_to_run_deterministic(native_lib, width, act, n):
    """The tcgen05 kernels hand work between issuing threads, TMEM regions and shared-memory buffers through
    mbarriers only; a missing dependency would show as run-to-run differences.  d_enc and the forward output
    involve no atomics, so they must be BIT-identical over many launches (the weight gradients are summed
    across CTAs with atomics and may differ in the last bits)."""
    a = {"relu": nat.ACT_RELU, "tanh": nat.ACT_TANH}[act]
    g = torch.Generator(device=DEV).manual_seed(n)
    enc = torch.randn(16, n, 2, device=DEV, generator=g) * 3e-2
    w1 = torch.randn(width, 32, device=DEV, generator=g) * 0.3
    w2 = torch.randn(16, width, device=DEV, generator=g) * 0.3
    d_out = torch.randn(n, 2, device=DEV, generator=g)
    outs, d_encs, g1s = [], [], []
    for _ in range(25):
        out = torch.empty(n, 2, device=DEV)
        d_enc = torch.empty_like(enc)
        g1, g2 = torch.zeros_like(w1), torch.zeros_like(w2)
        nat.check(native_lib.immoco_mlp_fwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), out.data_ptr(), n, width, a, 1,
                                            _s()), "fwd")
        nat.check(native_lib.immoco_mlp_bwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), d_out.data_ptr(),
                                            d_enc.data_ptr(), g1.data_ptr(), g2.data_ptr(), n, width, a, _s()), "bwd")
        outs.append(out)
        d_encs.append(d_enc)
        g1s.append(g1)
    for k in range(1, 25):
        assert torch.equal(outs[k], outs[0]), f"forward output differs in launch {k}"
        assert torch.equal(d_encs[k], d_encs[0]), f"d_enc differs in launch {k}"
        assert rel_l2(g1s[k], g1s[0]) < 1e-5


def test_hashgrid_row_swizzle_equals_reference_layout(native_lib):
    """3-D hash grid with the permuted row layout (immoco_grid_desc::swizzle) on a permuted table: features
    bit-identical to the reference layout, table gradients equal after un-permuting (float atomics: rounding)."""
    from miccai24_immoco_b200.encoding import grid_spec
    gs = grid_spec(3, mb.encoding_config)
    coords = mb.make_grids((4, 96, 80), "cuda").contiguous()
    n = coords.shape[0]
    swz = gs.row_swizzle(np.unique(coords[:, 0].cpu().numpy()))
    perm = torch.from_numpy(gs.row_permutation(swz)).cuda()
    g = torch.Generator().manual_seed(7)
    table = ((torch.rand(gs.n_rows, 2, generator=g) - 0.5) * 1e-2).cuda()
    d_enc = torch.randn(16, n, 2, generator=g).cuda()
    table_p = torch.empty_like(table)
    table_p.index_copy_(0, perm, table)
    s = torch.cuda.current_stream().cuda_stream
    res = {}
    for name, desc, tab in (("ref", gs.desc(), table), ("swz", gs.desc(swz), table_p)):
        enc = torch.empty(16, n, 2, device="cuda")
        grad = torch.zeros_like(tab)
        assert native_lib.immoco_hashgrid_fwd(C.byref(desc), coords.data_ptr(), tab.data_ptr(), enc.data_ptr(), n, s) == 0
        assert native_lib.immoco_hashgrid_bwd(C.byref(desc), coords.data_ptr(), d_enc.data_ptr(), grad.data_ptr(), n, s) == 0
        torch.cuda.synchronize()
        res[name] = (enc, grad)
    assert torch.equal(res["swz"][0], res["ref"][0])
    g_back = res["swz"][1][perm]
    assert float((g_back - res["ref"][1]).norm() / res["ref"][1].norm()) < 1e-6
    for impl in (0,):       # one-thread-per-point check kernels honour the layout word too
        native_lib.immoco_set_hashgrid_impl(impl)
        try:
            enc = torch.empty(16, n, 2, device="cuda")
            desc = gs.desc(swz)
            assert native_lib.immoco_hashgrid_fwd(C.byref(desc), coords.data_ptr(), table_p.data_ptr(), enc.data_ptr(), n, s) == 0
            torch.cuda.synchronize()
        finally:
            native_lib.immoco_set_hashgrid_impl(1)
        assert float((enc - res["ref"][0]).norm() / res["ref"][0].norm()) < 1e-6


@pytest.mark.parametrize("h,w", [(320, 320), (48, 40), (33, 17), (640, 368)])
def test_hashgrid_tap_indexed_storage_equals_reference_layout(native_lib, h, w):
    """2-D hash grid with its hashed levels stored tap-indexed (include/immoco_b200.h section 1c, immoco.py:GridTaps)
    on a permuted table: features BIT-identical to the hashing kernels on the reference layout, table gradients
    equal after un-permuting (float atomics: rounding), rows nobody touches stored behind the live ones."""
    from miccai24_immoco_b200.immoco import GridTaps
    gs = grid_spec(2, mb.encoding_config)
    coords = mb.immoco._identity_grid(h, w, "cuda").view(-1, 2).contiguous()
    n = coords.shape[0]
    desc = gs.desc()
    taps = GridTaps(gs, desc, coords)
    first, base = taps.first_level, gs.offsets[taps.first_level]
    assert first == 6 and taps.rows.shape == (16 - first, n, 4) and taps.rows.data_ptr() % 16 == 0
    perm = taps.perm
    assert torch.equal(torch.sort(perm).values, torch.arange(gs.n_rows, device="cuda"))        # a bijection ...
    assert torch.equal(perm[:base], torch.arange(base, device="cuda"))                          # ... identity on the dense levels
    # the tap list is the permuted image of the hash indices the kernels compute themselves
    raw = torch.empty((16 - first, n, 4), dtype=torch.int32, device="cuda")
    assert native_lib.immoco_hashgrid_tap_rows(C.byref(desc), coords.data_ptr(), n, first, 16, raw.data_ptr(), _s()) == 0
    lvl = torch.tensor(gs.offsets[first:16], device="cuda")[:, None, None]
    assert torch.equal(perm[raw.long() + lvl], taps.rows.long())
    touched = torch.unique(raw.long() + lvl)
    assert taps.n_active_rows == base + touched.numel()
    assert int(perm[touched].max()) == taps.n_active_rows - 1                                   # live rows first, no holes
    # ... ranked by first touch in (level, point, corner) order
    assert int(taps.rows[0, 0, 0]) == base
    g = torch.Generator().manual_seed(11)
    table = ((torch.rand(gs.n_rows, 2, generator=g) - 0.5) * 1e-2).cuda()
    d_enc = torch.randn(16, n, 2, generator=g).cuda()
    d_enc[:, ::7] = 0.0                                          # zero cotangents are skipped
    table_p = torch.empty_like(table)
    table_p.index_copy_(0, perm, table)
    t = taps.struct()
    enc_ref, enc_tap = torch.empty(16, n, 2, device="cuda"), torch.empty(16, n, 2, device="cuda")
    grad_ref, grad_tap = torch.zeros_like(table), torch.zeros_like(table)
    s = _s()
    assert native_lib.immoco_hashgrid_fwd(C.byref(desc), coords.data_ptr(), table.data_ptr(), enc_ref.data_ptr(), n, s) == 0
    assert native_lib.immoco_hashgrid_bwd(C.byref(desc), coords.data_ptr(), d_enc.data_ptr(), grad_ref.data_ptr(), n, s) == 0
    assert native_lib.immoco_hashgrid_fwd_taps(C.byref(desc), C.byref(t), coords.data_ptr(), table_p.data_ptr(),
                                               enc_tap.data_ptr(), n, s) == 0
    assert native_lib.immoco_hashgrid_bwd_taps(C.byref(desc), C.byref(t), coords.data_ptr(), d_enc.data_ptr(),
                                               grad_tap.data_ptr(), n, s) == 0
    torch.cuda.synchronize()
    assert torch.equal(enc_tap, enc_ref)
    assert float((grad_tap[perm] - grad_ref).norm() / grad_ref.norm()) < 1e-6
    assert float(grad_tap[taps.n_active_rows:].abs().max()) == 0.0 if taps.n_active_rows < gs.n_rows else True
    # argument checks: 3-D grids and a tap list of another point count are rejected
    gs3 = grid_spec(3, mb.encoding_config)
    d3 = gs3.desc()
    assert native_lib.immoco_hashgrid_fwd_taps(C.byref(d3), C.byref(t), coords.data_ptr(), table_p.data_ptr(),
                                               enc_tap.data_ptr(), n, s) == nat.ERR_UNSUPPORTED
    assert native_lib.immoco_hashgrid_fwd_taps(C.byref(desc), C.byref(t), coords.data_ptr(), table_p.data_ptr(),
                                               enc_tap.data_ptr(), n - 1, s) == nat.ERR_BAD_ARG


@pytest.mark.parametrize("m,h,w", [(4, 96, 80), (2, 48, 40), (8, 33, 17), (16, 20, 12), (5, 40, 36), (3, 33, 17), (4, 320, 320)])
@pytest.mark.parametrize("layout", ["lut", "swizzle", "reference"])
def test_hashgrid_grouped_kernels_equal_lane_pair_kernels(native_lib, m, h, w, layout):
    """immoco_hashgrid_fwd_grouped / _bwd_grouped (the 2 M lanes of a bundle = all groups x both dim-0 corners of one
    pixel) under the general linear layout (chunk tables), the Gray/exchange word and the reference layout: features
    BIT-identical to the lane-pair kernels on the reference layout, gradients equal after un-permuting."""
    gs = grid_spec(3, mb.encoding_config)
    coords = mb.make_grids((m, h, w), "cuda").contiguous()
    n, p = coords.shape[0], h * w
    u = torch.linspace(-1, 1, m).numpy()
    lut_t = None
    if layout == "lut":
        lut = gs.linear_layout(u)
        words = tuple(nat.LAYOUT_LUT if lut[l].any() else 0 for l in range(16))
        assert sum(1 for x in words if x) == 13
        assert np.array_equal(np.sort(gs.row_permutation_lut(lut)), np.arange(gs.n_rows))
        perm = torch.from_numpy(gs.row_permutation_lut(lut)).cuda()
        lut_t = torch.from_numpy(lut.view(np.int32)).cuda()
        desc = gs.desc(words, lut_t.data_ptr())
    elif layout == "swizzle":
        words = gs.row_swizzle(u)
        perm = torch.from_numpy(gs.row_permutation(words)).cuda()
        desc = gs.desc(words)
    else:
        perm = torch.arange(gs.n_rows, device="cuda")
        desc = gs.desc()
    g = torch.Generator().manual_seed(13 + m)
    table = ((torch.rand(gs.n_rows, 2, generator=g) - 0.5) * 1e-2).cuda()
    d_enc = torch.randn(16, n, 2, generator=g).cuda()
    d_enc[:, ::5] = 0.0
    table_p = torch.empty_like(table)
    table_p.index_copy_(0, perm, table)
    ref_desc = gs.desc()
    s = _s()
    enc_ref, enc_grp, enc_gen = (torch.empty(16, n, 2, device="cuda") for _ in range(3))
    grad_ref, grad_grp, grad_gen = (torch.zeros_like(table) for _ in range(3))
    assert native_lib.immoco_hashgrid_fwd(C.byref(ref_desc), coords.data_ptr(), table.data_ptr(), enc_ref.data_ptr(), n, s) == 0
    assert native_lib.immoco_hashgrid_bwd(C.byref(ref_desc), coords.data_ptr(), d_enc.data_ptr(), grad_ref.data_ptr(), n, s) == 0
    assert native_lib.immoco_hashgrid_fwd_grouped(C.byref(desc), coords.data_ptr(), table_p.data_ptr(), enc_grp.data_ptr(), p, m, s) == 0
    assert native_lib.immoco_hashgrid_bwd_grouped(C.byref(desc), coords.data_ptr(), d_enc.data_ptr(), grad_grp.data_ptr(), p, m, s) == 0
    # the generic entry points honour the same descriptor (chunk tables: one-thread-per-point kernels)
    assert native_lib.immoco_hashgrid_fwd(C.byref(desc), coords.data_ptr(), table_p.data_ptr(), enc_gen.data_ptr(), n, s) == 0
    assert native_lib.immoco_hashgrid_bwd(C.byref(desc), coords.data_ptr(), d_enc.data_ptr(), grad_gen.data_ptr(), n, s) == 0
    torch.cuda.synchronize()
    assert torch.equal(enc_grp, enc_ref)
    assert float((enc_gen - enc_ref).norm() / enc_ref.norm()) < 1e-6
    assert float((grad_grp[perm] - grad_ref).norm() / grad_ref.norm()) < 1e-6
    assert float((grad_gen[perm] - grad_ref).norm() / grad_ref.norm()) < 1e-6
    if layout == "lut":
        # group counts the bundles cannot hold, and consumers that only know the Gray/exchange word, are refused
        for bad in (1, 17):
            assert native_lib.immoco_hashgrid_fwd_grouped(C.byref(desc), coords.data_ptr(), table_p.data_ptr(),
                                                          enc_grp.data_ptr(), p, bad, s) == nat.ERR_UNSUPPORTED
        row_ptr = torch.empty(gs.n_rows + 1, dtype=torch.int32, device="cuda")
        assert native_lib.immoco_hashgrid_csr_build(C.byref(desc), coords.data_ptr(), n, row_ptr.data_ptr(),
                                                    row_ptr.data_ptr(), row_ptr.data_ptr(), 1 << 20, s) == nat.ERR_UNSUPPORTED


# ------------------------------------------------------------------------------------------------------------
# deterministic building blocks (include/immoco_b200.h sections 1b, 2, 7)
# ------------------------------------------------------------------------------------------------------------
def _build_csr(native_lib, gs, desc, x):
    from miccai24_immoco_b200.immoco import GridCsr
    return GridCsr(gs, desc, x)


@pytest.mark.parametrize("dims,kind", [(2, "grid"), (3, "grid"), (2, "ragged"), (3, "ragged"), (3, "tiny"), (2, "small"),
                                       (3, "small")])
def test_hashgrid_csr_structure_and_gather_backward(native_lib, dims, kind):
    """The row-sorted tap list equals the oracle's taps (same rows, same (point, corner) order within a row,
    same weights bit for bit); the gather backward equals autograd of the oracle encoding and is bit-identical
    run to run; rows nobody touches are not written."""
    gs = grid_spec(dims, mb.encoding_config)
    lv = orc.make_grid_levels(dims, orc.ENCODING_CONFIG)
    if kind == "small":
        x = (orc.identity_grid(48, 40).view(-1, 2) if dims == 2 else orc.make_grids((2, 48, 40))).contiguous().to(DEV)
    else:
        x = _coords(dims, kind).to(DEV)
    n = x.shape[0]
    d = gs.desc()
    csr = _build_csr(native_lib, gs, d, x)
    torch.cuda.synchronize()
    row_ptr = csr.row_ptr.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    taps = csr.taps.cpu().numpy()
    c = 1 << dims
    assert row_ptr[0] == 0 and row_ptr[-1] == n * c * 16 and np.all(np.diff(row_ptr) >= 0)
    if n <= 4000:        # full structural check against the oracle's taps
        idx, w = orc.all_taps(x.cpu(), lv)                       # (L, C, N)
        idx, w = idx.numpy(), w.numpy()
        rows = np.repeat(np.arange(gs.n_rows), np.diff(row_ptr))
        word = taps[:, 0].astype(np.int64) & 0xFFFFFFFF
        pts = word & 0x7FFFFFF                     # plane index = level * n + point ...
        rid = word >> 27                           # ... | (row within its level & 31) << 27
        wts = taps[:, 1].view(np.float32)
        for level in range(16):
            sl = slice(level * n * c, (level + 1) * n * c)
            # oracle taps of the level ordered by (row, point, corner)
            o_rows = idx[level].T.reshape(-1)                    # point-major, corner-minor
            o_pts = np.repeat(np.arange(n), c)
            o_w = w[level].T.reshape(-1)
            order = np.lexsort((np.arange(n * c), o_rows))
            assert np.array_equal(rows[sl], o_rows[order])
            assert np.array_equal(pts[sl], level * n + o_pts[order])
            assert np.array_equal(rid[sl], (o_rows[order] - gs.offsets[level]) & 31)
            assert np.allclose(wts[sl], o_w[order], rtol=0, atol=1e-7)
    # gather backward vs autograd of the oracle
    g = torch.Generator().manual_seed(3)
    table = ((torch.rand(gs.n_rows, 2, generator=g) * 2 - 1) * 1e-1).to(DEV).requires_grad_(True)
    d_enc = torch.randn(16, n, 2, generator=torch.Generator().manual_seed(4)).to(DEV)
    ref = orc.hashgrid_encode(x, table, lv, cache=False)
    (ref * d_enc.permute(1, 0, 2).reshape(n, 32)).sum().backward()
    cs = csr.struct()
    grads = []
    for _ in range(3):
        grad = torch.full_like(table, 7.0)        # sentinel: untouched rows must keep it
        nat.check(native_lib.immoco_hashgrid_bwd_csr(C.byref(d), C.byref(cs), d_enc.data_ptr(), grad.data_ptr(), _s()), "bwd_csr")
        grads.append(grad)
    touched = torch.from_numpy(np.diff(row_ptr) > 0).to(DEV)
    assert bool((grads[0][~touched] == 7.0).all())
    assert bool((table.grad[~touched] == 0).all())
    assert rel_l2(grads[0][touched], table.grad[touched]) < 2e-6
    assert torch.equal(grads[1], grads[0]) and torch.equal(grads[2], grads[0])
    # the atomic scatter gives the same values up to summation order
    grad_a = torch.zeros_like(table)
    nat.check(native_lib.immoco_hashgrid_bwd(C.byref(d), x.data_ptr(), d_enc.data_ptr(), grad_a.data_ptr(), n, _s()), "bwd")
    assert rel_l2(grads[0][touched], grad_a[touched]) < 2e-6


def test_hashgrid_csr_honours_row_swizzle(native_lib):
    gs = grid_spec(3, mb.encoding_config)
    coords = mb.make_grids((4, 64, 48), "cuda").contiguous()
    swz = gs.row_swizzle(np.linspace(-1, 1, 4, dtype=np.float32))
    assert any(swz)
    perm = torch.from_numpy(gs.row_permutation(swz)).cuda()
    d_enc = torch.randn(16, coords.shape[0], 2, generator=torch.Generator().manual_seed(1)).cuda()
    out = {}
    for name, desc in (("ref", gs.desc()), ("swz", gs.desc(swz))):
        csr = _build_csr(native_lib, gs, desc, coords)
        cs = csr.struct()
        grad = torch.zeros(gs.n_rows, 2, device="cuda")
        nat.check(native_lib.immoco_hashgrid_bwd_csr(C.byref(desc), C.byref(cs), d_enc.data_ptr(), grad.data_ptr(), _s()), "csr")
        out[name] = grad
    # same taps per logical row; the summation tree depends on the neighbouring rows of the physical layout (a warp
    # walks the taps of its 32 rows in steps of 32), so the two layouts agree to rounding, not bit for bit
    assert rel_l2(out["swz"][perm], out["ref"]) < 1e-6


def test_hashgrid_csr_fused_adam_equals_gather_then_adam(native_lib):
    """Gather + Adam in one kernel == gather, then immoco_adam_step on the table, bit for bit, over several
    steps; rows nobody touches never move."""
    gs = grid_spec(3, mb.encoding_config)
    x = mb.make_grids((2, 48, 40), "cuda").contiguous()
    d = gs.desc()
    csr = _build_csr(native_lib, gs, d, x)
    cs = csr.struct()
    g = torch.Generator().manual_seed(5)
    p0 = ((torch.rand(gs.n_rows, 2, generator=g) * 2 - 1) * 1e-4).cuda()
    pa, ma, va = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    pb, mb_, vb = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    grad = torch.zeros_like(p0)
    for step in range(1, 5):
        d_enc = (torch.randn(16, x.shape[0], 2, generator=g) * 10.0 ** float(step - 3)).cuda()
        nat.check(native_lib.immoco_hashgrid_bwd_csr_adam(C.byref(d), C.byref(cs), d_enc.data_ptr(), pa.data_ptr(),
                                                          ma.data_ptr(), va.data_ptr(), None, 1e-2, 0.9, 0.999, 1e-8,
                                                          step, _s()), "csr_adam")
        nat.check(native_lib.immoco_hashgrid_bwd_csr(C.byref(d), C.byref(cs), d_enc.data_ptr(), grad.data_ptr(), _s()), "csr")
        nat.check(native_lib.immoco_adam_step(pb.data_ptr(), grad.data_ptr(), mb_.data_ptr(), vb.data_ptr(), pb.numel(),
                                              1e-2, 0.9, 0.999, 1e-8, step, 0, _s()), "adam")
        assert torch.equal(pa, pb) and torch.equal(ma, mb_) and torch.equal(va, vb), step
    row_ptr = csr.row_ptr.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    untouched = torch.from_numpy(np.diff(row_ptr) == 0).cuda()
    assert int(untouched.sum()) > 0 and torch.equal(pa[untouched], p0[untouched])


@pytest.mark.parametrize("width,act,n", [(64, "tanh", 409600), (256, "relu", 102400), (64, "tanh", 1000), (256, "relu", 77)])
def test_mlp_backward_partials_are_deterministic_and_sum_to_the_gradient(native_lib, width, act, n):
    """Per-CTA weight-gradient blocks: bit-identical over launches, their ordered sum equals the atomically
    accumulated gradient (and torch's), d_enc is the same tensor as in the atomic mode; Adam over the blocks
    equals Adam over their sum."""
    a = {"relu": nat.ACT_RELU, "tanh": nat.ACT_TANH}[act]
    g = torch.Generator(device=DEV).manual_seed(n + width)
    enc = torch.randn(16, n, 2, device=DEV, generator=g) * 0.3
    w = torch.randn(width * 32 + 16 * width, device=DEV, generator=g) * 0.2
    w1, w2 = w[: width * 32], w[width * 32:]
    d_out = torch.randn(n, 2, device=DEV, generator=g)
    n_mlp = w.numel()
    n_part = native_lib.immoco_mlp_bwd_partial_count(n)
    assert 1 <= n_part <= torch.cuda.get_device_properties(0).multi_processor_count
    parts, d_encs = [], []
    for _ in range(4):
        part = torch.zeros(n_part, n_mlp, device=DEV)
        d_enc = torch.empty_like(enc)
        nat.check(native_lib.immoco_mlp_bwd_partials(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), d_out.data_ptr(),
                                                     d_enc.data_ptr(), part.data_ptr(), n, width, a, _s()), "partials")
        parts.append(part)
        d_encs.append(d_enc)
    for k in range(1, 4):
        assert torch.equal(parts[k], parts[0]) and torch.equal(d_encs[k], d_encs[0])
    g_at = torch.zeros(n_mlp, device=DEV)
    d_enc_at = torch.empty_like(enc)
    nat.check(native_lib.immoco_mlp_bwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), d_out.data_ptr(), d_enc_at.data_ptr(),
                                        g_at.data_ptr(), g_at.data_ptr() + 4 * width * 32, n, width, a, _s()), "bwd")
    assert torch.equal(d_enc_at, d_encs[0])
    g_sum = parts[0].double().sum(0).float()
    assert rel_l2(g_sum[: width * 32], g_at[: width * 32]) < 2e-6
    assert rel_l2(g_sum[width * 32: width * 34], g_at[width * 32: width * 34]) < 2e-6
    assert bool((parts[0][:, width * 34:] == 0).all())               # padded W2 rows are never written
    # Adam over the blocks (fixed summation tree inside the kernel: lane l adds blocks l, l + 32, ... in order,
    # then a 5-step butterfly) vs Adam over the same tree evaluated with torch fp32 adds
    lanes = torch.zeros(32, n_mlp, device=DEV)
    for c in range(n_part):
        lanes[c % 32] += parts[0][c]
    idx = torch.arange(32, device=DEV)
    for o in (16, 8, 4, 2, 1):
        lanes = lanes + lanes[idx ^ o]
    g_ord = lanes[0].contiguous()
    pa, ma, va = w.clone(), torch.zeros_like(w), torch.zeros_like(w)
    pb, mb_, vb = w.clone(), torch.zeros_like(w), torch.zeros_like(w)
    nat.check(native_lib.immoco_adam_step_partials(pa.data_ptr(), parts[0].data_ptr(), n_part, ma.data_ptr(), va.data_ptr(),
                                                   n_mlp, 1e-2, 0.9, 0.999, 1e-8, 1, _s()), "adam_partials")
    nat.check(native_lib.immoco_adam_step(pb.data_ptr(), g_ord.data_ptr(), mb_.data_ptr(), vb.data_ptr(), n_mlp, 1e-2, 0.9,
                                          0.999, 1e-8, 1, 0, _s()), "adam")
    assert torch.equal(pa, pb) and torch.equal(ma, mb_) and torch.equal(va, vb)


def test_release_streams(native_lib):
    """The library's only hidden state -- one auxiliary stream set per caller stream -- is freed on request."""
    case = orc.make_case(32, 32, 1, 6)
    mb.imcoco_motion_correction(case["kspace_motion"], case["masks"], 10)
    torch.cuda.synchronize()
    assert native_lib.immoco_release_streams() >= 1
    assert native_lib.immoco_release_streams() == 0
    im, _ = mb.imcoco_motion_correction(case["kspace_motion"], case["masks"], 10)      # re-created on demand
    torch.cuda.synchronize()
    assert bool(torch.isfinite(torch.view_as_real(im)).all())


@pytest.mark.parametrize("dims", [2, 3])
@pytest.mark.parametrize("kind", ["grid", "ragged"])
def test_hashgrid_stride_wrap_mode(native_lib, dims, kind):
    """encoding_config["stride_wrap"] = True: grid_index()'s stride kept in a uint32 like tiny-cuda-nn does --
    levels 12-15 (resolution >= 2^16) skip the hash and index densely with the wrapped strides.  Same kernels
    (the level kinds come from the host); forward, atomic backward and gather backward against the oracle."""
    cfg_o = dict(orc.ENCODING_CONFIG, stride_wrap=True)
    gs = grid_spec(dims, dict(mb.encoding_config, stride_wrap=True))
    lv = orc.make_grid_levels(dims, cfg_o)
    assert gs.hashed[12:] == (0, 0, 0, 0) and grid_spec(dims, mb.encoding_config).hashed[12:] == (1, 1, 1, 1)
    assert tuple(int(v) for v in lv.hashed) == gs.hashed
    x = _coords(dims, kind).to(DEV)
    n = x.shape[0]
    g = torch.Generator().manual_seed(3)
    table = ((torch.rand(gs.n_rows, 2, generator=g) * 2 - 1) * 1e-1).to(DEV).requires_grad_(True)
    enc = torch.empty((16, n, 2), device=DEV)
    d = gs.desc()
    nat.check(native_lib.immoco_hashgrid_fwd(C.byref(d), x.data_ptr(), table.data_ptr(), enc.data_ptr(), n, _s()), "fwd")
    ref = orc.hashgrid_encode(x, table, lv, cache=False)
    got = enc.permute(1, 0, 2).reshape(n, 32)
    assert rel_l2(got, ref) < 1e-6
    # the two modes really differ on the finest levels
    lv0 = orc.make_grid_levels(dims, orc.ENCODING_CONFIG)
    ref0 = orc.hashgrid_encode(x, table, lv0, cache=False)
    assert rel_l2(ref[:, 24:], ref0[:, 24:]) > 1e-2 and torch.equal(ref[:, :24], ref0[:, :24])
    d_enc = torch.randn(16, n, 2, generator=torch.Generator().manual_seed(4)).to(DEV)
    (ref * d_enc.permute(1, 0, 2).reshape(n, 32)).sum().backward()
    # in this mode the third coordinate drops out of levels 12-15, so thousands of points of a 3-D grid pile up on
    # the same rows: sums of that length differ between summation orders at the 1e-5 level in fp32 (measured 6e-6)
    grad = torch.zeros_like(table)
    nat.check(native_lib.immoco_hashgrid_bwd(C.byref(d), x.data_ptr(), d_enc.data_ptr(), grad.data_ptr(), n, _s()), "bwd")
    assert rel_l2(grad, table.grad) < 2e-5
    csr = _build_csr(native_lib, gs, d, x)
    cs = csr.struct()
    grad_c = torch.zeros_like(table)
    nat.check(native_lib.immoco_hashgrid_bwd_csr(C.byref(d), C.byref(cs), d_enc.data_ptr(), grad_c.data_ptr(), _s()), "csr")
    assert rel_l2(grad_c, table.grad) < 2e-5
    assert rel_l2(grad_c.double(), table.grad.double()) < 2e-5


def test_fit_in_stride_wrap_mode_matches_oracle():
    """A short fit with the tiny-cuda-nn-compatible level kinds on both sides (oracle config key / package default)."""
    from miccai24_immoco_b200 import encoding as enc_mod
    h, w, n_mov, seed, iters = 64, 48, 2, 3, 12
    case = orc.make_case(h, w, n_mov, seed)
    p_img, p_mot = case_params(seed)
    saved = dict(orc.ENCODING_CONFIG)
    try:
        enc_mod.set_stride_wrap_default(True)
        orc.ENCODING_CONFIG["stride_wrap"] = True
        _, _, trace = mb.imcoco_motion_correction(case["kspace_motion"].to(DEV), case["masks"].to(DEV), iters=iters,
                                                  image_params=p_img.to(DEV), motion_params=p_mot.to(DEV),
                                                  return_trace=True, deterministic=True)
        _, _, trace_o = orc.imcoco_motion_correction(case["kspace_motion"], case["masks"], iters=iters,
                                                     image_params=p_img, motion_params=p_mot, return_trace=True)
    finally:
        enc_mod.set_stride_wrap_default(False)
        orc.ENCODING_CONFIG.clear()
        orc.ENCODING_CONFIG.update(saved)
    rel = np.abs(trace - np.asarray(trace_o)) / np.abs(trace_o)
    assert rel[0] < 1e-5 and rel.max() < 1e-3, rel


@pytest.mark.parametrize("shape", [(4, 96, 80), (2, 48, 40), (1, 33, 17)])
@pytest.mark.parametrize("swizzled", [False, True])
def test_mlp_backward_with_fused_scatter_equals_separate_kernels(native_lib, shape, swizzled):
    """immoco_mlp_bwd_scatter + immoco_hashgrid_bwd_dense_levels == immoco_mlp_bwd + immoco_hashgrid_bwd: table
    gradients to the rounding of float atomics, dense-level feature planes bit for bit, weight gradients to rounding."""
    gs = grid_spec(3, mb.encoding_config)
    coords = mb.make_grids(shape, "cuda").contiguous()
    n = coords.shape[0]
    swz = gs.row_swizzle(np.linspace(-1, 1, shape[0], dtype=np.float32)) if swizzled else ()
    d = gs.desc(swz)
    g = torch.Generator(device=DEV).manual_seed(n)
    enc = torch.randn(16, n, 2, device=DEV, generator=g) * 0.3
    w = torch.randn(64 * 32 + 16 * 64, device=DEV, generator=g) * 0.2
    w1, w2 = w[: 64 * 32], w[64 * 32:]
    d_out = torch.randn(n, 2, device=DEV, generator=g)
    d_out[::5] = 0.0                                   # zero cotangents are skipped by both paths
    a = nat.ACT_TANH
    # separate kernels
    d_enc_a = torch.empty_like(enc)
    gw_a = torch.zeros_like(w)
    gt_a = torch.zeros(gs.n_rows, 2, device=DEV)
    nat.check(native_lib.immoco_mlp_bwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), d_out.data_ptr(), d_enc_a.data_ptr(),
                                        gw_a.data_ptr(), gw_a.data_ptr() + 4 * 64 * 32, n, 64, a, _s()), "bwd")
    nat.check(native_lib.immoco_hashgrid_bwd(C.byref(d), coords.data_ptr(), d_enc_a.data_ptr(), gt_a.data_ptr(), n, _s()), "hg")
    # fused
    d_enc_b = torch.full_like(enc, float("nan"))
    gw_b = torch.zeros_like(w)
    gt_b = torch.zeros(gs.n_rows, 2, device=DEV)
    nat.check(native_lib.immoco_mlp_bwd_scatter(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), d_out.data_ptr(),
                                                d_enc_b.data_ptr(), gw_b.data_ptr(), gw_b.data_ptr() + 4 * 64 * 32, C.byref(d),
                                                coords.data_ptr(), gt_b.data_ptr(), n, 64, a, _s()), "bwd_scatter")
    nat.check(native_lib.immoco_hashgrid_bwd_dense_levels(C.byref(d), coords.data_ptr(), d_enc_b.data_ptr(), gt_b.data_ptr(),
                                                          n, _s()), "hg_dense")
    dense = [l for l in range(16) if not gs.hashed[l]]
    hashed = [l for l in range(16) if gs.hashed[l]]
    assert dense == [0, 1, 2]
    assert torch.equal(d_enc_b[dense], d_enc_a[dense])
    assert bool(torch.isnan(d_enc_b[hashed]).all())              # never written: no round trip for those levels
    assert rel_l2(gt_b, gt_a) < 2e-6
    touched = gt_a.abs().sum(1) > 0
    assert bool((gt_b[~touched] == 0).all())
    assert rel_l2(gw_b, gw_a) < 1e-5
    assert native_lib.immoco_mlp_bwd_scatter(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), d_out.data_ptr(), d_enc_b.data_ptr(),
                                             gw_b.data_ptr(), gw_b.data_ptr(), C.byref(d), coords.data_ptr(), gt_b.data_ptr(),
                                             n, 256, a, _s()) == nat.ERR_UNSUPPORTED
