"""CPU: host-side logic of the product package and the C-ABI surface (no compute calls)."""
import ctypes as C
import os
import re
import shutil
import subprocess

import numpy as np
import pytest
import torch

import miccai24_immoco_b200 as mb
from miccai24_immoco_b200 import _native as nat
from miccai24_immoco_b200.encoding import grid_spec, mlp_spec, twiddles
from oracle import immoco_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol(native_lib):
    header = open(os.path.join(ROOT, "include", "immoco_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|void|immoco_profile\*)\s+(immoco_\w+)\s*\(", header, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(nat.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(native_lib, name) is not None
    assert native_lib.immoco_abi_version() == 2
    assert native_lib.immoco_launches_per_iteration(4) == 16 and len(nat.PROFILE_SLOTS) == 16
    assert native_lib.immoco_launches_per_iteration_mode(4, 1, 0) == 17
    assert native_lib.immoco_launches_per_iteration_mode(4, 1, 1) == 15
    assert native_lib.immoco_launches_per_iteration_mode(0, 1, 1) == 10
    sizes = (C.c_int32 * 4)()
    native_lib.immoco_struct_sizes(sizes)
    assert tuple(sizes) == (C.sizeof(nat.GridDesc), C.sizeof(nat.Lines), C.sizeof(nat.Fit), C.sizeof(nat.GridCsr))
    # the fp32 SIMT MLP kernels are test-side only (tests/checkers), not in the product library
    assert not hasattr(native_lib, "immoco_set_mlp_impl") and "mlp.cu" not in nat.SOURCES
    assert native_lib.immoco_get_deterministic() in (0, 1)


def test_library_targets_sm100a():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    mb.build()
    out = subprocess.run([cuobjdump, "-lelf", nat.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_grid_spec_equals_oracle_levels():
    for d in (2, 3):
        gs = grid_spec(d, mb.encoding_config)
        lv = orc.make_grid_levels(d, orc.ENCODING_CONFIG)
        assert gs.offsets == lv.offsets and gs.entries == lv.entries
        assert gs.resolutions == lv.resolutions and gs.scales == lv.scales
        assert tuple(bool(h) for h in gs.hashed) == lv.hashed


def test_configs_equal_oracle_configs():
    assert mb.network_config == orc.IMAGE_NETWORK_CONFIG
    assert mb.mot_network_config == orc.MOTION_NETWORK_CONFIG
    assert mb.encoding_config == orc.ENCODING_CONFIG
    assert mlp_spec(mb.network_config).n_params == 12288
    assert mlp_spec(mb.mot_network_config).n_params == 3072
    # both otype spellings of the reference are accepted
    assert mlp_spec({**mb.network_config, "otype": "FullyFusedMLP", "dtype": "float32"}).width == 256
    with pytest.raises(NotImplementedError):
        mlp_spec({**mb.network_config, "n_neurons": 128})


def test_lambda_schedule_equals_oracle():
    for iters in (10, 50, 200, 1000):
        assert mb.lambda_schedule(iters, 1e-2) == orc.lambda_schedule(iters, 1e-2)
        assert mb.lambda_schedule(iters, 1e-2, "downstream") == orc.lambda_schedule(iters, 1e-2, "downstream")
    with pytest.raises(ZeroDivisionError):
        mb.lambda_schedule(5, 1e-2)


@pytest.mark.parametrize("w", [1, 2, 7, 16, 64])
def test_extract_movement_groups_equals_oracle_random(w):
    g = torch.Generator().manual_seed(w)
    for trial in range(20):
        lines = torch.rand(w, generator=g) > 0.5
        for ml in (False, True):
            a = mb.extract_movement_groups(lines, make_list=ml)
            b = orc.extract_movement_groups(lines, make_list=ml)
            assert a.shape == b.shape and torch.equal(a, b)
        a = mb.extract_movement_groups(lines, make_list=True, height=2 * w + 1)
        b = orc.extract_movement_groups(lines, make_list=True, height=2 * w + 1)
        assert a.shape == b.shape and torch.equal(a, b)


def test_extract_movement_groups_golden(golden_dir):
    ops = np.load(os.path.join(golden_dir, "ops_small.npz"))
    for name in ("empty", "all", "runs", "edges", "single_last"):
        lines = torch.from_numpy(ops[f"groups_{name}_in"])
        for ml in (0, 1):
            got = mb.extract_movement_groups(lines, make_list=bool(ml))
            want = torch.from_numpy(ops[f"groups_{name}_{ml}"])
            assert got.shape == want.shape and torch.equal(got, want)


def test_make_grids_and_identity_equal_oracle():
    assert torch.equal(mb.make_grids((3, 5, 7)), orc.make_grids((3, 5, 7)))
    assert torch.equal(mb.make_grids((1, 4, 4)), orc.make_grids((1, 4, 4)))   # M == 1 -> m = -1
    from miccai24_immoco_b200.immoco import _identity_grid
    assert torch.equal(_identity_grid(6, 10, "cpu"), orc.identity_grid(6, 10))


def test_line_structure():
    lines = torch.tensor([0, 1, 1, 0, 0, 1, 0, 0], dtype=torch.bool)
    masks = mb.extract_movement_groups(lines, make_list=True, height=6)
    ls = mb.LineStructure(masks)
    assert (ls.m, ls.h, ls.w) == (2, 6, 8)
    assert ls.group_ofs.tolist() == [0, 2, 3] and ls.line_idx.tolist() == [1, 2, 5]
    assert ls.static_w.tolist() == [1, 0, 0, 1, 1, 0, 1, 1] and ls.max_lines == 2
    bad = masks.clone()
    bad[0, 0, 1] = 0
    with pytest.raises(NotImplementedError):
        mb.LineStructure(bad)
    empty = mb.LineStructure(torch.zeros((0, 6, 8), dtype=torch.long))
    assert empty.m == 0 and empty.static_w.tolist() == [1.0] * 8


def test_twiddles():
    t = twiddles(8)
    assert t.shape == (8, 2) and np.allclose(t[2], [0, -1], atol=1e-7)


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError):
        mb.FFT(torch.zeros(4, 4, dtype=torch.complex64))
    with pytest.raises(RuntimeError):
        mb.GradientEntropyLoss()(torch.zeros(4, 4, dtype=torch.complex64))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            mb.NetworkWithInputEncoding(2, 2, mb.encoding_config, mb.network_config)
        with pytest.raises(RuntimeError):
            mb.imcoco_motion_correction(torch.zeros(8, 8, dtype=torch.complex64),
                                        torch.zeros((1, 8, 8), dtype=torch.long), iters=10)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "miccai24_immoco_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("oracle/", "").lower() or "import oracle" not in src, fn
            assert "from oracle" not in src and "import oracle" not in src, fn


# ---- host replay of the kernels' __host__ __device__ maths (compiled with nvcc, run on CPU) ------
@pytest.fixture(scope="module")
def hostcheck():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    d = os.path.join(ROOT, "tests", "hostcheck")
    so = os.path.join(d, "_hostcheck.so")
    src = os.path.join(d, "hostcheck.cu")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC",
                        "-shared", "-I", os.path.join(ROOT, "include"), "-o", so, src], check=True)
    return C.CDLL(so)


@pytest.mark.parametrize("n", [320, 640, 368, 64, 46, 20, 2, 30])
def test_fft_butterflies_on_host(hostcheck, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))).astype(np.complex64)
    out = np.empty_like(x)
    for inv in (0, 1):
        assert hostcheck.hostcheck_fft(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), n, 3, inv) > 0
        x64 = x.astype(np.complex128)
        ref = np.fft.ifft(x64, axis=1) * n if inv else np.fft.fft(x64, axis=1)
        assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 1e-6


@pytest.mark.parametrize("dims", [2, 3])
def test_hash_index_maths_bit_exact_vs_oracle(hostcheck, dims):
    gs = grid_spec(dims, mb.encoding_config)
    lv = orc.make_grid_levels(dims, orc.ENCODING_CONFIG)
    d = gs.desc()
    coords = (orc.identity_grid(40, 24).view(-1, 2) if dims == 2 else orc.make_grids((3, 12, 10))).contiguous()
    n, cn = coords.shape[0], coords.numpy()
    for level in range(16):
        idx = np.empty((1 << dims, n), np.uint32)
        w = np.empty((1 << dims, n), np.float32)
        hostcheck.hostcheck_taps(C.byref(d), level, cn.ctypes.data_as(C.c_void_p), n,
                                 idx.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p))
        oi, ow = orc.hashgrid_taps(coords, lv, level)
        assert np.array_equal(idx.astype(np.int64), oi.numpy()), level
        assert np.array_equal(w, ow.numpy()), level


def test_row_swizzle_is_a_storage_permutation(hostcheck):
    """The physical row layout of the motion grid (GridSpec.row_swizzle / immoco_grid_desc::swizzle): the
    kernels' index routine under the layout word equals the oracle's index pushed through the host-side
    permutation, the permutation is a bijection per level, and both dim-0 corners of every tap pair end up in
    one 128-byte line (16 rows) -- the reason for the layout."""
    gs = grid_spec(3, mb.encoding_config)
    lv = orc.make_grid_levels(3, orc.ENCODING_CONFIG)
    for m in (2, 4, 5, 8):
        coords = orc.make_grids((m, 6, 5)).contiguous()
        u = np.unique(coords[:, 0].numpy())
        swz = gs.row_swizzle(u)
        assert all(w == 0 for lvl, w in enumerate(swz) if not gs.hashed[lvl]) and any(swz)
        perm = gs.row_permutation(swz)
        assert np.array_equal(np.sort(perm), np.arange(gs.n_rows))
        d = gs.desc(swz)
        n, cn = coords.shape[0], coords.numpy()
        for level in range(16):
            idx = np.empty((8, n), np.uint32)
            w = np.empty((8, n), np.float32)
            hostcheck.hostcheck_taps(C.byref(d), level, cn.ctypes.data_as(C.c_void_p), n,
                                     idx.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p))
            oi, _ = orc.hashgrid_taps(coords, lv, level)
            off = gs.offsets[level]
            want = perm[off + oi.numpy()] - off
            assert np.array_equal(idx.astype(np.int64), want), (m, level)
            if gs.hashed[level]:
                # corner c and c ^ 1 differ in the dim-0 bit: same 16-row line after the permutation
                assert np.array_equal(idx[0::2] >> 4, idx[1::2] >> 4), (m, level)
