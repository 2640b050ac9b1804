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
    assert native_lib.immoco_abi_version() == 3
    assert native_lib.immoco_launches_per_iteration(4) == 16 and len(nat.PROFILE_SLOTS) == 16
    assert native_lib.immoco_launches_per_iteration_mode(4, 1, 0) == 17
    assert native_lib.immoco_launches_per_iteration_mode(4, 1, 1) == 15
    assert native_lib.immoco_launches_per_iteration_mode(0, 1, 1) == 10
    sizes = (C.c_int32 * 5)()
    native_lib.immoco_struct_sizes(sizes)
    assert tuple(sizes) == (C.sizeof(nat.GridDesc), C.sizeof(nat.Lines), C.sizeof(nat.Fit), C.sizeof(nat.GridCsr),
                            C.sizeof(nat.GridTaps))
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


def test_stride_wrap_level_kinds_match_oracle_in_both_modes():
    """encoding_config["stride_wrap"]: the package's grid_spec and the oracle agree on which levels hash, in the
    default (SURVEY contract: levels 6-15 / 3-15 hashed) and in the tiny-cuda-nn uint32-stride mode (levels 12-15
    index densely); the process-wide default switch is honoured and restored."""
    from miccai24_immoco_b200 import encoding as enc
    from oracle import immoco_oracle as orc
    for dims, first_hashed in ((2, 6), (3, 3)):
        for wrap in (False, True):
            gs = enc.grid_spec(dims, dict(mb.encoding_config, stride_wrap=wrap))
            lv = orc.make_grid_levels(dims, dict(orc.ENCODING_CONFIG, stride_wrap=wrap))
            assert gs.hashed == tuple(int(v) for v in lv.hashed)
            want = [int(first_hashed <= l < (12 if wrap else 16)) for l in range(16)]
            assert list(gs.hashed) == want
            assert gs.offsets == lv.offsets and gs.entries == lv.entries
    try:
        enc.set_stride_wrap_default(True)
        assert enc.grid_spec(3, mb.encoding_config).hashed[12:] == (0, 0, 0, 0)
        assert enc.grid_spec(3, dict(mb.encoding_config, stride_wrap=False)).hashed[12:] == (1, 1, 1, 1)   # the key wins
    finally:
        enc.set_stride_wrap_default(False)
    assert enc.grid_spec(3, mb.encoding_config).hashed[12:] == (1, 1, 1, 1)
    # wrapped dense index of the oracle: third coordinate drops out at level 15 (res 2^19)
    import torch
    q = [torch.tensor([5, 5]), torch.tensor([7, 7]), torch.tensor([1, 900])]
    idx = orc.grid_corner_index(q, 1 << 19, 1 << 19, stride_wrap=True)
    assert int(idx[0]) == int(idx[1]) == 5


def test_batched_movement_group_labels_match_per_slice_extraction():
    """movement_masks_from_kspace labels a whole stack in one pass: (B, W) labels == the per-slice
    extract_movement_groups of the reference interface, label count == largest label."""
    import torch
    from miccai24_immoco_b200.motion_utils import movement_group_labels
    from oracle import immoco_oracle as orc
    g = torch.Generator().manual_seed(3)
    flags = torch.rand(9, 57, generator=g) > 0.6
    flags[0] = False
    flags[1] = True
    labels = movement_group_labels(flags)
    for b in range(flags.shape[0]):
        want = orc.extract_movement_groups(flags[b], make_list=True, height=4)
        assert int(labels[b].max()) == want.shape[0]
        got = (labels[b].view(1, 1, -1).expand(1, 4, -1) == torch.arange(1, want.shape[0] + 1).view(-1, 1, 1)).long()
        assert torch.equal(got, want)
        assert torch.equal(mb.extract_movement_groups(flags[b], make_list=True, height=4), want)


def test_bench_arms_share_one_config_and_every_baseline_config_is_selectable():
    """bench.py: the GPU arm and the --impl reference arm print the SAME config dict for the driver's comparison;
    every configuration BASELINE.json names has a --config entry whose workload string names its shape and n_M."""
    import json
    import bench
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(bench.CONFIGS) == len(base["configs"]) == 5
    assert sorted(c["index"] for c in bench.CONFIGS.values()) == [0, 1, 2, 3, 4]
    for name, c in bench.CONFIGS.items():
        a = bench.config_dict(name, c["n_mov"], 1000, 2)
        b = bench.config_dict(name, c["n_mov"], 1000, 2)
        assert a == b and f"{c['h']}x{c['w']}" in a["workload"] and a["baseline_config"] == c["index"]
    assert "n_M=4" in bench.workload_string("c2", 4, 1000) and "n_M=5" in bench.workload_string("c3", 5, 1000)
    assert bench.CONFIGS["c3"]["slices"] == 16 and bench.CONFIGS["c4"]["slices"] == 64 and bench.CONFIGS["c5"]["slices"] == 256
    assert bench.METRIC == base["metric"].split(" at ")[0]


def test_run_batched_and_deterministic_flags_are_part_of_the_public_surface():
    import inspect
    assert "deterministic" in inspect.signature(mb.imcoco_motion_correction).parameters
    assert {"deterministic", "batch"} <= set(inspect.signature(mb.reconstruct_batch).parameters)
    assert {"deterministic", "batch"} <= set(inspect.signature(mb.reconstruct_slices).parameters)
    assert callable(mb.run_batched) and mb.lib().immoco_max_fit_batch() == 8
    # positional signature of the reference call is untouched (immoco.py:116)
    names = list(inspect.signature(mb.imcoco_motion_correction).parameters)[:6]
    assert names == ["kspace_corr", "masks", "iters", "learning_rate", "lambda_ge", "debug"]


@pytest.mark.parametrize("m", [2, 4, 5, 8, 16])
def test_linear_row_layout_is_a_bijection_that_packs_the_bundles(m):
    """encoding.py:GridSpec.linear_layout: per hashed level a linear bijection of the index bits (chunk tables) that puts
    the 2 M rows one pixel corner needs over all groups into fewer 128-byte lines than one line per group, with the most
    frequent dim-0 pair in one 16-byte slot."""
    import miccai24_immoco_b200 as mb
    from miccai24_immoco_b200.encoding import grid_spec
    gs = grid_spec(3, mb.encoding_config)
    u = np.linspace(-1, 1, m)
    lut = gs.linear_layout(u)
    perm = gs.row_permutation_lut(lut)
    assert np.array_equal(np.sort(perm), np.arange(gs.n_rows))
    lines_total, pairs_in_slot = 0, 0
    for lvl in range(gs.n_levels):
        if not gs.hashed[lvl]:
            assert not lut[lvl].any() and np.array_equal(perm[gs.offsets[lvl]:gs.offsets[lvl + 1]],
                                                          np.arange(gs.offsets[lvl], gs.offsets[lvl + 1]))
            continue
        n = gs.entries[lvl]
        t = lut[lvl].astype(np.int64)
        s_of = lambda x: int(t[x & 127] ^ t[128 + ((x >> 7) & 63)] ^ t[192 + ((x >> 13) & 63)])   # noqa: E731
        rng = np.random.default_rng(lvl)
        for a, b in rng.integers(0, n, size=(50, 2)):
            assert s_of(int(a) ^ int(b)) == s_of(int(a)) ^ s_of(int(b))            # linear
        cells = [int(np.floor(np.float32(np.float64(np.float32(gs.scales[lvl])) * np.float64(np.float32(v)) + 0.5)))
                 & 0xFFFFFFFF for v in u]
        d = [(c & (n - 1), (c + 1) & (n - 1)) for c in cells]
        lines = {s_of(x ^ d[0][0]) >> 4 for pair in d for x in pair}
        assert len(lines) <= max(1, int(np.ceil(0.75 * m)))
        lines_total += len(lines)
        pairs_in_slot += sum(1 for a, b in d if s_of(a ^ b) == 1)
    n_hashed = sum(gs.hashed)
    assert lines_total <= 0.6 * m * n_hashed
    assert pairs_in_slot >= n_hashed * (m // 2) * 0.9 or m == 16
    assert not gs.linear_layout([0.0]).any()           # a single group: nothing to pack


def test_tap_indexed_storage_is_chosen_where_it_pays():
    """immoco.py:_taps_pay -- the image table is stored tap-indexed when at most ~70 % of a hashed level's rows are
    expected to be touched (measured: 320 x 320 gains 6 %, 640 x 368 loses 1.5 %); 3-D grids never are."""
    import miccai24_immoco_b200 as mb
    from miccai24_immoco_b200.encoding import grid_spec
    from miccai24_immoco_b200.immoco import _taps_pay, _taps_supported
    g2, g3 = grid_spec(2, mb.encoding_config), grid_spec(3, mb.encoding_config)
    assert _taps_supported(g2) and not _taps_supported(g3)
    assert _taps_pay(g2, 320 * 320) and _taps_pay(g2, 48 * 40) and not _taps_pay(g2, 640 * 368)
    assert not _taps_pay(g3, 320 * 320)
    # expected touched fraction at the switch-over: 1 - exp(-1.2) = 0.70
    assert abs((1 - np.exp(-4 * 320 * 320 / 2 ** 19)) - 0.542) < 1e-3
