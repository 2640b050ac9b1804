"""Hash-grid level table and parameter layout of the tcnn-style INRs.

Follows the configuration the reference hands to tiny-cuda-nn (src/models/immoco.py:11-37) and
tiny-cuda-nn's published Grid/Hash encoding rules (grid.h) as fixed by SURVEY.md Appendix B.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Tuple

import numpy as np

from . import _native as nat

OUT_PAD = 16          # tiny-cuda-nn pads the output layer to 16 rows
N_ENCODED = 32        # kernels are specialised for 16 levels x 2 features

_KNOWN_ENC_KEYS = {"otype", "type", "n_levels", "n_features_per_level", "log2_hashmap_size",
                   "base_resolution", "per_level_scale", "interpolation", "fine_resolution", "stride_wrap"}

# Process-wide default of the ``stride_wrap`` encoding option (see grid_spec); the config key wins.
_STRIDE_WRAP_DEFAULT = False


def set_stride_wrap_default(on: bool) -> None:
    """tiny-cuda-nn compatibility switch for every grid built afterwards without an explicit
    ``"stride_wrap"`` key in its encoding config (see ``grid_spec``)."""
    global _STRIDE_WRAP_DEFAULT
    _STRIDE_WRAP_DEFAULT = bool(on)
_ACTS = {"none": nat.ACT_NONE, "relu": nat.ACT_RELU, "tanh": nat.ACT_TANH}


@dataclass(frozen=True)
class GridSpec:
    n_dims: int
    n_levels: int
    scales: Tuple[float, ...]
    resolutions: Tuple[int, ...]
    entries: Tuple[int, ...]
    offsets: Tuple[int, ...]
    hashed: Tuple[int, ...]

    @property
    def n_rows(self) -> int:
        return self.offsets[-1]

    @property
    def n_table_params(self) -> int:
        return self.offsets[-1] * 2

    def desc(self, swizzle: Tuple[int, ...] = (), lut_ptr: int = 0) -> nat.GridDesc:
        """C-ABI descriptor; ``swizzle`` (one word per level, see ``row_swizzle``) selects the physical row
        layout of the hashed levels -- empty = the reference's own layout.  A word ``nat.LAYOUT_LUT`` selects the
        level's chunk tables of ``linear_layout`` at device address ``lut_ptr`` ((n_levels, 256) uint32)."""
        d = nat.GridDesc()
        d.n_dims = self.n_dims
        d.n_levels = self.n_levels
        for i in range(self.n_levels):
            d.scale[i] = self.scales[i]
            d.resolution[i] = self.resolutions[i]
            d.entries[i] = self.entries[i]
            d.hashed[i] = self.hashed[i]
            d.swizzle[i] = int(swizzle[i]) if swizzle else 0
        for i in range(self.n_levels + 1):
            d.offset[i] = self.offsets[i]
        d.layout_lut = int(lut_ptr) if lut_ptr else None
        if any(int(w) == nat.LAYOUT_LUT for w in swizzle) and not lut_ptr:
            raise ValueError("a LAYOUT_LUT level needs the device address of the layout tables")
        return d

    def row_swizzle(self, dim0_values) -> Tuple[int, ...]:
        """Row layout of the hashed levels for inputs whose FIRST coordinate only takes the few values
        ``dim0_values`` (the Motion INR: one value per movement group, src/models/immoco.py:48-53).

        The two dim-0 corners of a cell are the hash indices r and r ^ X, X = cell ^ (cell + 1) = 2^t - 1.
        For t <= 4 both rows share a 128-byte line; the group at coordinate +1 has cell = res - 1, i.e.
        t = log2(res) + 1, and its two corners are far apart.  The layout word asks the kernels to store
        row r at S(r) = swap_{a,b}(r ^ (r >> 1)): the Gray code maps every X = 2^t - 1 to the single bit
        t - 1, and the swap moves the large t - 1 (a) to a free low position (b), so all lane pairs of the
        level read / reduce two rows of ONE line.  0 = keep the reference's layout (nothing to gain)."""
        words = []
        for lvl in range(self.n_levels):
            n = self.entries[lvl]
            word = 0
            if self.hashed[lvl] and (n & (n - 1)) == 0 and len(dim0_values) > 0:
                tops = []
                for u in dim0_values:
                    pos = np.float32(np.float64(np.float32(self.scales[lvl])) * np.float64(np.float32(u)) + 0.5)
                    cell = int(np.floor(pos)) & 0xFFFFFFFF
                    x = (cell ^ ((cell + 1) & 0xFFFFFFFF)) & (n - 1)
                    tops.append(x.bit_length() - 1)
                small = {t for t in tops if t < 4}
                big = sorted({t for t in tops if t >= 4}, key=lambda t: -tops.count(t))
                free = [b for b in range(4) if b not in small]
                if big and free:
                    word = (1 << 31) | (free[0] << 8) | big[0]
            words.append(word)
        return tuple(words)

    def linear_layout(self, dim0_values) -> np.ndarray:
        """General linear row layout of the hashed levels for a grid whose FIRST coordinate only takes the few
        values ``dim0_values`` (one per movement group): a (n_levels, 256) uint32 lookup table, zero rows = keep
        the reference layout.

        For one (dim-1.., corner) hash H the rows a pixel needs over all groups and both dim-0 corners are
        H ^ d, d in D = {cell_g, cell_g + 1}: a 'bundle' of 2 M rows whose mutual differences span a small
        subspace V of the index bits (dimension k <= ~M + 1: the pair masks cell ^ (cell + 1) and the cross-group
        differences).  A linear bijection S that sends a basis of V to the unit vectors e_0 .. e_{k-1} puts every
        bundle into one aligned block of 2^k rows -- with the most frequent pair mask on e_0 (both rows of those
        pairs in one 16-byte slot, one RED.ADD.F32x4) and the basis ordered so that the bundle covers as few
        128-byte lines (16 rows) as possible: 2 instead of 4 lines at M = 4, 1 instead of 2 at M = 2.  ``row_swizzle``
        is the special case (Gray code + one exchange) that only aligns the pairs.  S is handed to the kernels as
        three chunk tables: S(x) = T0[x & 127] ^ T1[(x >> 7) & 63] ^ T2[(x >> 13) & 63] (linearity)."""
        lut = np.zeros((self.n_levels, 256), dtype=np.uint32)
        vals = [float(v) for v in dim0_values]
        if len(vals) < 2:
            return lut
        for lvl in range(self.n_levels):
            n = self.entries[lvl]
            bits = n.bit_length() - 1
            if not self.hashed[lvl] or (n & (n - 1)) != 0 or bits > 19 or bits < 8:
                continue
            cells = []
            for u in vals:
                pos = np.float32(np.float64(np.float32(self.scales[lvl])) * np.float64(np.float32(u)) + 0.5)
                cells.append(int(np.floor(pos)) & 0xFFFFFFFF)
            d = [((c & (n - 1)), ((c + 1) & 0xFFFFFFFF) & (n - 1)) for c in cells]
            masks = [a ^ b for a, b in d]
            cross = [d[g][0] ^ d[0][0] for g in range(1, len(d))]
            by_freq = sorted(set(masks), key=lambda m: (-masks.count(m), m))
            best = None
            for order in (by_freq + cross, [by_freq[0]] + cross + by_freq[1:], cross + by_freq):
                basis = _gf2_independent([v for v in order if v])
                mat = _gf2_layout(basis, bits)
                img = lambda x: _gf2_apply(mat, x)          # noqa: E731
                lines = {img(dd ^ d[0][0]) >> 4 for pair in d for dd in pair}
                merged = sum(1 for m in masks if img(m) == 1)
                key = (-merged, len(lines))       # the scatter is paced by the reduction count, the gather by lines
                if best is None or key < best[0]:
                    best = (key, mat)
            mat = best[1]
            for x in range(128):
                lut[lvl, x] = _gf2_apply(mat, x)
            for x in range(64):
                lut[lvl, 128 + x] = _gf2_apply(mat, x << 7)
                lut[lvl, 192 + x] = _gf2_apply(mat, x << 13)
        return lut

    def row_permutation_lut(self, lut: np.ndarray) -> np.ndarray:
        """perm[r] = physical row of logical (reference-layout) row r under ``linear_layout``'s tables."""
        perm = np.arange(self.n_rows, dtype=np.int64)
        for lvl in range(self.n_levels):
            if not lut[lvl].any():
                continue
            r = np.arange(self.entries[lvl], dtype=np.int64)
            t = lut[lvl].astype(np.int64)
            phys = t[r & 127] ^ t[128 + ((r >> 7) & 63)] ^ t[192 + ((r >> 13) & 63)]
            assert np.array_equal(np.sort(phys), r), "layout is not a bijection"
            perm[self.offsets[lvl]: self.offsets[lvl + 1]] = self.offsets[lvl] + phys
        return perm

    def row_permutation(self, swizzle: Tuple[int, ...]) -> np.ndarray:
        """perm[r] = physical row of logical (reference-layout) row r, over the whole table."""
        perm = np.arange(self.n_rows, dtype=np.int64)
        for lvl in range(self.n_levels):
            w = int(swizzle[lvl]) if swizzle else 0
            if w == 0 or not self.hashed[lvl] or (self.entries[lvl] & (self.entries[lvl] - 1)) != 0:
                continue
            r = np.arange(self.entries[lvl], dtype=np.int64)
            r ^= r >> 1
            a, b = w & 0xFF, (w >> 8) & 0xFF
            x = ((r >> a) ^ (r >> b)) & 1
            r ^= (x << a) | (x << b)
            perm[self.offsets[lvl]: self.offsets[lvl + 1]] = self.offsets[lvl] + r
        return perm



def _gf2_independent(vectors):
    """The vectors (ints = bit vectors) that are linearly independent of their predecessors, in order."""
    reduced, keep = [], []
    for v in vectors:
        x = v
        for b in reduced:
            x = min(x, x ^ b)
        if x:
            reduced.append(x)
            reduced.sort(reverse=True)
            keep.append(v)
    return keep


def _gf2_layout(basis, bits: int):
    """Columns S(e_j), j < bits, of the linear bijection with S(basis[i]) = e_i; the basis is completed with unit
    vectors (lowest bits first), which go to the remaining unit vectors in order."""
    cols = list(basis)
    for j in range(bits):
        if len(_gf2_independent(cols + [1 << j])) > len(cols):
            cols.append(1 << j)
    assert len(cols) == bits
    # invert B (columns `cols`: B e_i = cols[i]) by Gauss-Jordan on [B | I], rows as ints
    rows = [sum(((cols[c] >> r) & 1) << c for c in range(bits)) | (1 << (bits + r)) for r in range(bits)]
    for c in range(bits):
        piv = next(r for r in range(c, bits) if (rows[r] >> c) & 1)
        rows[c], rows[piv] = rows[piv], rows[c]
        for r in range(bits):
            if r != c and (rows[r] >> c) & 1:
                rows[r] ^= rows[c]
    inv_rows = [rw >> bits for rw in rows]                 # row r of S = B^-1
    return [sum(((inv_rows[r] >> j) & 1) << r for r in range(bits)) for j in range(bits)]      # columns of S


def _gf2_apply(cols, x: int) -> int:
    out, j = 0, 0
    while x:
        if x & 1:
            out ^= cols[j]
        x >>= 1
        j += 1
    return out


def grid_spec(n_dims: int, cfg: dict) -> GridSpec:
    """Level scales / resolutions / entry counts / offsets of a Grid-Hash encoding.

    ``cfg["stride_wrap"]`` (default False = SURVEY.md section 8's contract): how grid_index()'s dense-index
    stride is kept while the level kinds are decided.  False: as a mathematical integer -- with the reference
    configuration levels 6-15 (2-D) / 3-15 (3-D) are hashed.  True: in a uint32 like tiny-cuda-nn's
    ``grid_index`` keeps it -- for resolutions >= 2^16 (levels 12-15 here) ``stride * res`` wraps to 0, the
    hash branch is skipped and those levels index densely with the wrapped strides,
    (q0 + q1 * res) mod 2^19, the third coordinate dropping out.  The kernels take the per-level kind from the
    host (``immoco_grid_desc::hashed``) and their dense index arithmetic is uint32 anyway, so both modes run
    the same code (tests/test_gpu_kernels.py::test_hashgrid_stride_wrap_mode)."""
    if n_dims not in (2, 3):
        raise NotImplementedError("the IM-MoCo path uses 2-D and 3-D hash grids only")
    if str(cfg.get("otype", "Grid")) not in ("Grid", "HashGrid") or str(cfg.get("type", "Hash")) != "Hash":
        raise NotImplementedError("only the Grid/Hash encoding is implemented")
    if str(cfg.get("interpolation", "Linear")) != "Linear":
        raise NotImplementedError("only Linear interpolation is implemented")
    n_levels = int(cfg.get("n_levels", 16))
    n_feat = int(cfg.get("n_features_per_level", 2))
    if n_levels != 16 or n_feat != 2:
        raise NotImplementedError("kernels are specialised for n_levels=16, n_features_per_level=2")
    log2_t = int(cfg.get("log2_hashmap_size", 19))
    base = int(cfg.get("base_resolution", 16))
    log2_pls = math.log2(float(cfg.get("per_level_scale", 2.0)))
    wrap = bool(cfg.get("stride_wrap", _STRIDE_WRAP_DEFAULT))
    scales, ress, ents, offs, hashed = [], [], [], [0], []
    for lvl in range(n_levels):
        scale = float(np.float32(np.exp2(np.float32(lvl * log2_pls)) * np.float32(base) - np.float32(1.0)))
        res = int(math.ceil(scale)) + 1
        dense = res ** n_dims
        cap = 0xFFFFFFFF // 2
        n = cap if float(dense) > float(cap) else dense
        n = min((n + 7) // 8 * 8, 1 << log2_t)
        # hash iff the dense index range walked by grid_index() exceeds the level's entry count
        stride, dim = 1, 0
        while dim < n_dims and stride <= n:
            stride *= res
            if wrap:
                stride &= 0xFFFFFFFF
            dim += 1
        scales.append(scale)
        ress.append(res)
        ents.append(n)
        offs.append(offs[-1] + n)
        hashed.append(1 if n < stride else 0)
    if offs[-1] >= 2 ** 31:
        raise NotImplementedError("table too large")
    return GridSpec(n_dims, n_levels, tuple(scales), tuple(ress), tuple(ents), tuple(offs), tuple(hashed))


@dataclass(frozen=True)
class MlpSpec:
    width: int
    act: int

    @property
    def n_w1(self) -> int:
        return self.width * N_ENCODED

    @property
    def n_params(self) -> int:
        return self.width * N_ENCODED + OUT_PAD * self.width


def mlp_spec(cfg: dict) -> MlpSpec:
    """Accepts both otype strings the reference uses (CutLassMLP / FullyFusedMLP)."""
    otype = str(cfg.get("otype", "FullyFusedMLP")).lower()
    if otype not in ("cutlassmlp", "fullyfusedmlp"):
        raise NotImplementedError(f"network otype {cfg.get('otype')!r} is not on the IM-MoCo path")
    width = int(cfg.get("n_neurons", 64))
    if width not in (64, 256):
        raise NotImplementedError("n_neurons must be 64 or 256")
    if int(cfg.get("n_hidden_layers", 1)) != 1:
        raise NotImplementedError("n_hidden_layers must be 1")
    act = str(cfg.get("activation", "ReLU")).lower()
    if act not in ("relu", "tanh"):
        raise NotImplementedError("activation must be ReLU or Tanh")
    if str(cfg.get("output_activation", "None")).lower() != "none":
        raise NotImplementedError("output_activation must be None")
    return MlpSpec(width, _ACTS[act])


def twiddles(n: int) -> np.ndarray:
    """exp(-2 pi i t / n), t = 0..n-1, evaluated in float64 and rounded once -> (n, 2) float32."""
    t = np.arange(n, dtype=np.float64)
    ang = -2.0 * np.pi * t / n
    return np.stack([np.cos(ang), np.sin(ang)], axis=1).astype(np.float32)
