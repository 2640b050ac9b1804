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

    def desc(self, swizzle: Tuple[int, ...] = ()) -> nat.GridDesc:
        """C-ABI descriptor; ``swizzle`` (one word per level, see ``row_swizzle``) selects the physical row
        layout of the hashed levels -- empty = the reference's own layout."""
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
