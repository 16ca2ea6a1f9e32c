"""Launch planner: region geometry and temporal-blocking depth for the sm_100a kernels.

A CTA region is ``RH = 16*R`` rows x 512 bytes held in registers (R = band height per warp), of
which the interior ``TH x TWB`` (``TH <= RH - 2k``, ``TWB = 512 - 2*HB``, ``HB = roundup16(k*C)``)
is written back after ``k`` fused iterations. Deeper ``k`` divides the number of launches /
cross-GPU synchronisations by ``k`` (every BASELINE config is latency- or ALU-bound, never
link-bound: SURVEY.md §7.2) but recomputes a thicker halo; the planner minimises a simple time
model over ``(k, R, TH)`` including the wave quantisation over the 148 SMs.

The plan must be identical on every rank, so it only depends on quantities every rank knows
(the largest/smallest tile of the grid, the filter path, the SM count).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

from ..ops.host import KIND_F32, KIND_INT, KIND_SEP121

REGION_BYTES = 512          # kSP in csrc/cuda/stencil_core.cuh
WARPS = 16
DEFAULT_SMS = 148
# Band heights compiled into libpconv_cuda.so (csrc/cuda/stream_kernel.cu); refined from the
# library at runtime (ops.cuda.band_heights).
DEFAULT_BANDS = {KIND_SEP121: [2, 4, 7, 10, 14], KIND_INT: [2, 4, 7, 10], KIND_F32: [2, 4, 6, 8, 10]}

# Modelled SM throughput in region bytes per clock (from B200 measurements, profiles/): used only
# to rank candidate plans, never reported as a result.
BYTES_PER_CLK = {KIND_SEP121: 20.0, KIND_INT: 9.0, KIND_F32: 4.0}
CLOCK_GHZ = 1.8
LAUNCH_US = 4.0             # dependent-launch gap incl. the kernel prologue/epilogue
LOAD_BYTES_PER_US_PER_SM = 40e3  # ~6 TB/s / 148 SMs


@dataclass(frozen=True)
class StreamPlan:
    k: int            # fused iterations per launch
    R: int            # band height (rows per warp); region rows = 16*R
    TH: int           # interior rows per region
    TWB: int          # interior bytes per region
    HB: int           # column halo bytes
    ntx: int
    nty: int
    waves: int
    est_us_per_launch: float

    @property
    def RH(self) -> int:
        return WARPS * self.R

    def launches(self, loops: int) -> list[int]:
        """k of every launch for ``loops`` iterations (the last one may be shallower)."""
        full, rem = divmod(loops, self.k)
        return [self.k] * full + ([rem] if rem else [])


def _roundup(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def region_geometry(k: int, C: int) -> tuple[int, int]:
    HB = _roundup(max(k * C, 1), 16)
    return HB, REGION_BYTES - 2 * HB


def plan_stream(tile_rows: int, tile_cb: int, C: int, kind: int, loops: int, *,
                min_nbr_rows: int | None = None, min_nbr_cols: int | None = None,
                sms: int = DEFAULT_SMS, bands: list[int] | None = None,
                force_k: int | None = None, force_R: int | None = None) -> StreamPlan:
    """Pick (k, R, TH) for a tile of ``tile_rows`` x ``tile_cb`` bytes.

    ``min_nbr_rows`` / ``min_nbr_cols`` (pixels): smallest neighbouring tile extent — the halo may
    only reach into adjacent tiles, so ``k`` cannot exceed them.
    """
    loops = max(loops, 1)
    bands = sorted(bands or DEFAULT_BANDS[kind])
    if force_R:
        bands = [force_R]
    rate = BYTES_PER_CLK[kind] * CLOCK_GHZ * 1e3  # region bytes per us per SM
    kmax = min(loops, (WARPS * bands[-1] - 1) // 2, (REGION_BYTES // 2 - 64) // C)
    if min_nbr_rows is not None:
        kmax = min(kmax, min_nbr_rows)
    if min_nbr_cols is not None:
        kmax = min(kmax, min_nbr_cols)
    kmax = max(kmax, 1)
    ks = [force_k] if force_k else range(1, kmax + 1)
    best: tuple[float, StreamPlan] | None = None
    for k in ks:
        HB, TWB = region_geometry(k, C)
        if TWB < 64:
            continue
        ntx = math.ceil(tile_cb / TWB)
        for R in bands:
            RH = WARPS * R
            th_max = RH - 2 * k
            if th_max < 1:
                continue
            nty = math.ceil(tile_rows / th_max)
            TH = math.ceil(tile_rows / nty)      # balance the rows over the same number of tiles
            ntiles = ntx * nty
            waves = math.ceil(ntiles / sms)
            t_tile = (RH * REGION_BYTES) / LOAD_BYTES_PER_US_PER_SM \
                + k * RH * REGION_BYTES / rate \
                + (TH * TWB) / LOAD_BYTES_PER_US_PER_SM
            t_launch = LAUNCH_US + waves * t_tile
            total = math.ceil(loops / k) * t_launch
            cand = StreamPlan(k, R, TH, TWB, HB, ntx, nty, waves, t_launch)
            if best is None or total < best[0] - 1e-9:
                best = (total, cand)
    if best is None:
        raise ValueError("no feasible stream plan (tile too small for any halo depth?)")
    return best[1]


EXCHANGE_US = 3.0            # modelled cost of one in-kernel halo exchange (flags + strip traffic)
EXCHANGE_US_PEER = 6.0       # ... when neighbours sit on other GPUs (NVLink round trips)


def plan_resident(tile_rows: int, tile_cb: int, C: int, kind: int, loops: int, *,
                  min_nbr_rows: int | None = None, min_nbr_cols: int | None = None,
                  sms: int = DEFAULT_SMS, bands: list[int] | None = None, peers: bool = False,
                  force_k: int | None = None, force_R: int | None = None) -> StreamPlan | None:
    """Plan for the persistent resident kernel: the whole tile must fit the GPU's CTAs at once
    (``ntx*nty <= sms``). Returns None when it does not (the caller streams instead)."""
    loops = max(loops, 1)
    bands = sorted(bands or DEFAULT_BANDS[kind])
    if force_R:
        bands = [force_R]
    rate = BYTES_PER_CLK[kind] * CLOCK_GHZ * 1e3
    kmax = min(loops, (WARPS * bands[-1] - 1) // 2, (REGION_BYTES // 2 - 64) // C)
    if min_nbr_rows is not None:
        kmax = min(kmax, min_nbr_rows)
    if min_nbr_cols is not None:
        kmax = min(kmax, min_nbr_cols)
    kmax = max(kmax, 1)
    ks = [force_k] if force_k else range(1, kmax + 1)
    t_x = EXCHANGE_US_PEER if peers else EXCHANGE_US
    best: tuple[float, StreamPlan] | None = None
    for k in ks:
        HB, TWB = region_geometry(k, C)
        if TWB < max(64, HB):
            continue
        ntx = math.ceil(tile_cb / TWB)
        for R in bands:
            RH = WARPS * R
            th_max = RH - 2 * k
            if th_max < k:
                continue
            nty = math.ceil(tile_rows / th_max)
            if ntx * nty > sms:
                continue
            TH = max(math.ceil(tile_rows / nty), k)
            nblocks = math.ceil(loops / k)
            t_block = k * RH * REGION_BYTES / rate
            total = LAUNCH_US + (RH * REGION_BYTES + TH * TWB) / LOAD_BYTES_PER_US_PER_SM \
                + nblocks * t_block + (nblocks - 1) * t_x
            cand = StreamPlan(k, R, TH, TWB, HB, ntx, nty, 1, total)
            if best is None or total < best[0] - 1e-9:
                best = (total, cand)
    return best[1] if best else None
