"""Launch planner: region geometry and temporal-blocking depth for the sm_100a kernels.

A CTA region is ``RH = TH + 2k`` rows x 512 bytes of shared memory (two ping-pong copies), of
which the interior ``TH x TWB`` (``TWB = 512 - 2*HB``, ``HB = roundup16(k*C)``) is written back
after ``k`` fused iterations. Deeper ``k`` divides the number of launches / cross-GPU
synchronisations by ``k`` (every BASELINE config is latency- or ALU-bound, never link-bound:
SURVEY.md §7.2) but recomputes a thicker halo; the planner minimises a simple time model over
``(k, TH)`` including the wave quantisation over the 148 SMs.

The plan must be identical on every rank, so it only depends on quantities every rank knows
(the largest/smallest tile of the grid, the filter path, the SM count).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

from ..ops.host import KIND_F32, KIND_INT, KIND_SEP121

REGION_BYTES = 512          # kSP in csrc/cuda/stencil_core.cuh
DEFAULT_SMS = 148
DEFAULT_MAX_RH = 224        # (227 KB - slack) / (2 * 512 B); refined from the device at runtime

# Modelled SM throughput in region bytes per clock (measured on B200, profiles/): used only to
# rank candidate plans, never reported as a result.
BYTES_PER_CLK = {KIND_SEP121: 26.0, KIND_INT: 11.0, KIND_F32: 4.5}
CLOCK_GHZ = 1.8
LAUNCH_US = 4.0             # dependent-launch gap incl. the kernel prologue/epilogue
LOAD_BYTES_PER_US_PER_SM = 40e3  # ~6 TB/s / 148 SMs


@dataclass(frozen=True)
class StreamPlan:
    k: int            # fused iterations per launch
    TH: int           # interior rows per region
    TWB: int          # interior bytes per region
    HB: int           # column halo bytes
    ntx: int
    nty: int
    waves: int
    est_us_per_launch: float

    @property
    def RH(self) -> int:
        return self.TH + 2 * self.k

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
                sms: int = DEFAULT_SMS, max_rh: int = DEFAULT_MAX_RH,
                force_k: int | None = None) -> StreamPlan:
    """Pick (k, TH) for a tile of ``tile_rows`` x ``tile_cb`` bytes.

    ``min_nbr_rows`` / ``min_nbr_cols`` (pixels): smallest neighbouring tile extent — the halo may
    only reach into adjacent tiles, so ``k`` cannot exceed them.
    """
    if loops <= 0:
        loops = 1
    rate = BYTES_PER_CLK[kind] * CLOCK_GHZ * 1e3  # region bytes per us per SM
    kmax = min(loops, (max_rh - 1) // 2, (REGION_BYTES // 2 - 64) // C)
    if min_nbr_rows is not None:
        kmax = min(kmax, min_nbr_rows)
    if min_nbr_cols is not None:
        kmax = min(kmax, min_nbr_cols)
    kmax = max(kmax, 1)
    ks = [force_k] if force_k else range(1, kmax + 1)
    best: tuple[float, StreamPlan] | None = None
    for k in ks:
        if k > kmax and not force_k:
            continue
        HB, TWB = region_geometry(k, C)
        if TWB < 64:
            continue
        ntx = math.ceil(tile_cb / TWB)
        th_max = max_rh - 2 * k
        if th_max < 1:
            continue
        nty_min = math.ceil(tile_rows / th_max)
        for nty in range(nty_min, min(nty_min + 48, tile_rows) + 1):
            TH = math.ceil(tile_rows / nty)
            if TH > th_max:
                continue
            nty_eff = math.ceil(tile_rows / TH)
            ntiles = ntx * nty_eff
            waves = math.ceil(ntiles / sms)
            RH = TH + 2 * k
            rows_computed = sum(max(RH - 2 * it, 0) for it in range(1, k + 1))
            t_tile = (RH * REGION_BYTES) / LOAD_BYTES_PER_US_PER_SM \
                + rows_computed * REGION_BYTES / rate \
                + (TH * TWB) / LOAD_BYTES_PER_US_PER_SM
            t_launch = LAUNCH_US + waves * t_tile
            n_launch = math.ceil(loops / k)
            total = n_launch * t_launch
            cand = StreamPlan(k, TH, TWB, HB, ntx, nty_eff, waves, t_launch)
            if best is None or total < best[0] - 1e-9:
                best = (total, cand)
    if best is None:
        raise ValueError("no feasible stream plan (tile too small for any halo depth?)")
    return best[1]
