"""Launch plans for the sm_100a kernels — the Python face of the native planner.

The search itself is C++ (``csrc/host/planner.cpp``, shared with the stand-alone native CLI): it
picks the temporal-blocking depth ``k``, the band height ``R`` (region = 16*R rows x 512 bytes held
in registers) and the interior rows ``TH`` per CTA by minimising a time model that includes the
wave quantisation over the SMs. ``plan_persist`` (the persistent kernel: ``warps`` = 8 runs two CTAs
per SM, 16 one) additionally requires every CTA of the tile to be co-resident and returns ``None``
when the tile does not fit on chip.
"""
from __future__ import annotations

from dataclasses import dataclass

from ..ops import host as H
from ..ops.host import KIND_F32, KIND_INT, KIND_SEP121

REGION_BYTES = 512          # kSP in csrc/cuda/stencil_core.cuh
WARPS = 16
DEFAULT_SMS = 148
# Band heights compiled into libpconv_cuda.so (csrc/cuda/stream_kernel.cu); refined from the
# library at runtime (ops.cuda.band_heights).
DEFAULT_BANDS = {KIND_SEP121: [2, 4, 7, 10, 14], KIND_INT: [2, 4, 7, 10], KIND_F32: [2, 4, 6, 8, 10]}
# persistent kernel (csrc/cuda/persist_kernel.cuh)
DEFAULT_PERSIST_BANDS = {KIND_SEP121: [2, 3, 4, 5, 6, 7, 8, 10, 12, 14], KIND_INT: [2, 4, 7, 10],
                         KIND_F32: [2, 4, 6, 8, 10]}


@dataclass(frozen=True)
class StreamPlan:
    k: int            # fused iterations per launch (stream) / per exchange (persistent)
    R: int            # band height (rows per warp); region rows = W*R
    TH: int           # interior rows per region
    TWB: int          # interior bytes per region
    HB: int           # column halo bytes
    ntx: int
    nty: int
    waves: int
    est_us_per_launch: float
    W: int = WARPS    # warps per CTA (persistent kernel: 8 or 16)

    @property
    def RH(self) -> int:
        return self.W * self.R

    def launches(self, loops: int) -> list[int]:
        """k of every launch for ``loops`` iterations (the last one may be shallower)."""
        full, rem = divmod(loops, self.k)
        return [self.k] * full + ([rem] if rem else [])


def _to_plan(r, warps: int = WARPS) -> StreamPlan | None:
    if r is None:
        return None
    return StreamPlan(r.k, r.R, r.TH, r.TWB, r.HB, r.ntx, r.nty, r.waves, float(r.est_us), warps)


def plan_stream(tile_rows: int, tile_cb: int, C: int, kind: int, loops: int, *,
                min_nbr_rows: int | None = None, min_nbr_cols: int | None = None,
                sms: int = DEFAULT_SMS, bands: list[int] | None = None,
                force_k: int | None = None, force_R: int | None = None) -> StreamPlan:
    p = _to_plan(H.plan(tile_rows, tile_cb, C, kind, loops, sorted(bands or DEFAULT_BANDS[kind]),
                        min_nbr_rows=min_nbr_rows or 0, min_nbr_cols=min_nbr_cols or 0, sms=sms,
                        force_k=force_k or 0, force_R=force_R or 0))
    if p is None:
        raise ValueError("no feasible stream plan (tile too small for any halo depth?)")
    return p


def plan_persist(tile_rows: int, tile_cb: int, C: int, kind: int, loops: int, *,
                 min_nbr_rows: int | None = None, min_nbr_cols: int | None = None,
                 sms: int = DEFAULT_SMS, bands: list[int] | None = None, peers: bool = False,
                 force_k: int | None = None, force_R: int | None = None,
                 warps: int = 8) -> StreamPlan | None:
    """Plan of the persistent kernel with ``warps`` warps per CTA (8: two CTAs share an SM and hide each
    other's exchanges; 16: one CTA per SM). None when the tile does not fit on chip."""
    if warps not in (8, 16):
        raise ValueError("warps per CTA must be 8 or 16")
    return _to_plan(H.plan(tile_rows, tile_cb, C, kind, loops, sorted(bands or DEFAULT_PERSIST_BANDS[kind]),
                           min_nbr_rows=min_nbr_rows or 0, min_nbr_cols=min_nbr_cols or 0, sms=sms,
                           force_k=force_k or 0, force_R=force_R or 0, resident=True, peers=peers,
                           warps=warps), warps)
