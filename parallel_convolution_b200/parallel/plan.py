"""Launch plans for the sm_100a kernels — the Python face of the native planner.

The search itself is C++ (``csrc/host/planner.cpp``, shared with the stand-alone native CLI): it
picks the temporal-blocking depth ``k``, the band height ``R`` (region = 16*R rows x 512 bytes held
in registers) and the interior rows ``TH`` per CTA by minimising a time model that includes the
wave quantisation over the SMs. ``plan_resident`` additionally requires every CTA of the tile to be
co-resident (``ntx*nty <= #SMs``) and returns ``None`` when the tile does not fit on chip.
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
# split-phase resident kernel (csrc/cuda/resident2_kernel.cuh): rows per warp per virtual tile
DEFAULT_DUAL_BANDS = {KIND_SEP121: [2, 4, 5, 7], KIND_INT: [2, 4, 5], KIND_F32: [2, 4, 5]}


@dataclass(frozen=True)
class StreamPlan:
    k: int            # fused iterations per launch (stream) / per exchange (resident)
    R: int            # band height (rows per warp); region rows = 16*R
    TH: int           # interior rows per region
    TWB: int          # interior bytes per region
    HB: int           # column halo bytes
    ntx: int
    nty: int
    waves: int
    est_us_per_launch: float
    dual: bool = False   # split-phase resident kernel: nty counts virtual tile rows, two per CTA

    @property
    def RH(self) -> int:
        return WARPS * self.R

    def launches(self, loops: int) -> list[int]:
        """k of every launch for ``loops`` iterations (the last one may be shallower)."""
        full, rem = divmod(loops, self.k)
        return [self.k] * full + ([rem] if rem else [])


def _to_plan(r, dual: bool = False) -> StreamPlan | None:
    if r is None:
        return None
    return StreamPlan(r.k, r.R, r.TH, r.TWB, r.HB, r.ntx, r.nty, r.waves, float(r.est_us), dual)


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


def plan_resident(tile_rows: int, tile_cb: int, C: int, kind: int, loops: int, *,
                  min_nbr_rows: int | None = None, min_nbr_cols: int | None = None,
                  sms: int = DEFAULT_SMS, bands: list[int] | None = None, peers: bool = False,
                  force_k: int | None = None, force_R: int | None = None,
                  dual: bool | int = False) -> StreamPlan | None:
    """``dual``: 1/True plans the split-phase kernel (``bands`` are then its per-virtual-tile heights),
    2 its protocol-warp variant (regions of 15*R rows)."""
    default = DEFAULT_DUAL_BANDS if dual else DEFAULT_BANDS
    return _to_plan(H.plan(tile_rows, tile_cb, C, kind, loops, sorted(bands or default[kind]),
                           min_nbr_rows=min_nbr_rows or 0, min_nbr_cols=min_nbr_cols or 0, sms=sms,
                           force_k=force_k or 0, force_R=force_R or 0, resident=True, peers=peers,
                           dual=int(dual)), bool(dual))
