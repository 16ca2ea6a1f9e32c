"""Fused stencil + halo backend (the product): drives ``stream_kernel`` / ``persist_kernel``.

One *tile driver* per rank owns two ping-pong tile buffers, a flag pad, a status block and an
*exchange arena* in plain ``cudaMalloc`` memory (IPC-exportable, see ``peer_memory.py``). A streaming
launch gets the base pointers of the up-to-eight neighbouring tiles; the kernel pulls halo rows /
columns / corners straight out of those buffers (NVLink P2P loads across GPUs) and synchronises with
per-launch epoch flags written by the neighbours with ``st.release.sys``. The persistent kernel keeps
the tile in registers for a whole run and its CTAs *push* their borders into the neighbours' inbox
slots inside the arenas (NVLink P2P stores across GPUs). There is no NCCL and no host round trip on
this path; the host enqueues every launch of a run up front.

Replaces the reference's ``MPI_Isend/Irecv`` + ``MPI_Wait`` iteration loop
(``mpi/mpi_conv.c:164-248``) for the GPU path.

``VirtualCluster`` runs N tile drivers on ONE device in one process (peer pointers degenerate to
local pointers) — the way multi-GPU logic is tested on a 1-GPU box.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ..models.filters import Filter
from ..ops import cuda as K
from .decomposition import TileGrid
from .plan import StreamPlan, plan_stream

import os

FLAG_BYTES = 256      # bank 0 (+0): launch epochs, bank 1 (+64): "input ready" step counters
BANK1 = 64
STATUS_WORDS = 14     # u64: error, max wait ns, sum wait ns, waits; persistent-kernel probe CTA: push /
                      # signal+wait+fetch / inbox->register / compute ns, exchanges, of which signal / own poll
# control block layout (one cudaMalloc, one IPC handle): flags | counter | status
OFF_COUNTER, OFF_STATUS = FLAG_BYTES, FLAG_BYTES + 64
CTL_BYTES = OFF_STATUS + 128
# Exchange arena of the persistent kernel (csrc/common/xplan.h): 64 KB of per-CTA flag words + two
# parities of one inbox slot per CTA. Fixed size so that it can be IPC-mapped once; plans whose slots
# do not fit are not offered (a 1920x2520 RGB image on one GPU needs 5 MB).
ARENA_BYTES = int(os.environ.get("PCONV_ARENA_MIB", "48")) << 20


def tile_pitch(grid: TileGrid) -> int:
    """Common byte pitch of every tile buffer: widest tile rounded up to 128 B."""
    cb = grid.max_tile_cols() * grid.channels
    return (cb + 127) // 128 * 128


@dataclass
class RankMemory:
    """Device memory of one rank. ``owned`` is False for peer-mapped views of another rank."""

    bufs: list[int]
    flags: int
    counter: int = 0
    status: int = 0
    nbytes: int = 0
    owned: bool = True
    arena: int = 0

    @staticmethod
    def allocate(rows: int, pitch: int, slots: int = 1) -> "RankMemory":
        """``slots`` independent ping-pong buffer pairs (slot s = bufs[2s], bufs[2s+1]) so that the
        host copies of one step overlap the iterations of another."""
        nbytes = rows * pitch + 256
        bufs = [K.malloc(nbytes) for _ in range(2 * slots)]
        ctl = K.malloc(CTL_BYTES)
        return RankMemory(bufs, ctl, ctl + OFF_COUNTER, ctl + OFF_STATUS, nbytes, True, K.malloc(ARENA_BYTES))

    @staticmethod
    def view(bufs: list[int], ctl: int, nbytes: int, arena: int = 0) -> "RankMemory":
        """Peer-mapped view of another rank's memory (same control-block layout)."""
        return RankMemory(bufs, ctl, ctl + OFF_COUNTER, ctl + OFF_STATUS, nbytes, False, arena)

    def free(self) -> None:
        if self.owned:
            for p in self.bufs:
                K.free(p)
            K.free(self.flags)
            if self.arena:
                K.free(self.arena)
        self.bufs, self.flags, self.arena = [], 0, 0


@dataclass
class RunStats:
    launches: int = 0
    watchdog: bool = False
    max_wait_us: float = 0.0
    sum_wait_us: float = 0.0
    waits: int = 0
    plan: StreamPlan | None = None


class TileDriver:
    """Enqueues the fused kernel launches for one tile."""

    def __init__(self, grid: TileGrid, rank: int, filt: Filter, mem: RankMemory,
                 neighbors: list[RankMemory | None], pitch: int, stream: int = 0,
                 timeout_ms: float = 10_000.0):
        self.grid, self.rank, self.filt, self.mem = grid, rank, filt, mem
        self.neighbors = neighbors          # 9 entries, index 4 = self.mem
        self.pitch, self.stream = pitch, stream
        self.timeout_ns = int(timeout_ms * 1e6)
        self.tile = grid.tile(rank)
        self.inject_delay_ns = 0            # fault injection (tests / --inject-*)
        self.inject_drop_signal = False
        self._geom = self._geometry()

    # -- geometry ---------------------------------------------------------------------------
    def _geometry(self) -> K.Geometry:
        g, t = self.grid, self.tile
        rb, cbn = g.row_bounds(), g.col_bounds()
        Cn = g.channels

        def window(bounds: list[int], i: int, n: int) -> list[int]:
            lo_prev = bounds[i - 1] if i > 0 else bounds[i]
            hi_next = bounds[i + 2] if i + 1 < n else bounds[i + 1]
            return [lo_prev, bounds[i], bounds[i + 1], hi_next]

        geo = K.Geometry()
        geo.img_rows, geo.img_cb = g.height, g.width * Cn
        geo.row_lo = (C.c_int64 * 4)(*window(rb, t.ty, g.grid_rows))
        geo.cb_lo = (C.c_int64 * 4)(*[v * Cn for v in window(cbn, t.tx, g.grid_cols)])
        geo.pitch, geo.C = self.pitch, Cn
        return geo

    # -- launches ---------------------------------------------------------------------------
    def _params(self, plan: StreamPlan, k: int, src_parity: int, wait: int, signal: int,
                grid_ctas: int = 0, slot: int = 0) -> K.StreamParams:
        p = K.StreamParams()
        p.g = self._geom
        src_parity += 2 * slot
        for d in range(9):
            nb = self.neighbors[d]
            p.src[d] = nb.bufs[src_parity] if nb is not None else None
            p.src_pitch[d] = self.pitch
            p.sync.remote_flags[d] = (nb.flags + 4 * (8 - d)) if (nb is not None and d != 4) else None
        p.dst = self.mem.bufs[src_parity ^ 1]
        p.filt = self.filt.spec
        p.k, p.TH, p.TWB, p.HB, p.grid, p.R = k, plan.TH, plan.TWB, plan.HB, grid_ctas, plan.R
        # (a shallower tail launch, k < plan.k, keeps TH/HB: HB >= k*C still holds and RH shrinks)
        p.sync.local_flags = self.mem.flags
        p.sync.done_counter = self.mem.counter
        p.sync.status = self.mem.status
        p.sync.wait_value, p.sync.signal_value = wait, signal
        p.sync.timeout_ns = self.timeout_ns
        p.sync.inject_delay_ns = self.inject_delay_ns
        p.sync.inject_drop_signal = 1 if self.inject_drop_signal else 0
        return p

    def persist_params(self, plan: StreamPlan, loops: int, src_parity: int, wait: int, signal: int,
                       cta_epoch: int, slot: int = 0, wait_bank1: bool = False,
                       peer_scope: bool = False) -> K.PersistParams:
        """One launch of the persistent kernel: reads buffer ``src_parity`` (and the neighbours'), writes
        the other buffer of the pair; exchange b carries epoch ``cta_epoch + b``."""
        p = K.PersistParams()
        p.g = self._geom
        src_parity += 2 * slot
        for d in range(9):
            nb = self.neighbors[d]
            p.src[d] = nb.bufs[src_parity] if nb is not None else None
            p.arena[d] = nb.arena if nb is not None else None
            p.sync.remote_flags[d] = (nb.flags + 4 * (8 - d)) if (nb is not None and d != 4) else None
        p.dst = self.mem.bufs[src_parity ^ 1]
        p.arena_bytes = ARENA_BYTES
        p.filt = self.filt.spec
        p.loops, p.k, p.TH, p.TWB, p.HB, p.R, p.W = loops, plan.k, plan.TH, plan.TWB, plan.HB, plan.R, plan.W
        p.nty, p.ntx = plan.nty, plan.ntx
        p.epoch_base = cta_epoch & 0xFFFFFFFF
        p.peer_scope = 1 if peer_scope else 0
        p.debug_skip = int(os.environ.get("PCONV_PERSIST_SKIP", "0"))   # timing experiments only
        p.sync.local_flags = self.mem.flags + (BANK1 if wait_bank1 else 0)
        p.sync.done_counter = self.mem.counter
        p.sync.status = self.mem.status
        p.sync.wait_value, p.sync.signal_value = wait, signal
        p.sync.timeout_ns = self.timeout_ns
        p.sync.inject_drop_signal = 1 if self.inject_drop_signal else 0
        return p

    def sync_params(self, wait: int, signal: int, bank: int = 0) -> K.Sync:
        """Parameters of a stand-alone protocol step on flag bank 0 (launch epochs) or 1 (inputs). The struct
        is cached per bank (this runs once or twice per step of a ~200 us step): only the epochs change."""
        cache = self.__dict__.setdefault("_sync_cache", {})
        s = cache.get(bank)
        if s is not None:
            s.wait_value, s.signal_value = wait, signal
            s.inject_drop_signal = 1 if self.inject_drop_signal else 0
            return s
        s = cache[bank] = K.Sync()
        off = BANK1 if bank else 0
        s.local_flags = self.mem.flags + off
        for d in range(9):
            nb = self.neighbors[d]
            s.remote_flags[d] = (nb.flags + off + 4 * (8 - d)) if (nb is not None and d != 4) else None
        s.done_counter, s.status = self.mem.counter, self.mem.status
        s.wait_value, s.signal_value, s.timeout_ns = wait, signal, self.timeout_ns
        s.inject_drop_signal = 1 if self.inject_drop_signal else 0
        return s

    # -- data movement ----------------------------------------------------------------------
    def upload(self, tile: np.ndarray, parity: int = 0) -> None:
        t = self.tile
        cb = t.cols * self.grid.channels
        arr = np.ascontiguousarray(tile, dtype=np.uint8).reshape(t.rows, cb)
        K.memcpy2d_async(self.mem.bufs[parity], self.pitch, arr.ctypes.data, cb, cb, t.rows,
                         K.H2D, self.stream)
        K.stream_sync(self.stream)  # arr may be a temporary

    def download(self, parity: int) -> np.ndarray:
        t = self.tile
        cb = t.cols * self.grid.channels
        out = np.empty((t.rows, cb), dtype=np.uint8)
        K.memcpy2d_async(out.ctypes.data, cb, self.mem.bufs[parity], self.pitch, cb, t.rows,
                         K.D2H, self.stream)
        K.stream_sync(self.stream)
        return out

    def synth(self, seed: int, parity: int = 0) -> None:
        t, Cn = self.tile, self.grid.channels
        K.synth_rect(seed, self.grid.row_bytes, t.row0, t.rows, t.col0 * Cn, t.cols * Cn,
                     self.mem.bufs[parity], self.pitch, self.stream)

    def reset_status(self) -> None:
        K.memset_async(self.mem.status, 0, 8 * STATUS_WORDS, self.stream)

    def read_status(self) -> tuple[bool, float, float, int]:
        out = np.zeros(STATUS_WORDS, dtype=np.uint64)
        K.memcpy_async(out.ctypes.data, self.mem.status, 8 * STATUS_WORDS, K.D2H, self.stream)
        K.stream_sync(self.stream)
        self.probe = {"publish_us": float(out[4]) / 1e3, "wait_us": float(out[5]) / 1e3,
                      "reload_us": float(out[6]) / 1e3, "compute_us": float(out[7]) / 1e3,
                      "exchanges": int(out[8]), "signal_us": float(out[9]) / 1e3, "poll_us": float(out[10]) / 1e3,
                      "push_hz_us": float(out[11]) / 1e3, "push_bar1_us": float(out[12]) / 1e3,
                      "push_vert_us": float(out[13]) / 1e3}
        return bool(out[0]), float(out[1]) / 1e3, float(out[2]) / 1e3, int(out[3])


def make_persist_plan(grid: TileGrid, filt: Filter, loops: int, halo_depth: int | None = None,
                      sms: int | None = None, bands: list[int] | None = None,
                      band: int | None = None, peers: bool = False, warps: int = 8) -> StreamPlan | None:
    """Plan for the persistent kernel with ``warps`` warps per CTA, or None when the tile does not fit
    on chip (or its inbox slots do not fit the exchange arena)."""
    from .plan import plan_persist
    if not grid.exact and grid.workers > 1:
        return None      # CTA coordinates wrap into the neighbouring rank: identical tile shapes everywhere
    Cn = grid.channels
    multi_r, multi_c = grid.grid_rows > 1, grid.grid_cols > 1
    kw = {}
    if sms:
        kw["sms"] = sms
    if bands:
        kw["bands"] = bands
    p = plan_persist(grid.max_tile_rows(), grid.max_tile_cols() * Cn, Cn, filt.kind, max(loops, 1),
                     min_nbr_rows=grid.min_tile_rows() if multi_r else None,
                     min_nbr_cols=grid.min_tile_cols() if multi_c else None,
                     force_k=halo_depth, force_R=band, peers=peers, warps=warps, **kw)
    if p is not None and persist_arena_bytes(p) > ARENA_BYTES:
        return None
    return p


def persist_arena_bytes(p: StreamPlan) -> int:
    """csrc/common/xplan.h: 64 KB of flags + 2 parities x CTAs x slot (blocks N,S | W,E | 4 corners)."""
    nh, nc = p.HB // 16, p.TWB // 16
    chunks = 2 * p.k * nc + 2 * p.TH * nh + 4 * p.k * nh
    chunks = (chunks + 7) & ~7
    return 1024 * 64 + 2 * p.nty * p.ntx * chunks * 16


def make_plan(grid: TileGrid, filt: Filter, loops: int, halo_depth: int | None = None,
              sms: int | None = None, bands: list[int] | None = None,
              band: int | None = None) -> StreamPlan:
    """The (rank-independent) stream plan for a grid."""
    Cn = grid.channels
    multi_r, multi_c = grid.grid_rows > 1, grid.grid_cols > 1
    kw = {}
    if sms:
        kw["sms"] = sms
    if bands:
        kw["bands"] = bands
    return plan_stream(grid.max_tile_rows(), grid.max_tile_cols() * Cn, Cn, filt.kind,
                       max(loops, 1),
                       min_nbr_rows=grid.min_tile_rows() if multi_r else None,
                       min_nbr_cols=grid.min_tile_cols() if multi_c else None,
                       force_k=halo_depth, force_R=band, **kw)


class VirtualCluster:
    """N tiles on one device in one process (``FusedRuntime`` with ``world == 1``).

    The kernels, flag protocol and owner lookup are exactly the multi-GPU ones; only the pointers
    are local. Used by the tests and by ``ConvEngine(workers=N)`` on a single GPU.
    """

    def __init__(self, grid: TileGrid, filt: Filter, halo_depth: int | None = None,
                 band: int | None = None, mode: str = "stream", timeout_ms: float = 10_000.0,
                 warps: int | None = None):
        from .fused_runtime import FusedRuntime
        self.grid = grid
        self.rt = FusedRuntime(grid, filt, 0, 1, None, halo_depth=halo_depth, band=band, mode=mode,
                               timeout_ms=timeout_ms, warps=warps)
        self.sms, self.bands = self.rt.sms, self.rt.bands

    def close(self) -> None:
        self.rt.close()

    def upload(self, image: np.ndarray) -> None:
        g = self.grid
        img = np.ascontiguousarray(image, dtype=np.uint8).reshape(g.height, g.row_bytes)
        Cn = g.channels
        self.rt.upload({t.rank: img[t.row0:t.row0 + t.rows, t.col0 * Cn:(t.col0 + t.cols) * Cn]
                        for t in g.tiles()})

    def synth(self, seed: int) -> None:
        self.rt.synth(seed)

    def download(self) -> np.ndarray:
        g, Cn = self.grid, self.grid.channels
        out = np.empty((g.height, g.row_bytes), dtype=np.uint8)
        for r, tile in self.rt.download().items():
            t = g.tile(r)
            out[t.row0:t.row0 + t.rows, t.col0 * Cn:(t.col0 + t.cols) * Cn] = tile
        return out

    def run(self, loops: int) -> RunStats:
        """Enqueue ``loops`` iterations (asynchronous; ``download`` synchronises)."""
        for d in self.rt.drivers.values():
            d.reset_status()
        plan, launches = self.rt.enqueue(loops)
        self.last = RunStats(launches=launches, plan=plan)
        return self.last

    def collect_status(self) -> RunStats:
        s = self.last
        for d in self.rt.drivers.values():
            err, mx, sm, n = d.read_status()
            s.watchdog |= err
            s.max_wait_us = max(s.max_wait_us, mx)
            s.sum_wait_us += sm
            s.waits += n
        return s
