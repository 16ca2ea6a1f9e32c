"""NCCL baseline backend: pack -> ncclSend/ncclRecv -> unpack -> stencil kernel.

"A path that only calls NCCL/MPI for the halo exchange is the baseline, not the product"
(BASELINE.json) — this is that baseline, kept honest so the fused backend has something real to
beat. It is the GPU analogue of the reference's MPI schedule (``mpi/mpi_conv.c:164-248``):

    every k iterations:  ONE pack kernel (all border strips -> dense send buffers, 128-bit copies)
                         ONE ncclGroup of send/recv pairs (torch.distributed.batch_isend_irecv)
                         ONE unpack kernel (dense receive buffers -> ghost strips)
                         one stream_kernel launch of k fused iterations

and the whole run (every exchange and launch of ``loops`` iterations) is captured once into a CUDA graph
and replayed, so no Python or launch latency sits between the steps — a tuned send/recv implementation,
not a strawman. The graph replay is opt-in (``PCONV_NCCL_GRAPH=1``, used by ``bench/halo_sweep.py``):
captured NCCL groups need every rank to capture and replay in lockstep, which a library default must not
assume of its callers.

The stencil kernel is the same sm_100a kernel as the fused backend; the only difference is where
the halo comes from: instead of peer-mapped neighbour tiles, its eight neighbour "sources" are
small local *ghost* buffers addressed through virtual base pointers
(``pconv_stream_params_t.src_pitch``), filled by NCCL. Same k-deep temporal blocking, so the
comparison isolates the cost of staging + NCCL + the extra launches.
"""
from __future__ import annotations

import os

import numpy as np

from ..models.filters import Filter
from ..ops import cuda as K
from .decomposition import NEIGHBOR_OFFSETS, TileGrid
from .fused import RankMemory, TileDriver, make_plan, tile_pitch
from .plan import REGION_BYTES, StreamPlan


def _r16(v: int) -> int:
    return (v + 15) // 16 * 16


class NcclRuntime:
    def __init__(self, grid: TileGrid, filt: Filter, rank: int, world: int, dist,
                 halo_depth: int | None = None, band: int | None = None):
        import torch
        if world > 1 and grid.workers != world:
            raise ValueError("one tile per process: grid size must equal the world size")
        self.torch, self.dist = torch, dist
        self.grid, self.filt, self.rank, self.world = grid, filt, rank, world
        self.halo_depth, self.band = halo_depth, band
        self.pitch = tile_pitch(grid)
        self.tstream = torch.cuda.current_stream()
        self.stream = int(self.tstream.cuda_stream)
        self.use_graph = os.environ.get("PCONV_NCCL_GRAPH", "0") not in ("", "0")
        self._graphs: dict = {}
        self.sms = K.sm_count()
        self.bands = K.band_heights(filt.kind)
        self.local_ranks = [rank] if world > 1 else [0]
        self.me = self.local_ranks[0]
        self.mems = {self.me: RankMemory.allocate(grid.max_tile_rows(), self.pitch)}
        self.nbr_ranks = grid.neighbors(self.me)
        self.driver = TileDriver(grid, self.me, filt, self.mems[self.me],
                                 [self.mems[self.me] if d == 4 else None for d in range(9)],
                                 self.pitch, self.stream)
        self.drivers = {self.me: self.driver}
        self.parity, self.slot = 0, 0
        self.multi = False            # no in-kernel flags on this path
        self._plans: dict[int, StreamPlan] = {}
        self._ghost_key = None
        self.exchanges = 0
        # PCONV_POISON_HALO=1: refill the ghost strips with 0xA5 before every exchange, so a region the
        # kernel reads but the exchange did not deliver shows up as wrong pixels
        self.poison = os.environ.get("PCONV_POISON_HALO", "0") not in ("", "0")

    # -- lifecycle / data (same surface as FusedRuntime) --------------------------------------------
    def close(self) -> None:
        self.torch.cuda.synchronize()
        for m in self.mems.values():
            m.free()
        self.mems = {}

    def upload(self, tiles: dict[int, np.ndarray]) -> None:
        self.driver.upload(tiles[self.me], 0)
        self.parity = 0

    def synth(self, seed: int) -> None:
        self.driver.synth(seed, 0)
        self.parity = 0

    def download(self) -> dict[int, np.ndarray]:
        return {self.me: self.driver.download(self.parity)}

    def begin_input(self) -> int:
        return 0

    def end_input(self) -> int:
        return 0

    def plan(self, loops: int) -> StreamPlan:
        if loops not in self._plans:
            self._plans[loops] = make_plan(self.grid, self.filt, loops, self.halo_depth, self.sms,
                                           self.bands, self.band)
        return self._plans[loops]

    # -- ghost strips ---------------------------------------------------------------------------------
    def _build_ghosts(self, plan: StreamPlan) -> None:
        """Allocate send/recv/ghost buffers for this plan's (k, HB, RH)."""
        key = (plan.k, plan.HB, plan.RH)
        if self._ghost_key == key:
            return
        torch, g, C = self.torch, self.grid, self.grid.channels
        dev = torch.device("cuda", torch.cuda.current_device())
        me = g.tile(self.me)
        my_rows, my_cb = me.rows, me.cols * C
        k, HB, RH = plan.k, plan.HB, plan.RH
        self.links = []
        for d, ((dy, dx), r) in enumerate(zip(NEIGHBOR_OFFSETS, self.nbr_ranks)):
            if r < 0 or d == 4:
                continue
            nt = g.tile(r)
            n_rows, n_cb = nt.rows, nt.cols * C
            # --- what I receive: the neighbour's strip facing me, placed in a ghost of its tile ----
            if dy < 0:      # neighbour above: its last k rows
                g_row0, g_rows, fill_r0, fill_rows = n_rows - k, k, n_rows - k, k
            elif dy > 0:    # below: its first k rows (ghost deep enough for any region overhang)
                g_row0, g_rows, fill_r0, fill_rows = 0, min(n_rows, RH), 0, k
            else:           # same tile row: all my rows
                g_row0, g_rows, fill_r0, fill_rows = 0, n_rows, 0, n_rows
            if dx < 0:      # neighbour to the left: its last HB bytes (16-byte aligned start)
                c0 = max(0, (n_cb - HB) // 16 * 16)
                g_c0, g_cb, fill_c0, fill_cb = c0, n_cb - c0, c0, n_cb - c0
            elif dx > 0:    # right: its first HB bytes (ghost wide enough for the region overhang)
                g_c0, g_cb, fill_c0, fill_cb = 0, min(n_cb, REGION_BYTES), 0, min(HB, n_cb)
            else:
                g_c0, g_cb, fill_c0, fill_cb = 0, n_cb, 0, n_cb
            gp = _r16(g_cb)
            ghost = torch.zeros(g_rows * gp + 64, dtype=torch.uint8, device=dev)
            recv = torch.empty(fill_rows * fill_cb, dtype=torch.uint8, device=dev)
            vbase = ghost.data_ptr() - g_row0 * gp - g_c0        # virtual base of the neighbour tile
            # --- what I send: my strip facing the neighbour (mirror of the above from its side) -----
            if dy < 0:
                s_r0, s_rows = 0, k
            elif dy > 0:
                s_r0, s_rows = my_rows - k, k
            else:
                s_r0, s_rows = 0, my_rows
            if dx < 0:
                s_c0, s_cb = 0, min(HB, my_cb)
            elif dx > 0:
                s_c0 = max(0, (my_cb - HB) // 16 * 16)
                s_cb = my_cb - s_c0
            else:
                s_c0, s_cb = 0, my_cb
            send = torch.empty(max(s_rows * s_cb, 16), dtype=torch.uint8, device=dev)[:s_rows * s_cb]
            self.links.append({"d": d, "rank": r, "ghost": ghost, "gp": gp, "vbase": vbase,
                               "recv": recv, "fill": (fill_r0 - g_row0, fill_c0 - g_c0, fill_rows, fill_cb),
                               "send": send, "strip": (s_r0, s_c0, s_rows, s_cb)})
        self._ghost_key = key
        self._graphs = {}
        # descriptor batches of the fused pack / unpack kernels (the pack side per source parity)
        n = len(self.links)
        self._packs = []
        for par in range(2):
            arr = (K.Rect * max(n, 1))()
            for i, L in enumerate(self.links):
                r0, c0, nr, ncb = L["strip"]
                arr[i] = K.Rect(self.mems[self.me].bufs[par], self.pitch, r0, c0, nr, ncb, L["send"].data_ptr())
            self._packs.append(arr)
        self._unpack = (K.Rect * max(n, 1))()
        for i, L in enumerate(self.links):
            r0, c0, nr, ncb = L["fill"]
            self._unpack[i] = K.Rect(L["ghost"].data_ptr(), L["gp"], r0, c0, nr, ncb, L["recv"].data_ptr())
        self._ops = []
        for L in self.links:
            self._ops.append(self.dist.P2POp(self.dist.isend, L["send"], L["rank"]))
            self._ops.append(self.dist.P2POp(self.dist.irecv, L["recv"], L["rank"]))
        if self._ops:     # connect the send/recv channels now (lazy in NCCL): a CUDA graph capture cannot
            for req in self.dist.batch_isend_irecv(self._ops):
                req.wait()
            torch.cuda.synchronize()

    def _cur_stream(self) -> int:
        return int(self.torch.cuda.current_stream().cuda_stream)

    def _exchange(self, src_parity: int) -> int:
        """One halo exchange of the current state; returns kernels launched."""
        st = self._cur_stream()
        n = len(self.links)
        if self.poison:
            for L in self.links:
                L["ghost"].fill_(0xA5)
        if n:
            K.rect_batch(self._packs[src_parity], n, True, st)
            for req in self.dist.batch_isend_irecv(self._ops):     # one ncclGroup
                req.wait()            # stream-ordered wait (NCCL): the next kernels see the data
            K.rect_batch(self._unpack, n, False, st)
        self.exchanges += 1
        return 2 if n else 0

    # -- compute --------------------------------------------------------------------------------------
    def _enqueue_eager(self, plan: StreamPlan, loops: int) -> int:
        launched = 0
        st = self._cur_stream()
        for k in plan.launches(loops):
            if self.world > 1:
                launched += self._exchange(self.parity)
            p = self.driver._params(plan, k, self.parity, 0, 0)
            if self.world > 1:
                for L in self.links:
                    p.src[L["d"]] = L["vbase"]
                    p.src_pitch[L["d"]] = L["gp"]
            K.stream_launch(p, st)
            launched += 1
            self.parity ^= 1
        return launched

    def enqueue(self, loops: int) -> tuple[StreamPlan | None, int]:
        if loops <= 0:
            return None, 0
        plan = self.plan(loops)
        if self.world > 1:
            self._build_ghosts(plan)
        if not (self.use_graph and self.world > 1 and not self.poison):
            return plan, self._enqueue_eager(plan, loops)
        # Whole run as ONE CUDA graph (captured per (loops, starting parity); NCCL send/recv groups are
        # capturable): replaying it costs one launch from the host.
        key = (loops, self.parity)
        if key not in self._graphs:
            torch = self.torch
            p0, x0 = self.parity, self.exchanges
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    launched = self._enqueue_eager(plan, loops)
                self._graphs[key] = (g, launched, self.parity, self.exchanges - x0)
            except Exception:                       # capture not possible here: stay eager
                self.use_graph = False
                self.parity, self.exchanges = p0, x0
                torch.cuda.synchronize()
                return plan, self._enqueue_eager(plan, loops)
            self.parity, self.exchanges = p0, x0    # capturing did not execute anything
        g, launched, parity_after, nx = self._graphs[key]
        g.replay()
        self.parity = parity_after
        self.exchanges += nx
        return plan, launched

    def run(self, loops: int, rep, sync: bool = True) -> None:
        torch = self.torch
        if loops > 0:
            K.kernel_prepare(self.filt.kind, self.grid.channels, self.plan(loops).R)
        if self.world > 1:
            torch.cuda.synchronize()
            self.dist.barrier()
        e0, e1 = K.Event(), K.Event()
        e0.record(self.stream)
        plan, launches = self.enqueue(loops)
        e1.record(self.stream)
        rep.launches = launches
        rep.plan = dict(plan.__dict__) if plan else {}
        rep.plan["nccl_exchanges"] = self.exchanges
        if sync:
            e1.synchronize()
            rep.elapsed_ms = e0.elapsed_ms(e1)
