"""FusedRuntime: the per-process driver of the fused backend.

Unifies the two process models behind one interface used by ``ConvEngine``:

* single process — ``world == 1``: this process owns every tile of the grid as *virtual ranks* on
  one device (1 tile for a plain single-GPU run);
* one process per GPU — ``world > 1``: this process owns tile ``rank`` and maps its neighbours'
  buffers through CUDA IPC (``peer_memory.PeerGroup``).

Two kernels drive a run (``mode``): ``"resident"`` — ONE cooperative launch of the persistent kernel
(``persist_kernel.cuh``: the tile stays in registers, CTAs push their borders into each other's inbox
slots every k iterations, across GPUs through the peer-mapped exchange arenas) whenever the tile fits
on chip — and ``"stream"`` (``stream_kernel.cuh``: k fused iterations per launch over tiles of any
size, all launches of a run enqueued by one native call). Cross-rank ordering is entirely in-kernel
(epoch flags). The host synchronises twice per run (before: inputs in place; after: results readable),
never in between.
"""
from __future__ import annotations

import os

import numpy as np

from ..models.filters import Filter
from ..ops import cuda as K
from .decomposition import TileGrid
from .fused import (RankMemory, TileDriver, make_persist_plan, make_plan, tile_pitch)
from .plan import StreamPlan

MODES = ("stream", "resident")


class FusedRuntime:
    def __init__(self, grid: TileGrid, filt: Filter, rank: int, world: int, dist,
                 halo_depth: int | None = None, mode: str = "auto", timeout_ms: float = 10_000.0,
                 band: int | None = None, slots: int = 1, warps: int | None = None):
        if mode not in ("auto",) + MODES:
            raise ValueError(f"unknown kernel mode {mode!r} (auto | stream | resident)")
        if warps not in (None, 8, 16):
            raise ValueError("warps per CTA of the persistent kernel: 8 or 16")
        self.grid, self.filt, self.rank, self.world, self.dist = grid, filt, rank, world, dist
        self.halo_depth, self.mode, self.band, self.warps = halo_depth, mode, band, warps
        self.pitch = tile_pitch(grid)
        self.slots, self.slot = max(1, int(slots)), 0
        self._pipe = None
        self._step = 0
        self.stream = K.stream_create()
        self.sms = K.sm_count()
        self.bands = K.band_heights(filt.kind)
        self.pbands = K.persist_band_heights(filt.kind)
        rows = grid.max_tile_rows()
        self.peers = None
        if world > 1:
            if grid.workers != world:
                raise ValueError("one tile per process: grid size must equal the world size")
            from .peer_memory import PeerGroup
            self.local_ranks = [rank]
            mem = RankMemory.allocate(rows, self.pitch, self.slots)
            self.mems = {rank: mem}
            self.peers = PeerGroup(rank, world)
            nbrs = {rank: self.peers.exchange(mem, grid.neighbors(rank))}
        else:
            self.local_ranks = list(range(grid.workers))
            self.mems = {r: RankMemory.allocate(rows, self.pitch, self.slots) for r in self.local_ranks}
            nbrs = {r: [self.mems[n] if n >= 0 else None for n in grid.neighbors(r)]
                    for r in self.local_ranks}
        self.drivers = {r: TileDriver(grid, r, filt, self.mems[r], nbrs[r], self.pitch, self.stream,
                                      timeout_ms) for r in self.local_ranks}
        self.parity = 0
        self.epoch = 0
        self.cta_epoch = 0            # persistent kernel: per-CTA exchange epochs (monotonic, contiguous)
        self.multi = grid.workers > 1
        self._plans: dict[int, StreamPlan] = {}
        self._pplans: dict[int, StreamPlan | None] = {}
        self._tuned: dict[int, str] = {}
        # Virtual ranks in resident mode need their persistent kernels co-resident: one stream each, and
        # no two of them on one hardware queue. CUDA multiplexes streams onto CUDA_DEVICE_MAX_CONNECTIONS
        # queues (read when the context is created); beyond what is usable, stream mode only.
        self._rank_streams = ({r: K.stream_create() for r in self.local_ranks}
                              if 1 < len(self.local_ranks) <= K.usable_connections() - 2 else {})

    # -- lifecycle ------------------------------------------------------------------------------
    def close(self) -> None:
        self.pipeline_sync()
        if self.peers is not None:
            K.device_sync()
            self.peers.barrier()      # nobody may still be reading my buffers
            self.peers.close()
        for m in self.mems.values():
            m.free()
        self.mems = {}
        for st in self._rank_streams.values():
            K.stream_destroy(st)
        if self._pipe is not None:
            K.stream_destroy(self._pipe["s_in"])
            K.stream_destroy(self._pipe["s_out"])
        K.stream_destroy(self.stream)

    def _barrier(self) -> None:
        if self.peers is not None:
            K.stream_sync(self.stream)
            self.peers.barrier()

    # -- data -----------------------------------------------------------------------------------
    def upload(self, tiles: dict[int, np.ndarray]) -> None:
        self._barrier()               # previous run's readers are done with my buffers
        for r, tile in tiles.items():
            self.drivers[r].upload(tile, 0)
        self.parity, self.slot = 0, 0

    def synth(self, seed: int) -> None:
        self._barrier()
        for d in self.drivers.values():
            d.synth(seed, 0)
        self.parity, self.slot = 0, 0

    def download(self) -> dict[int, np.ndarray]:
        return {r: d.download(2 * self.slot + self.parity) for r, d in self.drivers.items()}

    def buffer_ptr(self, rank: int, parity: int | None = None) -> int:
        return self.mems[rank].bufs[2 * self.slot + (self.parity if parity is None else parity)]

    # -- compute ----------------------------------------------------------------------------------
    def plan(self, loops: int) -> StreamPlan:
        if loops not in self._plans:
            self._plans[loops] = make_plan(self.grid, self.filt, loops, self.halo_depth, self.sms,
                                           self.bands, self.band)
        return self._plans[loops]

    # -- in-stream input hand-over (no host barrier between back-to-back steps) -------------------
    def begin_input(self) -> int:
        """Enqueue: wait until every neighbour finished the previous run (they may still be reading
        my buffers). Call before overwriting a tile buffer in-stream. Returns kernels launched."""
        if not self.multi or self.epoch == 0:
            return 0
        if getattr(self, "_last_exchanges", 0) >= 1:
            # The previous run was a persistent launch with at least one in-kernel exchange: a neighbour reads
            # my input buffer only in its initial region load, BEFORE it pushes its first halo, and my kernel
            # could not finish without receiving that push — so every neighbour is done with the buffer by
            # the time this (stream-ordered) point is reached. No flag round trip needed.
            return 0
        for d in self.drivers.values():
            K.sync_launch(d.sync_params(self.epoch, 0), self.stream)
        return len(self.drivers)

    def end_input(self) -> int:
        """Enqueue: announce "my input tile is in place"; the next run's first launch waits for it."""
        if not self.multi:
            return 0
        self.epoch += 1
        for d in self.drivers.values():
            K.sync_launch(d.sync_params(0, self.epoch), self.stream)
        self._first_wait = self.epoch
        return len(self.drivers)

    def resident_plan(self, loops: int) -> StreamPlan | None:
        """Plan of the persistent kernel (the cheaper of 8 / 16 warps per CTA by the model, unless
        ``warps`` pins it); None if the tile does not fit on chip or the geometry is not supported."""
        if len(self.local_ranks) > 1 and not self._rank_streams:
            return None        # too many virtual ranks for concurrent persistent kernels
        if loops not in self._pplans:
            best = None
            for w in ((self.warps,) if self.warps else (8, 16)):
                p = self._persist_candidate(loops, self.halo_depth, self.band, w)
                if p is not None and (best is None or p.est_us_per_launch < best.est_us_per_launch):
                    best = p
            self._pplans[loops] = best
        return self._pplans[loops]

    def _persist_candidate(self, loops: int, k: int | None, R: int | None, warps: int) -> StreamPlan | None:
        nloc = len(self.local_ranks)
        try:
            p = make_persist_plan(self.grid, self.filt, loops, k, self.sms // nloc, self.pbands, R,
                                  peers=self.world > 1, warps=warps)
        except ValueError:
            return None
        if p is None:
            return None
        # the device must really keep that many CTAs resident (registers / shared memory of the instance)
        try:
            d = self.drivers[self.local_ranks[0]]
            cap = K.persist_capacity(d.persist_params(p, loops, 0, 0, 0, 0))
        except K.CudaError:
            return None
        return p if p.nty * p.ntx * nloc <= cap else None

    def _plan_cache(self, mode: str) -> dict:
        return {"stream": self._plans, "resident": self._pplans}[mode]

    def plan_for(self, mode: str, loops: int) -> StreamPlan | None:
        return self.plan(loops) if mode == "stream" else self.resident_plan(loops)

    # -- on-device autotuning ----------------------------------------------------------------------
    def autotune(self, loops: int, max_candidates: int = 16, reps: int = 2) -> dict:
        """Measure candidate (mode, warps, k, R) plans for ``loops`` iterations on this machine and keep
        the fastest. The model-based planner is good at ranking but blind to box-specific effects
        (launch gaps, cross-GPU skew); a few milliseconds of measurement settle it. Collective: every
        rank times the same candidates in the same order, executes the same collectives whether or not
        a candidate failed locally, and the decision uses the max over ranks — so all ranks adopt the
        same plan. Overwrites the tile buffers with synthetic data: call it before loading real input."""
        if loops <= 0:
            return {}
        cands: list[tuple[str, StreamPlan]] = []

        def add(mode: str, k: int | None, R: int | None, warps: int = 16) -> None:
            try:
                if mode == "stream":
                    p = make_plan(self.grid, self.filt, loops, k, self.sms, self.bands, R)
                else:
                    p = self._persist_candidate(loops, k, R, warps)
            except ValueError:
                return
            if p is not None and (mode, p) not in cands:
                cands.append((mode, p))

        modes = [m for m in MODES if self.mode in ("auto", m)]
        per_mode = max(4, max_candidates // max(1, len(modes)))
        for mode in modes:
            mine = len(cands)
            variants = (16,) if mode == "stream" else ((self.warps,) if self.warps else (8, 16))
            bands = self.bands if mode == "stream" else self.pbands
            for w in variants:
                n0 = len(cands)
                add(mode, self.halo_depth, self.band, w)
                if len(cands) == n0:
                    continue
                base = cands[n0][1]
                for k in sorted({max(1, base.k - 2), max(1, base.k - 1), base.k + 1, base.k + 2, base.k + 4,
                                 max(1, base.k // 2), base.k * 2} - {base.k}):
                    if self.halo_depth or k > loops:
                        continue
                    add(mode, k, self.band, w)          # the model's best band height for that depth
                    if not self.band:
                        add(mode, k, base.R, w)
            # rank this mode's candidates by the model and keep its most promising ones (the models of
            # the two kernels are not comparable with each other — the measurement decides that)
            own = sorted(cands[mine:], key=lambda mp: mp[1].est_us_per_launch *
                         (-(-loops // mp[1].k) if mp[0] == "stream" else 1))
            cands[mine:] = own[:per_mode]
        saved = (self.mode, dict(self._plans), dict(self._pplans), dict(self._tuned))
        self._tuned.pop(loops, None)
        results = []
        for mode, p in cands:
            self.mode = mode
            self._plan_cache(mode)[loops] = p
            best = None
            try:
                self.prepare(loops)
                for d in self.drivers.values():
                    d.reset_status()
                K.stream_sync(self.stream)
                for trial in range(reps + 1):           # trial 0 warms up
                    steps = 1 if trial == 0 else 4
                    e0, e1 = K.Event(), K.Event()
                    e0.record(self.stream)
                    for _ in range(steps):               # back-to-back steps, ordered in-stream like
                        self.begin_input()               # the streaming/bench path (no host barrier)
                        for d in self.drivers.values():
                            d.synth(17, 2 * self.slot)
                        self.parity = 0
                        self.end_input()
                        self.enqueue(loops)
                    e1.record(self.stream)
                    e1.synchronize()
                    if trial:
                        ms = e0.elapsed_ms(e1) / steps
                        best = ms if best is None else min(best, ms)
                if any(d.read_status()[0] for d in self.drivers.values()):
                    best = None                          # a watchdog fired: not a candidate
            except Exception:          # a candidate the kernels reject is simply not a candidate
                best = None
            # the same collectives on every rank, failed or not
            self._barrier()
            if self.world > 1:
                import torch
                t = torch.tensor([best if best is not None else float("inf")], dtype=torch.float64, device="cuda")
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
                best = float(t.item())
                if best == float("inf"):
                    best = None
            if best is not None:
                results.append((best, mode, p))
        self.mode, self._plans, self._pplans, self._tuned = saved
        if not results:
            return {}
        results.sort(key=lambda r: r[0])
        ms, mode, p = results[0]
        if mode == "stream":
            # The tuner's four-step measurement cannot resolve differences below ~2 %. Within that band prefer
            # the persistent kernel: one launch per run instead of several, so less host work and launch
            # jitter per step in a long sequence of steps (which is how it is then used).
            for r in results[1:]:
                if r[1] == "resident" and r[0] <= 1.02 * ms:
                    ms, mode, p = r
                    break
        self._tuned[loops] = mode
        self._plan_cache(mode)[loops] = p
        return {"mode": mode, "plan": dict(p.__dict__), "ms": ms,
                "tried": [(round(r[0], 4), r[1], r[2].W if r[1] != "stream" else 0, r[2].k, r[2].R) for r in results]}

    def choose_mode(self, loops: int) -> str:
        tuned = self._tuned.get(loops)
        if tuned and self.mode == "auto":
            return tuned
        if self.mode in MODES:
            if self.mode != "stream" and self.plan_for(self.mode, loops) is None:
                raise ValueError("resident mode requested but the tile does not fit on chip "
                                 "(or its geometry is not supported by the persistent kernel)")
            return self.mode
        # "auto": the persistent kernel whenever the tile fits on chip — one launch, the image never
        # leaves the registers, exchanges hidden behind a co-resident CTA (profiles/README.md)
        return "resident" if self.resident_plan(loops) is not None else "stream"

    def _enqueue_resident(self, loops: int) -> tuple[StreamPlan, int]:
        plan = self.resident_plan(loops)
        nblocks = -(-loops // plan.k)
        first_wait, self._first_wait = getattr(self, "_first_wait", 0), 0
        signal = self.epoch + 1 if self.multi else 0
        bank1, self._first_bank1 = getattr(self, "_first_bank1", False), False
        # the launch structs are cached per (plan, loops, buffers, flag bank): only the epochs change per run
        ckey = (plan, loops, self.parity, self.slot, bank1)
        cache = self.__dict__.setdefault("_pparams", {})
        if ckey not in cache:
            if len(cache) > 64:
                cache.clear()
            cache[ckey] = {r: self.drivers[r].persist_params(plan, loops, self.parity, 0, 0, 0, slot=self.slot,
                                                             wait_bank1=bank1, peer_scope=self.world > 1)
                           for r in self.local_ranks}
        params = cache[ckey]
        for r, p in params.items():
            p.sync.inject_drop_signal = 1 if self.drivers[r].inject_drop_signal else 0
            p.sync.wait_value = first_wait if self.multi else 0
            p.sync.signal_value = signal
            p.epoch_base = self.cta_epoch & 0xFFFFFFFF
        if self._rank_streams:
            fork = K.Event().record(self.stream)
            joins = []
            for r, st in self._rank_streams.items():
                K.stream_wait_event(st, fork)
                K.persist_launch(params[r], st)
                joins.append(K.Event().record(st))
            for ev in joins:
                K.stream_wait_event(self.stream, ev)
            self._keep = (fork, joins)
        else:
            K.persist_launch(params[self.local_ranks[0]], self.stream)
        if self.multi:
            self.epoch += 1
        self.cta_epoch += nblocks - 1     # contiguous epochs: the inbox parity keeps alternating
        self._last_exchanges = nblocks - 1
        self.parity ^= 1                  # reads one buffer of the pair, writes the other
        return plan, len(self.local_ranks)

    def enqueue(self, loops: int) -> tuple[StreamPlan | None, int]:
        """Enqueue ``loops`` iterations on the runtime stream; returns (plan, kernel launches)."""
        if loops <= 0:
            return None, 0
        mode = self.choose_mode(loops)
        if mode != "stream":
            return self._enqueue_resident(loops)
        plan = self.plan(loops)
        ks = plan.launches(loops)
        n = len(self.local_ranks)
        blocks = (K.StreamParams * (2 * n))()
        for i, r in enumerate(self.local_ranks):
            d = self.drivers[r]
            blocks[2 * i] = d._params(plan, plan.k, self.parity, 0, 0, slot=self.slot)
            blocks[2 * i + 1] = d._params(plan, plan.k, self.parity ^ 1, 0, 0, slot=self.slot)
        first_wait, self._first_wait = getattr(self, "_first_wait", 0), 0
        bank1, self._first_bank1 = getattr(self, "_first_bank1", False), False
        from .fused import BANK1
        ff = [self.mems[r].flags + BANK1 for r in self.local_ranks] if (bank1 and first_wait) else None
        K.stream_run(blocks, n, ks, self.epoch, self.multi, self.stream, first_wait, ff)
        self._last_exchanges = 0
        self.epoch += len(ks)
        self.parity ^= len(ks) & 1
        return plan, len(ks) * n

    # -- pipelined end-to-end steps: H2D(i+1) || iterate(i) || D2H(i-1) ---------------------------------
    def _pipeline(self):
        if self._pipe is None:
            S = self.slots
            self._pipe = {"s_in": K.stream_create(), "s_out": K.stream_create(),
                          "in_done": [K.Event() for _ in range(S)], "comp_done": [K.Event() for _ in range(S)],
                          "out_done": [K.Event() for _ in range(S)], "used": [False] * S,
                          "last_epoch": [0] * S}
        return self._pipe

    def enqueue_step_pipelined(self, host_in: dict, host_out: dict, loops: int) -> int:
        """One end-to-end step on buffer slot ``step % slots``: pinned host -> device on the copy-in
        stream, ``loops`` iterations on the compute stream, device -> pinned host on the copy-out
        stream. Steps on different slots overlap; cross-rank ordering is in-stream (flag bank 1 =
        "input of step s is in place", bank 0 = launch epochs). ``host_in[r]`` / ``host_out[r]``:
        pinned 2-D arrays of local tile r (one set per slot is the caller's business)."""
        P, S = self._pipeline(), self.slots
        step = self._step
        slot = step % S
        self._step += 1
        n = 0
        s_in, s_out = P["s_in"], P["s_out"]
        if P["used"][slot]:
            K.stream_wait_event(s_in, P["out_done"][slot])          # my previous result left the slot
            if self.multi and P["last_epoch"][slot]:                # neighbours finished reading it
                for d in self.drivers.values():
                    K.sync_launch(d.sync_params(P["last_epoch"][slot], 0, bank=0), s_in)
                n += len(self.drivers)
        for r in self.local_ranks:
            a = host_in[r]
            K.memcpy2d_async(self.mems[r].bufs[2 * slot], self.pitch, a.ctypes.data, a.shape[1],
                             a.shape[1], a.shape[0], K.H2D, s_in)
        if self.multi:
            for d in self.drivers.values():
                K.sync_launch(d.sync_params(0, step + 1, bank=1), s_in)
            n += len(self.drivers)
        P["in_done"][slot].record(s_in)
        K.stream_wait_event(self.stream, P["in_done"][slot])
        self.slot, self.parity = slot, 0
        if self.multi:
            self._first_wait, self._first_bank1 = step + 1, True
        _, launches = self.enqueue(loops)
        P["last_epoch"][slot] = self.epoch
        P["used"][slot] = True
        P["comp_done"][slot].record(self.stream)
        K.stream_wait_event(s_out, P["comp_done"][slot])
        for r in self.local_ranks:
            o = host_out[r]
            K.memcpy2d_async(o.ctypes.data, o.shape[1], self.mems[r].bufs[2 * slot + self.parity],
                             self.pitch, o.shape[1], o.shape[0], K.D2H, s_out)
        P["out_done"][slot].record(s_out)
        return n + launches

    def slot_sync(self, slot: int) -> None:
        """Block until the pipelined step that last used buffer slot ``slot`` has delivered its result to the
        host (its D2H copy is complete) — finer than ``pipeline_sync``: later steps keep running."""
        if self._pipe is not None and self._pipe["used"][slot]:
            self._pipe["out_done"][slot].synchronize()

    def pipeline_sync(self) -> None:
        K.stream_sync(self.stream)
        if self._pipe is not None:
            K.stream_sync(self._pipe["s_in"])
            K.stream_sync(self._pipe["s_out"])

    def prepare(self, loops: int) -> None:
        """Plan and force the module load of the kernel that ``loops`` will use (one-time cost)."""
        if loops <= 0:
            return
        mode = self.choose_mode(loops)
        plan = self.plan_for(mode, loops)
        key = (plan.R, plan.W if mode != "stream" else 0)
        if key not in getattr(self, "_prepared", set()):
            if mode == "stream":
                K.kernel_prepare(self.filt.kind, self.grid.channels, plan.R)
            else:
                K.persist_capacity(self.drivers[self.local_ranks[0]].persist_params(plan, loops, 0, 0, 0, 0))
            self._prepared = getattr(self, "_prepared", set()) | {key}

    def run(self, loops: int, rep, sync: bool = True) -> None:
        self.prepare(loops)
        for d in self.drivers.values():
            d.reset_status()
        self._barrier()               # every rank's input tile is in place
        e0, e1 = K.Event(), K.Event()
        e0.record(self.stream)
        plan, launches = self.enqueue(loops)
        e1.record(self.stream)
        rep.launches = launches
        rep.plan = dict(plan.__dict__) if plan else {}
        if not sync:
            self._pending = (e0, e1)
            return
        e1.synchronize()
        rep.elapsed_ms = e0.elapsed_ms(e1)
        for d in self.drivers.values():
            err, mx, sm, nw = d.read_status()
            rep.watchdog |= err
            rep.max_wait_us = max(rep.max_wait_us, mx)
            rep.sum_wait_us += sm
            rep.waits += nw
            if getattr(d, "probe", None) and d.probe.get("exchanges"):
                rep.probe = d.probe
        self._barrier()
        if rep.watchdog:
            raise RuntimeError("halo watchdog fired: a neighbouring rank never signalled "
                               f"(timeout {self.drivers[self.local_ranks[0]].timeout_ns / 1e6:.0f} ms)")
