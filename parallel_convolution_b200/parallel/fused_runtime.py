"""FusedRuntime: the per-process driver of the fused backend.

Unifies the two process models behind one interface used by ``ConvEngine``:

* single process — ``world == 1``: this process owns every tile of the grid as *virtual ranks* on
  one device (1 tile for a plain single-GPU run);
* one process per GPU — ``world > 1``: this process owns tile ``rank`` and maps its neighbours'
  buffers through CUDA IPC (``peer_memory.PeerGroup``).

A run is enqueued with ONE native call (``pconv_stream_run``): all launches of all local tiles go
onto one stream back to back; cross-rank ordering is entirely in-kernel (epoch flags). The host
synchronises twice per run (before: inputs in place; after: results readable), never in between.
"""
from __future__ import annotations

import os

import numpy as np

from ..models.filters import Filter
from ..ops import cuda as K
from .decomposition import TileGrid
from .fused import (RankMemory, TileDriver, make_plan, make_resident_plan,
                    tile_pitch)
from .plan import StreamPlan

# Whether "auto" lets the autotuner try the split-phase resident kernel (explicit --mode resident2
# always works). Single-GPU runs only until it has been validated across NVLink.
AUTO_RESIDENT2 = os.environ.get("PCONV_RESIDENT2", "0") == "1"


class FusedRuntime:
    def __init__(self, grid: TileGrid, filt: Filter, rank: int, world: int, dist,
                 halo_depth: int | None = None, mode: str = "auto", timeout_ms: float = 10_000.0,
                 band: int | None = None, slots: int = 1):
        self.grid, self.filt, self.rank, self.world, self.dist = grid, filt, rank, world, dist
        self.halo_depth, self.mode, self.band = halo_depth, mode, band
        self.pitch = tile_pitch(grid)
        self.slots, self.slot = max(1, int(slots)), 0
        self._pipe = None
        self._step = 0
        self.stream = K.stream_create()
        self.sms = K.sm_count()
        self.bands = K.band_heights(filt.kind)
        self.bands2 = K.resident2_band_heights(filt.kind)    # split-phase resident kernel
        self.bands3 = K.resident3_band_heights(filt.kind)    # ... with a protocol warp (experimental)
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
        self.cta_epoch = 0            # resident kernel: per-CTA exchange epochs (monotonic)
        self.multi = grid.workers > 1
        self._plans: dict[int, StreamPlan] = {}
        self._rplans: dict[int, StreamPlan | None] = {}
        self._r2plans: dict[int, StreamPlan | None] = {}
        self._r3plans: dict[int, StreamPlan | None] = {}
        # virtual ranks in resident mode need their persistent kernels co-resident: one stream each,
        # and no two of them on one hardware queue (ops.cuda.MAX_CONNECTIONS) — beyond that, stream mode
        self._rank_streams = ({r: K.stream_create() for r in self.local_ranks}
                              if 1 < len(self.local_ranks) < K.MAX_CONNECTIONS - 2 else {})

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

    def resident_plan(self, loops: int, dual: bool | int = False) -> StreamPlan | None:
        """Plan of the persistent kernel (``dual``: 1/True its split-phase variant, 2 the split-phase
        variant with a protocol warp); None if it cannot hold the tile."""
        dual = int(dual)
        cache = (self._rplans, self._r2plans, self._r3plans)[dual]
        if len(self.local_ranks) > 1 and not self._rank_streams:
            return None        # too many virtual ranks for concurrent persistent kernels
        if loops not in cache:
            nloc = len(self.local_ranks)
            bands = (self.bands, self.bands2, self.bands3)[dual]
            if dual and (not bands or (self.band and self.band not in bands)):
                cache[loops] = None
            else:
                cache[loops] = make_resident_plan(
                    self.grid, self.filt, loops, self.halo_depth, self.sms // nloc, bands, self.band,
                    peers=self.world > 1, dual=dual)
        return cache[loops]

    _VARIANT = {"stream": -1, "resident": 0, "resident2": 1, "resident3": 2}

    def _plan_cache(self, mode: str) -> dict:
        return {"stream": self._plans, "resident": self._rplans, "resident2": self._r2plans,
                "resident3": self._r3plans}[mode]

    def plan_for(self, mode: str, loops: int) -> StreamPlan | None:
        return self.plan(loops) if mode == "stream" else self.resident_plan(loops, self._VARIANT[mode])

    # -- on-device autotuning ----------------------------------------------------------------------
    def autotune(self, loops: int, max_candidates: int = 14, reps: int = 2) -> dict:
        """Measure candidate (mode, k, R) plans for ``loops`` iterations on this machine and keep the
        fastest. The model-based planner is good at ranking but blind to box-specific effects
        (launch gaps, cross-GPU skew); a few milliseconds of measurement settle it. Collective: every
        rank times the same candidates in the same order and the decision uses the max over ranks,
        so all ranks adopt the same plan. Overwrites the tile buffers with synthetic data — call it
        before loading real input."""
        if loops <= 0:
            return {}
        cands: list[tuple[str, StreamPlan]] = []

        def add(mode: str, k: int, R: int) -> None:
            try:
                if mode == "stream":
                    p = make_plan(self.grid, self.filt, loops, k, self.sms, self.bands, R)
                else:
                    var = self._VARIANT[mode]
                    p = make_resident_plan(self.grid, self.filt, loops, k, self.sms // len(self.local_ranks),
                                           (self.bands, self.bands2, self.bands3)[var], R,
                                           peers=self.world > 1, dual=max(var, 0))
            except ValueError:
                return
            if p is not None and (mode, p) not in cands:
                cands.append((mode, p))

        modes = [m for m in ("stream", "resident") if self.mode in ("auto", m)]
        if self.bands2 and (self.mode == "resident2" or (self.mode == "auto" and AUTO_RESIDENT2)):
            modes.append("resident2")
        if self.bands3 and self.mode == "resident3":      # never picked automatically (not run on hardware yet)
            modes.append("resident3")
        per_mode = max(4, max_candidates // max(1, len(modes)))
        for mode in modes:
            base = self.plan_for(mode, loops)
            if base is None:
                continue
            mine = len(cands)
            cands.append((mode, base))
            bands = {"resident2": self.bands2, "resident3": self.bands3}.get(mode, self.bands)
            for k in sorted({max(1, base.k - 2), max(1, base.k - 1), base.k + 1, base.k + 2, base.k + 5,
                             max(1, base.k // 2), base.k * 2} - {base.k}):
                if self.halo_depth or k > loops:
                    continue
                for R in ([base.R] + [b for b in bands if b != base.R]):
                    if self.band and R != self.band:
                        continue
                    add(mode, k, R)
            # rank this mode's candidates by the model and keep its most promising ones (the models of
            # different kernels are not comparable with each other — the measurement decides that)
            own = sorted(cands[mine:], key=lambda mp: mp[1].est_us_per_launch *
                         (-(-loops // mp[1].k) if mp[0] == "stream" else 1))
            cands[mine:] = own[:per_mode]
        saved = (self.mode, dict(self._plans), dict(self._rplans), dict(self._r2plans), dict(self._r3plans))
        results = []
        for mode, p in cands:
            self.mode = mode
            self._plan_cache(mode)[loops] = p
            best = None
            try:
                self.prepare(loops)
                self._barrier()
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
                self._barrier()
            except Exception:          # a candidate the kernels reject is simply not a candidate
                best = None
            if best is not None and self.world > 1:
                import torch
                t = torch.tensor([best], dtype=torch.float64, device="cuda")
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
                best = float(t.item())
            if best is not None:
                results.append((best, mode, p))
        self.mode, self._plans, self._rplans, self._r2plans, self._r3plans = saved
        if not results:
            return {}
        results.sort(key=lambda r: r[0])
        ms, mode, p = results[0]
        self._tuned = getattr(self, "_tuned", {})
        self._tuned[loops] = mode
        self._plan_cache(mode)[loops] = p
        return {"mode": mode, "plan": dict(p.__dict__), "ms": ms,
                "tried": [(round(r[0], 4), r[1], r[2].k, r[2].R) for r in results]}

    def choose_mode(self, loops: int) -> str:
        tuned = getattr(self, "_tuned", {}).get(loops)
        if tuned and self.mode == "auto":
            return tuned
        if self.mode in ("stream", "resident", "resident2", "resident3"):
            if self.mode != "stream" and self.plan_for(self.mode, loops) is None:
                raise ValueError(f"{self.mode} mode requested but the tile does not fit on chip")
            return self.mode
        # "auto", from measurements on B200 (profiles/README.md): on 1-2 GPUs one in-kernel exchange
        # costs about as much as a relaunch + reload, and streaming wins by 4-8 % (0.47 vs 0.51 ms on
        # one GPU, 0.31 vs 0.32 ms on two); from 4 GPUs up the per-launch cross-GPU skew dominates and
        # the persistent kernel wins (0.21 vs 0.26 ms on eight).
        if self.world >= 4 and self.resident_plan(loops) is not None:
            return "resident"
        return "stream"

    def _enqueue_resident(self, loops: int, variant: int = 0) -> tuple[StreamPlan, int]:
        plan = self.resident_plan(loops, variant)
        nblocks = -(-loops // plan.k)
        first_wait, self._first_wait = getattr(self, "_first_wait", 0), 0
        signal = self.epoch + 1 if self.multi else 0
        bank1, self._first_bank1 = getattr(self, "_first_bank1", False), False
        params = {r: self.drivers[r].resident_params(plan, loops, self.parity, first_wait if self.multi else 0,
                                                     signal, self.cta_epoch, slot=self.slot,
                                                     wait_bank1=bank1) for r in self.local_ranks}
        for p in params.values():
            p.dual = variant            # 0 single-phase, 1 split-phase, 2 split-phase + protocol warp
        if self._rank_streams:
            fork = K.Event().record(self.stream)
            joins = []
            for r, st in self._rank_streams.items():
                K.stream_wait_event(st, fork)
                K.resident_launch(params[r], st)
                joins.append(K.Event().record(st))
            for ev in joins:
                K.stream_wait_event(self.stream, ev)
            self._keep = (fork, joins)
        else:
            K.resident_launch(params[self.local_ranks[0]], self.stream)
        if self.multi:
            self.epoch += 1
        self.cta_epoch += nblocks
        self.parity ^= nblocks & 1
        return plan, len(self.local_ranks)

    def enqueue(self, loops: int) -> tuple[StreamPlan | None, int]:
        """Enqueue ``loops`` iterations on the runtime stream; returns (plan, kernel launches)."""
        if loops <= 0:
            return None, 0
        mode = self.choose_mode(loops)
        if mode != "stream":
            return self._enqueue_resident(loops, self._VARIANT[mode])
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
        resident = self._VARIANT[mode] + 1
        key = (plan.R, resident)
        if key not in getattr(self, "_prepared", set()):
            K.kernel_prepare(self.filt.kind, self.grid.channels, plan.R, resident)
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
