"""ConvEngine — the public API: iterate a 3x3 filter over a raw image on 1..N workers.

One engine instance == one rank. Process model:

* **single process** (no ``torch.distributed``): ``backend="fused"`` runs on one GPU; with
  ``workers=N`` the image is decomposed into N *virtual ranks* on that GPU (same kernels, flags and
  owner lookup as the multi-GPU path). ``backend="cpu"`` runs the native CPU engine.
* **one process per GPU** (``torchrun``; ``torch.distributed`` initialised, backend nccl or gloo):
  rank r owns tile r of the grid. ``fused`` pulls halos through NVLink peer memory inside the
  kernel; ``nccl`` is the send/recv baseline; ``cpu`` is the MPI-like host backend over gloo.

Typical use (the call a user makes; ``bench.py`` times exactly this)::

    eng = ConvEngine(width=1920, height=2520, channels=3, filter="gaussian")
    out = eng.run(image, loops=100)          # numpy uint8 in, numpy uint8 out (rank 0)

Replaces the reference's three ``main`` programs (``mpi/mpi_conv.c:18-297``,
``mpi_omp/mpi_omp_conv.c``, ``cuda/main.c:10-55`` + ``gpuConvolute``).
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass, field

import numpy as np

from .models.filters import Filter
from .ops import host as H
from .parallel.decomposition import TileGrid


@dataclass
class EngineConfig:
    width: int
    height: int
    channels: int = 1
    filter: str | Filter = "gaussian"
    backend: str = "auto"            # auto | fused | nccl | cpu
    workers: int | None = None       # single-process: number of virtual ranks (default 1)
    grid: tuple[int, int] | None = None
    halo_depth: int | None = None    # temporal-blocking depth k (None: planner decides)
    mode: str = "auto"               # fused backend: auto | stream | resident (persistent kernel)
    warps: int | None = None         # persistent kernel: warps per CTA (8 | 16; None: planner/autotuner)
    threads: int = 0                 # cpu backend: OpenMP threads per rank
    cpu_transport: str = "auto"      # cpu backend, several ranks: auto | shm (shared image, one node) | tcp (native messages) | gloo
    device: int | None = None        # CUDA device index (default: LOCAL_RANK or 0)
    timeout_ms: float = 10_000.0     # in-kernel watchdog for neighbour flags
    slots: int = 1                   # buffer sets for pipelined streaming steps (H2D || iterate || D2H)
    standalone: bool = False         # ignore torch.distributed: a private single-process engine inside a rank


@dataclass
class RunReport:
    loops: int = 0
    elapsed_ms: float = 0.0          # device-timed (fused/nccl) or host-timed (cpu), max over ranks
    launches: int = 0
    backend: str = ""
    plan: dict = field(default_factory=dict)
    max_wait_us: float = 0.0         # exposed halo wait: longest single spin of any CTA
    sum_wait_us: float = 0.0
    waits: int = 0
    watchdog: bool = False
    probe: dict = field(default_factory=dict)   # resident kernel: in-kernel phase times of one interior CTA

    def mpixels_per_s(self, width: int, height: int) -> float:
        return width * height * self.loops / max(self.elapsed_ms, 1e-9) / 1e3


def _dist():
    import sys
    if "torch.distributed" not in sys.modules:
        return None      # never imported => cannot be initialised; keeps CPU/CLI start-up torch-free
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


class ConvEngine:
    def __init__(self, width: int | None = None, height: int | None = None, channels: int = 1,
                 filter: str | Filter = "gaussian", config: EngineConfig | None = None, **kw):
        cfg = config or EngineConfig(width=width, height=height, channels=channels, filter=filter, **kw)
        self.cfg = cfg
        self.filter = cfg.filter if isinstance(cfg.filter, Filter) else Filter.parse(str(cfg.filter))
        self.dist = None if cfg.standalone else _dist()
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        self.distributed = self.world > 1
        self.backend = self._pick_backend(cfg.backend)
        workers = self.world if self.distributed else (cfg.workers or 1)
        if self.backend == "cpu" and not self.distributed:
            workers = 1
        self.grid = TileGrid.create(cfg.height, cfg.width, cfg.channels, workers, cfg.grid)
        self.report = RunReport(backend=self.backend)
        self._impl = None
        self._setup()

    # ------------------------------------------------------------------------------------------
    def _pick_backend(self, want: str) -> str:
        want = {"host": "cpu", "gpu": "fused", "cuda": "fused"}.get(want, want)
        if want not in ("auto", "fused", "nccl", "cpu"):
            raise ValueError(f"unknown backend {want!r}")
        if want == "auto":
            from .ops import cuda as K
            if K.available():
                return "fused"
            if os.path.exists("/dev/nvidiactl") and os.environ.get("PCONV_ALLOW_CPU_FALLBACK") != "1":
                # a GPU box must never silently run the CPU engine: say why the CUDA path is unusable
                try:
                    K.lib()
                    why = "libpconv_cuda.so loaded but reports no usable device"
                except Exception as e:  # noqa: BLE001
                    why = f"libpconv_cuda.so could not be built/loaded: {e}"
                raise RuntimeError(f"NVIDIA driver present but the fused backend is unavailable ({why}); "
                                   "pass --backend cpu (or PCONV_ALLOW_CPU_FALLBACK=1) to run on the CPU")
            return "cpu"
        if want in ("fused", "nccl"):
            from .ops import cuda as K
            if not K.available():
                raise RuntimeError(f"backend {want!r} needs a CUDA device and libpconv_cuda.so")
        return want

    def _pick_cpu_transport(self, want: str) -> str:
        """How CPU ranks trade halos: through a shared image when they all sit on one node (no messages, see
        parallel/shm_backend.py); otherwise private tiles and messages with the reference's MPI schedule — over the
        native TCP mesh (parallel/net_backend.py), or over gloo driven from Python (parallel/host_backend.py)."""
        want = os.environ.get("PCONV_CPU_TRANSPORT", want or "auto")
        if want not in ("auto", "shm", "tcp", "gloo"):
            raise ValueError(f"unknown cpu transport {want!r}")
        if not self.distributed:
            return "local"
        if want in ("gloo", "tcp"):
            return want
        from .parallel.shm_backend import same_node, segment_bytes
        ok = same_node(self.dist, segment_bytes(self.grid))   # collective: every rank takes the same branch
        if want == "shm" and not ok:
            raise RuntimeError("cpu transport 'shm' needs every rank on one node and room for both images in /dev/shm")
        return "shm" if ok else "tcp"

    def _setup(self) -> None:
        cfg = self.cfg
        if self.backend == "cpu":
            self.cpu_transport = self._pick_cpu_transport(cfg.cpu_transport)
            if self.cpu_transport == "shm":
                from .parallel.shm_backend import ShmTile
                self._impl = ShmTile(self.grid, self.rank, self.filter, cfg.halo_depth, cfg.threads, self.dist,
                                     timeout_ms=cfg.timeout_ms)
                return
            if self.cpu_transport == "tcp":
                from .parallel.net_backend import NetTile
                self._impl = NetTile(self.grid, self.rank, self.filter, cfg.halo_depth, cfg.threads, self.dist,
                                     timeout_ms=max(cfg.timeout_ms, 30_000.0))
                return
            from .parallel.host_backend import HostTile
            # messages are dear (gloo: 0.5-2 ms per exchange of eight): by default trade them for a 6-deep halo and
            # its redundant ring; --halo-depth 1 is the reference's schedule (one exchange per iteration)
            k = cfg.halo_depth or (max(1, min(6, self.grid.min_tile_rows(), self.grid.min_tile_cols()))
                                   if self.distributed else 1)
            self._impl = HostTile(self.grid, self.rank if self.distributed else 0, self.filter, k,
                                  cfg.threads, self.dist if self.distributed else None)
            return
        from .ops import cuda as K
        dev = cfg.device if cfg.device is not None else int(os.environ.get("LOCAL_RANK", "0"))
        if dev >= K.device_count():
            dev = dev % max(K.device_count(), 1)
        K.set_device(dev)
        self.device = dev
        if self.backend == "fused":
            from .parallel.fused_runtime import FusedRuntime
            self._impl = FusedRuntime(self.grid, self.filter, self.rank, self.world, self.dist,
                                      halo_depth=cfg.halo_depth, mode=cfg.mode,
                                      timeout_ms=cfg.timeout_ms, slots=cfg.slots, warps=cfg.warps)
        else:
            from .parallel.nccl_backend import NcclRuntime
            self._impl = NcclRuntime(self.grid, self.filter, self.rank, self.world, self.dist,
                                     halo_depth=cfg.halo_depth)

    def close(self) -> None:
        if self._impl is not None and hasattr(self._impl, "close"):
            self._impl.close()          # synchronises every stream first
        if self.backend != "cpu" and self._impl is not None:
            from .ops import cuda as K
            for bufs in getattr(self, "_pool", []):
                for p in bufs.values():
                    K.free(p)
            for p in getattr(self, "_pinned", []):
                K.host_free(p)
            self._pool, self._pinned = [], []
        self._impl = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- tiles this process owns -------------------------------------------------------------------
    @property
    def local_tiles(self):
        if self.distributed:
            return [self.grid.tile(self.rank)]
        return self.grid.tiles()

    def _slice(self, img2d: np.ndarray, t) -> np.ndarray:
        C = self.grid.channels
        return img2d[t.row0:t.row0 + t.rows, t.col0 * C:(t.col0 + t.cols) * C]

    # -- data in -------------------------------------------------------------------------------------
    def load_image(self, image: np.ndarray) -> None:
        """Every rank passes the full image (or at least its own tile region is read from it)."""
        g = self.grid
        img = np.asarray(image, dtype=np.uint8).reshape(g.height, g.row_bytes)
        if self.backend == "cpu":
            self._impl.load(self._slice(img, self._impl.tile))
        else:
            self._impl.upload({t.rank: self._slice(img, t) for t in self.local_tiles})

    def load_file(self, path: str) -> None:
        """Rank-parallel read: every rank reads exactly its tile rows (64-bit offsets)."""
        g, C = self.grid, self.grid.channels
        need = g.height * g.row_bytes
        have = H.file_size(path)
        if have < need:
            raise ValueError(f"{path}: {have} bytes, need {need} for {g.width}x{g.height}x{C}")
        if self.backend != "cpu" and hasattr(self._impl, "mems"):
            return self._load_file_pipelined(path)
        tiles = {t.rank: H.read_rect(path, g.row_bytes, t.row0, t.rows, t.col0 * C, t.cols * C)
                 for t in self.local_tiles}
        if self.backend == "cpu":
            self._impl.load(tiles[self._impl.tile.rank])
        else:
            self._impl.upload(tiles)

    # -- overlapped rank-parallel file I/O of the GPU backends ------------------------------------------
    # The reference's ranks seek/read (write) their own tile rows concurrently (mpi/mpi_conv.c:133-147,
    # 253-271); here every rank additionally overlaps its pread/pwrite with its PCIe copies: the tile moves
    # in row chunks through two pinned staging buffers, the disk side of chunk j+1 running while the copy
    # engine moves chunk j.
    _IO_CHUNK_BYTES = 8 << 20

    def _io_staging(self, tile_cb: int):
        import ctypes
        from .ops import cuda as K
        rows = max(16, self._IO_CHUNK_BYTES // max(tile_cb, 1))
        ptrs = [K.host_alloc(rows * tile_cb) for _ in range(2)]
        bufs = [np.ctypeslib.as_array((ctypes.c_uint8 * (rows * tile_cb)).from_address(p)).reshape(rows, tile_cb)
                for p in ptrs]
        return rows, ptrs, bufs

    def _load_file_pipelined(self, path: str) -> None:
        from .ops import cuda as K
        g, C, rt = self.grid, self.grid.channels, self._impl
        if hasattr(rt, "_barrier"):
            rt._barrier()                 # previous run's readers are done with my buffers
        for t in self.local_tiles:
            cb = t.cols * C
            rows, ptrs, bufs = self._io_staging(cb)
            ev = [K.Event(), K.Event()]
            try:
                for j, r0 in enumerate(range(0, t.rows, rows)):
                    b, n = j & 1, min(rows, t.rows - r0)
                    if j >= 2:
                        ev[b].synchronize()                  # the copy out of this staging buffer is done
                    H.read_rect(path, g.row_bytes, t.row0 + r0, n, t.col0 * C, cb, out=bufs[b][:n])
                    K.memcpy2d_async(rt.mems[t.rank].bufs[0] + r0 * rt.pitch, rt.pitch, ptrs[b], cb, cb, n,
                                     K.H2D, rt.stream)
                    ev[b].record(rt.stream)
                K.stream_sync(rt.stream)
            finally:
                for p in ptrs:
                    K.host_free(p)
        rt.parity = 0
        if hasattr(rt, "slot"):
            rt.slot = 0

    def _save_file_pipelined(self, path: str) -> None:
        from .ops import cuda as K
        g, C, rt = self.grid, self.grid.channels, self._impl
        K.stream_sync(rt.stream)
        slot = getattr(rt, "slot", 0)
        for t in self.local_tiles:
            cb = t.cols * C
            rows, ptrs, bufs = self._io_staging(cb)
            ev = [K.Event(), K.Event()]
            src = rt.mems[t.rank].bufs[2 * slot + rt.parity]
            chunks = [(r0, min(rows, t.rows - r0)) for r0 in range(0, t.rows, rows)]
            try:
                for j, (r0, n) in enumerate(chunks):
                    b = j & 1
                    K.memcpy2d_async(ptrs[b], cb, src + r0 * rt.pitch, rt.pitch, cb, n, K.D2H, rt.stream)
                    ev[b].record(rt.stream)
                    if j >= 1:                               # write chunk j-1 while chunk j comes down
                        pr0, pn = chunks[j - 1]
                        ev[1 - b].synchronize()
                        H.write_rect(path, g.row_bytes, t.row0 + pr0, t.col0 * C, bufs[1 - b][:pn])
                if chunks:
                    lr0, ln = chunks[-1]
                    ev[(len(chunks) - 1) & 1].synchronize()
                    H.write_rect(path, g.row_bytes, t.row0 + lr0, t.col0 * C, bufs[(len(chunks) - 1) & 1][:ln])
            finally:
                K.stream_sync(rt.stream)
                for p in ptrs:
                    K.host_free(p)

    def load_synthetic(self, seed: int = 1) -> None:
        """Deterministic random image, generated per tile (on the device for GPU backends)."""
        g, C = self.grid, self.grid.channels
        if self.backend == "cpu":
            t = self._impl.tile
            self._impl.load(H.synth_rect(seed, g.row_bytes, t.row0, t.rows, t.col0 * C, t.cols * C))
        else:
            self._impl.synth(seed)

    # -- compute -------------------------------------------------------------------------------------
    def autotune(self, loops: int) -> dict:
        """Fused backend: measure candidate launch plans for ``loops`` iterations on this machine
        and adopt the fastest (a few ms; overwrites the loaded tiles — call before loading input)."""
        if self.backend == "fused" and hasattr(self._impl, "autotune"):
            return self._impl.autotune(loops)
        return {}

    def iterate(self, loops: int, sync: bool = True) -> RunReport:
        """Apply the filter ``loops`` times to the loaded tiles. Device-timed for GPU backends."""
        rep = RunReport(loops=loops, backend=self.backend)
        if self.backend == "cpu":
            if self.distributed:
                self.dist.barrier()
            t0 = time.perf_counter()
            self._impl.iterate(loops)
            ms = (time.perf_counter() - t0) * 1e3
            rep.elapsed_ms = self._max_over_ranks(ms)
            rep.launches = 0
            rep.plan = {"k": self._impl.k, "exchanges": self._impl.exchanges,
                        "transport": getattr(self, "cpu_transport", "local")}
        else:
            self._impl.run(loops, rep, sync=sync)
            if sync:
                rep.elapsed_ms = self._max_over_ranks(rep.elapsed_ms)
        self.report = rep
        return rep

    def _max_over_ranks(self, v: float) -> float:
        if not self.distributed:
            return v
        import torch
        dev = "cpu" if (self.backend == "cpu" or self.dist.get_backend() == "gloo") else f"cuda:{self.device}"
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    # -- data out ------------------------------------------------------------------------------------
    def local_results(self) -> dict[int, np.ndarray]:
        if self.backend == "cpu":
            return {self._impl.tile.rank: self._impl.result()}
        return self._impl.download()

    def gather(self) -> np.ndarray | None:
        """Full result image on rank 0 (None elsewhere)."""
        g, C = self.grid, self.grid.channels
        local = self.local_results()
        if not self.distributed:
            out = np.empty((g.height, g.row_bytes), dtype=np.uint8)
            for r, tile in local.items():
                self._slice(out, g.tile(r))[:] = tile
            return self._shape(out)
        table = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(local[self.rank], table, dst=0)
        if self.rank != 0:
            return None
        out = np.empty((g.height, g.row_bytes), dtype=np.uint8)
        for r, tile in enumerate(table):
            self._slice(out, g.tile(r))[:] = tile
        return self._shape(out)

    def _shape(self, out2d: np.ndarray) -> np.ndarray:
        g = self.grid
        return out2d.reshape(g.height, g.width) if g.channels == 1 else out2d.reshape(g.height, g.width, g.channels)

    def save_file(self, path: str) -> None:
        """Rank-parallel write (rank 0 creates/truncates, then every rank writes its tile)."""
        g, C = self.grid, self.grid.channels
        if self.rank == 0:
            H.create_file(path, g.height * g.row_bytes)
        if self.distributed:
            self.dist.barrier()
        if self.backend != "cpu" and hasattr(self._impl, "mems"):
            self._save_file_pipelined(path)
        else:
            for r, tile in self.local_results().items():
                t = g.tile(r)
                H.write_rect(path, g.row_bytes, t.row0, t.col0 * C, np.ascontiguousarray(tile))
        if self.distributed:
            self.dist.barrier()

    # -- streaming steps (GPU backends): asynchronous, no host barrier between steps -------------------
    def make_input_pool(self, count: int, seed: int = 1) -> int:
        """Device-resident pool of ``count`` synthetic input images (per local tile)."""
        from .ops import cuda as K
        rt = self._impl
        self._pool = []
        for i in range(count):
            bufs = {}
            for r in rt.local_ranks:
                t, C = self.grid.tile(r), self.grid.channels
                p = K.malloc(rt.mems[r].nbytes)
                K.synth_rect(seed + i, self.grid.row_bytes, t.row0, t.rows, t.col0 * C, t.cols * C,
                             p, rt.pitch, rt.stream)
                bufs[r] = p
            self._pool.append(bufs)
        K.stream_sync(rt.stream)
        return count

    def enqueue_step_from_pool(self, index: int, loops: int) -> int:
        """Enqueue one step: pool image -> tile buffers (device copy) -> ``loops`` iterations.
        Returns the number of kernels launched. Ordering against neighbours is in-stream."""
        from .ops import cuda as K
        rt = self._impl
        n = rt.begin_input()
        for r in rt.local_ranks:
            K.memcpy_async(rt.mems[r].bufs[0], self._pool[index % len(self._pool)][r],
                           rt.mems[r].nbytes - 256, K.D2D, rt.stream)
        rt.parity, rt.slot = 0, 0
        n += rt.end_input()
        _, launches = rt.enqueue(loops)
        return n + launches

    def enqueue_step_pipelined(self, host_in: dict, host_out: dict, loops: int) -> int:
        """Like ``enqueue_step_pinned`` but on rotating buffer slots and three streams, so the host
        copies of neighbouring steps overlap the iterations (needs ``slots >= 2``)."""
        return self._impl.enqueue_step_pipelined(host_in, host_out, loops)

    def pinned_buffers(self) -> tuple[dict, dict]:
        """Pinned host staging (input, output) per local tile, dense rows x tile_bytes."""
        from .ops import cuda as K
        import ctypes
        ins, outs = {}, {}
        for r in self._impl.local_ranks:
            t, C = self.grid.tile(r), self.grid.channels
            n = t.rows * t.cols * C
            for table in (ins, outs):
                p = K.host_alloc(n)
                self.__dict__.setdefault("_pinned", []).append(p)
                arr = np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(p)).reshape(t.rows, t.cols * C)
                table[r] = arr
        return ins, outs

    def enqueue_step_pinned(self, host_in: dict, host_out: dict, loops: int) -> int:
        """Enqueue one end-to-end step: H2D of the input tiles (pinned), ``loops`` iterations, D2H of
        the result tiles (pinned). Asynchronous; ``synchronize()`` completes it."""
        from .ops import cuda as K
        rt = self._impl
        n = rt.begin_input()
        for r in rt.local_ranks:
            a = host_in[r]
            K.memcpy2d_async(rt.mems[r].bufs[0], rt.pitch, a.ctypes.data, a.shape[1], a.shape[1],
                             a.shape[0], K.H2D, rt.stream)
        rt.parity, rt.slot = 0, 0
        n += rt.end_input()
        _, launches = rt.enqueue(loops)
        for r in rt.local_ranks:
            o = host_out[r]
            K.memcpy2d_async(o.ctypes.data, o.shape[1], rt.mems[r].bufs[2 * rt.slot + rt.parity], rt.pitch,
                             o.shape[1], o.shape[0], K.D2H, rt.stream)
        return n + launches

    def slot_sync(self, slot: int) -> None:
        """Wait for the pipelined step that last used buffer slot ``slot`` (its result is on the host)."""
        if hasattr(self._impl, "slot_sync"):
            self._impl.slot_sync(slot)
        else:
            self.synchronize()

    def synchronize(self) -> None:
        from .ops import cuda as K
        if hasattr(self._impl, "pipeline_sync"):
            self._impl.pipeline_sync()
        else:
            K.stream_sync(self._impl.stream)

    @property
    def stream(self) -> int:
        return self._impl.stream

    # -- one-call conveniences -----------------------------------------------------------------------
    def run(self, image: np.ndarray, loops: int) -> np.ndarray | None:
        self.load_image(image)
        self.iterate(loops)
        return self.gather()

    def run_tensor(self, x, loops: int):
        """PyTorch interface of the GPU backends: ``x`` is a uint8 CUDA tensor on this engine's device —
        the whole image ``(H, W[, C])`` in a single process, this rank's tile ``(rows, cols[, C])``
        under ``torch.distributed``. Returns a new tensor of the same shape holding the result of
        ``loops`` iterations. Device to device: the data never visits the host."""
        import torch
        from .ops import cuda as K
        if self.backend not in ("fused", "nccl"):
            raise RuntimeError("run_tensor needs a GPU backend (fused or nccl)")
        g, C = self.grid, self.grid.channels
        tiles = list(self.local_tiles)
        rows, cols = (tiles[0].rows, tiles[0].cols) if self.distributed else (g.height, g.width)
        want = (rows, cols) if C == 1 else (rows, cols, C)
        if (not isinstance(x, torch.Tensor) or x.dtype != torch.uint8 or not x.is_cuda
                or tuple(x.shape) != want or x.device.index != self.device):
            raise ValueError(f"expected a uint8 CUDA tensor of shape {want} on cuda:{self.device}")
        x = x.contiguous()
        out = torch.empty_like(x)
        rt = self._impl
        src_pitch = cols * C
        torch.cuda.current_stream(x.device).synchronize()     # x is complete; `out` is allocated
        if hasattr(rt, "_barrier"):
            rt._barrier()                                     # previous run's readers are done with my buffers
        for t in tiles:
            off = 0 if self.distributed else (t.row0 * g.width + t.col0) * C
            K.memcpy2d_async(rt.mems[t.rank].bufs[0], rt.pitch, x.data_ptr() + off, src_pitch,
                             t.cols * C, t.rows, K.D2D, rt.stream)
        rt.parity = 0
        if hasattr(rt, "slot"):
            rt.slot = 0
        self.iterate(loops)                                   # synchronises the runtime stream
        for t in tiles:
            off = 0 if self.distributed else (t.row0 * g.width + t.col0) * C
            K.memcpy2d_async(out.data_ptr() + off, src_pitch, rt.mems[t.rank].bufs[rt.parity], rt.pitch,
                             t.cols * C, t.rows, K.D2D, rt.stream)
        K.stream_sync(rt.stream)
        return out

    def run_batch(self, paths_in: list[str], loops: int, out_dir: str | None = None) -> dict:
        """Data-parallel-in-time batch mode: filter many images of the engine's shape.

        On the fused backend the images stream through the rotating buffer slots — the file read
        and H2D copy of image i+1 and the D2H copy and file write of image i-1 overlap the
        iterations of image i (``slots >= 2``). Other backends process the list sequentially.
        Outputs are ``blur_<basename>`` in ``out_dir`` (default: next to each input). The
        reference handles exactly one image per program run (SURVEY.md §2.6 "DP: No").
        """
        import time as _time
        g, C = self.grid, self.grid.channels
        t0 = _time.perf_counter()

        def out_path(p: str) -> str:
            d, b = os.path.split(p)
            return os.path.join(out_dir if out_dir is not None else d, "blur_" + b)

        pipelined = self.backend == "fused" and getattr(self._impl, "slots", 1) >= 2 and loops > 0
        if not pipelined:
            for p in paths_in:
                self.run_file(p, loops, out_path(p))
            return {"images": len(paths_in), "seconds": _time.perf_counter() - t0, "pipelined": False}
        slots = self._impl.slots
        stage = [self.pinned_buffers() for _ in range(slots)]
        pending: list[tuple[int, str]] = []      # (slot, output path) in submission order

        def flush(slot: int, path: str) -> None:
            if self.rank == 0:
                H.create_file(path, g.height * g.row_bytes)
            if self.distributed:
                self.dist.barrier()
            for r, tile in stage[slot][1].items():
                t = g.tile(r)
                H.write_rect(path, g.row_bytes, t.row0, t.col0 * C, tile)

        for i, p in enumerate(paths_in):
            slot = i % slots
            if len(pending) >= slots:            # the slot's previous result must leave before reuse
                self.synchronize()
                while pending:
                    flush(*pending.pop(0))
            ins, outs = stage[slot]
            for r, arr in ins.items():
                t = g.tile(r)
                H.read_rect(p, g.row_bytes, t.row0, t.rows, t.col0 * C, t.cols * C, out=arr)
            self.enqueue_step_pipelined(ins, outs, loops)
            pending.append((slot, out_path(p)))
        self.synchronize()
        while pending:
            flush(*pending.pop(0))
        return {"images": len(paths_in), "seconds": _time.perf_counter() - t0, "pipelined": True}

    def run_file(self, path_in: str, loops: int, path_out: str | None = None) -> RunReport:
        """``prog image width height loops {rgb|grey}``: read, iterate, write ``blur_<image>``.

        The output name prefixes the *basename* (the reference prefixes the whole path, defect D8).
        """
        if path_out is None:
            d, b = os.path.split(path_in)
            path_out = os.path.join(d, "blur_" + b)
        self.load_file(path_in)
        rep = self.iterate(loops)
        self.save_file(path_out)
        return rep
