"""ctypes binding of ``libpconv_host.so`` (csrc/host/pconv_host.h).

Everything that runs on the CPU is native C++: the oracle, the vectorised/OpenMP CPU engine
(the framework's analogue of the reference's ``mpi_conv`` / ``mpi_omp_conv`` compute), raw I/O,
the decomposition chooser and the synthetic-image generator. Python only marshals pointers.
"""
from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache

import numpy as np

from .. import build as _build

KIND_F32, KIND_INT, KIND_SEP121 = 0, 1, 2
KIND_NAMES = {KIND_F32: "f32-exact", KIND_INT: "int-swar", KIND_SEP121: "sep121-swar"}


class FilterSpec(C.Structure):
    """Mirror of ``pconv_filter_t`` (csrc/common/pconv_types.h)."""

    _fields_ = [("h", C.c_float * 9), ("w", C.c_int32 * 9), ("shift", C.c_int32),
                ("kind", C.c_int32)]

    def taps(self) -> np.ndarray:
        return np.array(list(self.h), dtype=np.float32).reshape(3, 3)

    def weights(self) -> np.ndarray:
        return np.array(list(self.w), dtype=np.int32).reshape(3, 3)


class PlanRequest(C.Structure):
    _fields_ = [("tile_rows", C.c_int64), ("tile_cb", C.c_int64), ("min_nbr_rows", C.c_int64),
                ("min_nbr_cols", C.c_int64), ("C", C.c_int32), ("kind", C.c_int32),
                ("loops", C.c_int32), ("sms", C.c_int32), ("bands", C.c_int32 * 16),
                ("nbands", C.c_int32), ("force_k", C.c_int32), ("force_R", C.c_int32),
                ("resident", C.c_int32), ("peers", C.c_int32), ("warps", C.c_int32)]


class PlanResult(C.Structure):
    _fields_ = [("k", C.c_int32), ("R", C.c_int32), ("TH", C.c_int32), ("TWB", C.c_int32),
                ("HB", C.c_int32), ("ntx", C.c_int32), ("nty", C.c_int32), ("waves", C.c_int32),
                ("est_us", C.c_double)]


@lru_cache(maxsize=1)
def lib() -> C.CDLL:
    path = os.environ.get("PCONV_HOST_LIB") or _build.ensure_host()   # override: sanitizer builds
    L = C.CDLL(os.fspath(path))
    i64, u8p, fsp = C.c_int64, C.c_void_p, C.POINTER(FilterSpec)
    L.pconv_filter_from_int.argtypes = [C.POINTER(C.c_int32), C.c_double, fsp]
    L.pconv_filter_from_float.argtypes = [C.POINTER(C.c_float), fsp]
    L.pconv_divide_rows.argtypes = [i64, i64, C.c_int]
    L.pconv_choose_grid.argtypes = [i64, i64, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int)]
    L.pconv_split.argtypes = [i64, C.c_int, C.c_int, C.POINTER(i64), C.POINTER(i64)]
    L.pconv_split.restype = None
    L.pconv_oracle_run.argtypes = [u8p, u8p, u8p, i64, i64, C.c_int, fsp, C.c_int]
    L.pconv_xplan_emulate.argtypes = [u8p, u8p, i64, i64, C.c_int, fsp, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_uint32]
    L.pconv_cpu_convolve.argtypes = [u8p, u8p, i64, C.c_int, fsp, i64, i64, i64, i64, C.c_int]
    L.pconv_cpu_run_image.argtypes = [u8p, u8p, i64, i64, C.c_int, fsp, C.c_int, C.c_int]
    L.pconv_cpu_run_image_blocked.argtypes = [u8p, u8p, i64, i64, C.c_int, fsp, C.c_int, C.c_int, C.c_int]
    L.pconv_cpu_run_file.argtypes = [C.c_char_p, C.c_char_p, i64, i64, C.c_int, fsp, C.c_int, C.c_int, C.c_int, i64]
    L.pconv_file_size.argtypes = [C.c_char_p]
    L.pconv_file_size.restype = i64
    L.pconv_read_rect.argtypes = [C.c_char_p, i64, i64, i64, i64, i64, u8p, i64]
    L.pconv_create_file.argtypes = [C.c_char_p, i64]
    L.pconv_write_rect.argtypes = [C.c_char_p, i64, i64, i64, i64, i64, u8p, i64]
    L.pconv_synth_rect.argtypes = [C.c_uint64, i64, i64, i64, i64, i64, u8p, i64]
    L.pconv_synth_rect.restype = None
    L.pconv_host_version.restype = C.c_char_p
    L.pconv_plan.argtypes = [C.POINTER(PlanRequest), C.POINTER(PlanResult)]
    L.pconv_shm_create.argtypes = [C.c_char_p, i64, i64, C.c_int, C.c_int]
    L.pconv_shm_create.restype = C.c_void_p
    L.pconv_shm_attach.argtypes = [C.c_char_p]
    L.pconv_shm_attach.restype = C.c_void_p
    L.pconv_shm_close.argtypes = [C.c_void_p]
    L.pconv_shm_close.restype = None
    L.pconv_shm_pixels.argtypes = [C.c_void_p, C.c_int]
    L.pconv_shm_pixels.restype = C.c_void_p
    for fn in (L.pconv_shm_pitch, L.pconv_shm_bytes):
        fn.argtypes = [C.c_void_p]
        fn.restype = i64
    L.pconv_shm_epoch.argtypes = [C.c_void_p, C.c_int]
    L.pconv_shm_epoch.restype = i64
    for fn in (L.pconv_shm_reset, L.pconv_shm_abort):
        fn.argtypes = [C.c_void_p, C.c_int]
        fn.restype = None
    L.pconv_shm_run.argtypes = [C.c_void_p, C.c_int, i64, i64, i64, i64, C.POINTER(C.c_int32), fsp, C.c_int,
                                C.c_int, C.c_int, C.c_double, C.POINTER(i64)]
    L.pconv_net_init.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_double]
    L.pconv_net_init.restype = C.c_void_p
    L.pconv_net_close.argtypes = [C.c_void_p]
    L.pconv_net_close.restype = None
    L.pconv_net_error.argtypes = [C.c_void_p]
    L.pconv_net_error.restype = C.c_char_p
    L.pconv_net_barrier.argtypes = [C.c_void_p]
    L.pconv_net_max.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.pconv_net_tile_pitch.argtypes = [i64, C.c_int, C.c_int]
    L.pconv_net_tile_pitch.restype = i64
    L.pconv_net_run.argtypes = [C.c_void_p, i64, i64, C.c_int, i64, i64, i64, i64, C.POINTER(C.c_int32), u8p, u8p, fsp,
                                C.c_int, C.c_int, C.c_int, C.POINTER(i64)]
    return L


def _ptr(a: np.ndarray) -> int:
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"], "uint8 C-contiguous array required"
    return a.ctypes.data


def filter_from_int(w, div: float) -> FilterSpec:
    arr = (C.c_int32 * 9)(*[int(v) for v in np.asarray(w).reshape(-1)])
    f = FilterSpec()
    if lib().pconv_filter_from_int(arr, float(div), C.byref(f)) != 0:
        raise ValueError("bad integer filter (zero divisor?)")
    return f


def filter_from_float(h) -> FilterSpec:
    arr = (C.c_float * 9)(*[float(v) for v in np.asarray(h, dtype=np.float32).reshape(-1)])
    f = FilterSpec()
    if lib().pconv_filter_from_float(arr, C.byref(f)) != 0:
        raise ValueError("bad float filter")
    return f


def divide_rows(rows: int, cols: int, workers: int) -> int:
    return int(lib().pconv_divide_rows(rows, cols, workers))


def choose_grid(rows: int, cols: int, workers: int) -> tuple[int, int, bool]:
    r, c, e = C.c_int(), C.c_int(), C.c_int()
    rc = lib().pconv_choose_grid(rows, cols, workers, C.byref(r), C.byref(c), C.byref(e))
    if rc != 0:
        raise ValueError(f"cannot decompose {rows}x{cols} over {workers} workers")
    return r.value, c.value, bool(e.value)


def split(extent: int, parts: int, idx: int) -> tuple[int, int]:
    s, n = C.c_int64(), C.c_int64()
    lib().pconv_split(extent, parts, idx, C.byref(s), C.byref(n))
    return s.value, n.value


def plan(tile_rows: int, tile_cb: int, Cn: int, kind: int, loops: int, bands: list[int], *,
         min_nbr_rows: int = 0, min_nbr_cols: int = 0, sms: int = 148, force_k: int = 0,
         force_R: int = 0, resident: bool = False, peers: bool = False,
         warps: int = 16) -> PlanResult | None:
    """Native launch planner (csrc/host/planner.cpp). None when no plan is feasible."""
    q = PlanRequest()
    q.tile_rows, q.tile_cb, q.min_nbr_rows, q.min_nbr_cols = tile_rows, tile_cb, min_nbr_rows, min_nbr_cols
    q.C, q.kind, q.loops, q.sms = Cn, kind, loops, sms
    for i, b in enumerate(bands[:16]):
        q.bands[i] = b
    q.nbands = min(len(bands), 16)
    q.force_k, q.force_R = force_k or 0, force_R or 0
    q.resident, q.peers, q.warps = int(resident), int(peers), int(warps)
    out = PlanResult()
    rc = lib().pconv_plan(C.byref(q), C.byref(out))
    if rc < 0:
        raise ValueError("bad plan request")
    return out if rc == 0 else None


def oracle(image: np.ndarray, f: FilterSpec, loops: int) -> np.ndarray:
    """Golden reference. ``image``: (H, W) or (H, W, C) uint8."""
    img = np.ascontiguousarray(image, dtype=np.uint8)
    H, W = img.shape[:2]
    Cn = 1 if img.ndim == 2 else img.shape[2]
    out, scratch = np.empty_like(img), np.empty_like(img)
    rc = lib().pconv_oracle_run(_ptr(img), _ptr(out), _ptr(scratch), H, W, Cn, C.byref(f), loops)
    if rc != 0:
        raise RuntimeError(f"oracle failed ({rc})")
    return out


def xplan_emulate(image: np.ndarray, f: FilterSpec, loops: int, grid: tuple[int, int], k: int, R: int,
                  W: int, TH: int = 0, epoch_base: int = 0) -> np.ndarray | int:
    """CPU replay of the persistent kernel's pushed halo exchange (csrc/host/xplan_emu.cpp): the result
    image, or the negative code of a geometry the launcher rejects."""
    img = np.ascontiguousarray(image, dtype=np.uint8)
    H, Wpx = img.shape[:2]
    Cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.zeros_like(img)
    rc = lib().pconv_xplan_emulate(_ptr(img), _ptr(out), H, Wpx, Cn, C.byref(f), loops, grid[0], grid[1],
                                   k, R, W, TH, epoch_base)
    return out if rc == 0 else int(rc)


def cpu_run_image(image: np.ndarray, f: FilterSpec, loops: int, threads: int = 0) -> np.ndarray:
    img = np.ascontiguousarray(image, dtype=np.uint8)
    H, W = img.shape[:2]
    Cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    rc = lib().pconv_cpu_run_image(_ptr(img), _ptr(out), H, W, Cn, C.byref(f), loops, threads)
    if rc != 0:
        raise RuntimeError(f"cpu engine failed ({rc})")
    return out


def cpu_run_image_blocked(image: np.ndarray, f: FilterSpec, loops: int, k: int, threads: int = 0) -> np.ndarray:
    """Temporally blocked CPU run (k iterations per pass over memory); bit-identical to the oracle."""
    img = np.ascontiguousarray(image, dtype=np.uint8)
    H, W = img.shape[:2]
    Cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    rc = lib().pconv_cpu_run_image_blocked(_ptr(img), _ptr(out), H, W, Cn, C.byref(f), loops, threads, k)
    if rc != 0:
        raise RuntimeError(f"blocked cpu engine failed ({rc})")
    return out


def cpu_run_file(path_in: str, path_out: str, height: int, width: int, Cn: int, f: FilterSpec, loops: int,
                 threads: int = 0, k: int = 0, mem_bytes: int = 0) -> None:
    """Out-of-core file -> file run of the native CPU engine: row strips with k-deep halos,
    ceil(loops/k) passes, at most ~``mem_bytes`` of the image in memory. Bit-identical to the oracle."""
    rc = lib().pconv_cpu_run_file(os.fsencode(path_in), os.fsencode(path_out), height, width, Cn,
                                  C.byref(f), loops, threads, k, mem_bytes)
    if rc != 0:
        raise OSError(-rc if rc < 0 else rc, f"out-of-core run failed ({os.strerror(-rc) if rc < 0 else rc}): "
                                             f"{path_in} -> {path_out}")


class ShmSegment:
    """The shared image of the CPU ranks of one node (csrc/host/shm_ranks.cpp): two ping-pong images with
    their zero frame and one epoch counter per rank. Rank 0 creates it, the others attach."""

    def __init__(self, path: str, height: int = 0, width: int = 0, Cn: int = 0, nranks: int = 0,
                 create: bool = False):
        L = lib()
        self.path, self._h = path, None
        h = (L.pconv_shm_create(os.fsencode(path), height, width, Cn, nranks) if create
             else L.pconv_shm_attach(os.fsencode(path)))
        if not h:
            raise OSError(f"cannot {'create' if create else 'attach to'} shared segment {path}")
        self._h = h
        self.pitch = int(L.pconv_shm_pitch(h))
        self.nbytes = int(L.pconv_shm_bytes(h))

    def image(self, which: int, height: int, row_bytes: int) -> np.ndarray:
        """Writable (height, row_bytes) view of image ``which`` (0/1), frame excluded."""
        base = lib().pconv_shm_pixels(self._h, which & 1)
        buf = (C.c_uint8 * ((height - 1) * self.pitch + row_bytes)).from_address(base)
        flat = np.frombuffer(buf, dtype=np.uint8)
        return np.lib.stride_tricks.as_strided(flat, shape=(height, row_bytes), strides=(self.pitch, 1))

    def epoch(self, rank: int) -> int:
        return int(lib().pconv_shm_epoch(self._h, rank))

    def reset(self, rank: int) -> None:
        lib().pconv_shm_reset(self._h, rank)

    def abort(self, rank: int) -> None:
        lib().pconv_shm_abort(self._h, rank)

    def run(self, rank: int, rect: tuple[int, int, int, int], nbr: list[int], f: FilterSpec, loops: int, k: int,
            threads: int = 1, timeout_s: float = 0.0) -> int:
        """Advance ``rank``'s tile ``rect`` = (r0, r1, c0, c1) by ``loops`` iterations; returns the epochs run."""
        arr = (C.c_int32 * 9)(*[int(v) for v in nbr])
        ran = C.c_int64(0)
        rc = lib().pconv_shm_run(self._h, rank, *[int(v) for v in rect], arr, C.byref(f), loops, k,
                                 max(1, int(threads)), float(timeout_s), C.byref(ran))
        if rc != 0:
            raise OSError(-rc, f"shared-memory rank {rank}: {os.strerror(-rc)} "
                               f"({'a neighbour stopped advancing' if rc == -110 else 'a neighbour gave up' if rc == -125 else 'bad arguments'})")
        return int(ran.value)

    def close(self) -> None:
        if self._h:
            lib().pconv_shm_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NetComm:
    """A rank's end of the TCP mesh of csrc/host/net_ranks.cpp (rendezvous through rank 0 at host:port;
    ``master_fd`` hands rank 0 an already listening socket)."""

    def __init__(self, rank: int, world: int, host: str, port: int, master_fd: int = -1, timeout_s: float = 120.0):
        self._h = lib().pconv_net_init(rank, world, os.fsencode(host), port, master_fd, timeout_s)
        if not self._h:
            raise OSError(f"rank {rank}: cannot join the TCP mesh at {host}:{port}")
        self.rank, self.world = rank, world

    def _check(self, rc: int, what: str) -> int:
        if rc < 0:
            raise OSError(-rc, f"{what}: {lib().pconv_net_error(self._h).decode()}")
        return rc

    def barrier(self) -> None:
        self._check(lib().pconv_net_barrier(self._h), "barrier")

    def max(self, v: float) -> float:
        x = C.c_double(v)
        self._check(lib().pconv_net_max(self._h, C.byref(x)), "max over ranks")
        return float(x.value)

    def run(self, shape: tuple[int, int, int], tile: tuple[int, int, int, int], nbr: list[int], bufs: tuple[np.ndarray, np.ndarray],
            f: FilterSpec, loops: int, k: int, threads: int = 1) -> tuple[int, int]:
        """``shape`` = (H, W, C) of the image, ``tile`` = (row0, rows, col0, cols); ``bufs``: two zero-framed
        (rows + 2k, pitch) arrays, input in bufs[0]. Returns (index of the buffer holding the result, exchanges)."""
        arr = (C.c_int32 * 9)(*[int(v) for v in nbr])
        nx = C.c_int64(0)
        rc = lib().pconv_net_run(self._h, shape[0], shape[1], shape[2], tile[0], tile[1], tile[2], tile[3], arr,
                                 _ptr(bufs[0]), _ptr(bufs[1]), C.byref(f), loops, k, max(1, int(threads)), C.byref(nx))
        return self._check(rc, "halo exchange"), int(nx.value)

    def close(self) -> None:
        if self._h:
            lib().pconv_net_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def net_tile_pitch(cols: int, Cn: int, k: int) -> int:
    return int(lib().pconv_net_tile_pitch(cols, Cn, k))


def cpu_convolve(src: np.ndarray, dst: np.ndarray, halo: int, Cn: int, f: FilterSpec,
                 r0: int, r1: int, c0: int, c1: int, threads: int = 0) -> None:
    """One pass over rows [r0,r1) x pixel columns [c0,c1) of a halo-padded tile.

    ``src``/``dst``: 2-D uint8 arrays (rows+2*halo, pitch) whose interior pixel (0,0) sits at
    [halo, halo*Cn]. Coordinates may be negative / exceed the interior as long as the 1-pixel
    stencil footprint stays inside the padded array (used by the k-deep temporal blocking of
    the host backend).
    """
    assert src.shape == dst.shape and src.ndim == 2
    pitch = src.strides[0]
    off = halo * pitch + halo * Cn
    rc = lib().pconv_cpu_convolve(src.ctypes.data + off, dst.ctypes.data + off, pitch, Cn,
                                  C.byref(f), r0, r1, c0, c1, threads)
    if rc != 0:
        raise RuntimeError(f"cpu_convolve failed ({rc})")


def file_size(path: str) -> int:
    n = int(lib().pconv_file_size(os.fsencode(path)))
    if n < 0:
        raise OSError(-n, os.strerror(-n), path)
    return n


def read_rect(path: str, row_bytes: int, row0: int, nrows: int, byte0: int, nbytes: int,
              out: np.ndarray | None = None) -> np.ndarray:
    if out is None:
        out = np.empty((nrows, nbytes), dtype=np.uint8)
    assert out.dtype == np.uint8 and out.strides[1] == 1
    rc = lib().pconv_read_rect(os.fsencode(path), row_bytes, row0, nrows, byte0, nbytes,
                               out.ctypes.data, out.strides[0])
    if rc != 0:
        raise OSError(-rc, os.strerror(-rc), path)
    return out


def create_file(path: str, total_bytes: int) -> None:
    rc = lib().pconv_create_file(os.fsencode(path), total_bytes)
    if rc != 0:
        raise OSError(-rc, os.strerror(-rc), path)


def write_rect(path: str, row_bytes: int, row0: int, byte0: int, data: np.ndarray) -> None:
    assert data.dtype == np.uint8 and data.ndim == 2 and data.strides[1] == 1
    rc = lib().pconv_write_rect(os.fsencode(path), row_bytes, row0, data.shape[0], byte0,
                                data.shape[1], data.ctypes.data, data.strides[0])
    if rc != 0:
        raise OSError(-rc, os.strerror(-rc), path)


def synth_rect(seed: int, row_bytes: int, row0: int, nrows: int, byte0: int,
               nbytes: int) -> np.ndarray:
    out = np.empty((nrows, nbytes), dtype=np.uint8)
    lib().pconv_synth_rect(seed & (2**64 - 1), row_bytes, row0, nrows, byte0, nbytes,
                           out.ctypes.data, out.strides[0])
    return out


def omp_max_threads() -> int:
    return int(lib().pconv_omp_max_threads())
