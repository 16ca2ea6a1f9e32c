"""ctypes binding of ``libpconv_cuda.so`` (csrc/cuda/pconv_cuda.h).

The kernels are hand-written sm_100a CUDA; this module only mirrors the launch structs and
marshals pointers. There is no PyTorch fallback for the compute: if the library is missing or a
launch fails on a GPU box, these functions raise.
"""
from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache

from .. import build as _build
from .host import FilterSpec

DIRS = 9
SELF = 4
# Persistent kernels of several virtual ranks must run CONCURRENTLY (they spin on each other's flags).
# Streams are multiplexed onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); two rank
# streams sharing a queue serialise their kernels behind each other's event records — a deadlock that
# only the watchdog ends.
MAX_CONNECTIONS = 32


def usable_connections() -> int:
    """Hardware queues concurrent streams can count on. CUDA_DEVICE_MAX_CONNECTIONS is read when the CUDA
    context is created (default 8): the package never sets it behind the user's back — export it (up to
    32) before the first CUDA call to run more virtual ranks with the persistent kernel; with fewer
    usable queues than virtual ranks the runtime falls back to the streaming kernel."""
    try:
        v = int(os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "8"))
    except ValueError:
        v = 8
    return max(1, min(v, MAX_CONNECTIONS))


class Geometry(C.Structure):
    _fields_ = [("img_rows", C.c_int64), ("img_cb", C.c_int64), ("row_lo", C.c_int64 * 4),
                ("cb_lo", C.c_int64 * 4), ("pitch", C.c_int64), ("C", C.c_int32),
                ("pad_", C.c_int32)]


class Sync(C.Structure):
    _fields_ = [("local_flags", C.c_void_p), ("remote_flags", C.c_void_p * 9),
                ("done_counter", C.c_void_p), ("wait_value", C.c_uint32),
                ("signal_value", C.c_uint32), ("timeout_ns", C.c_uint64),
                ("status", C.c_void_p), ("inject_delay_ns", C.c_int32),
                ("inject_drop_signal", C.c_int32)]


class StreamParams(C.Structure):
    _fields_ = [("g", Geometry), ("src", C.c_void_p * 9), ("src_pitch", C.c_int64 * 9),
                ("dst", C.c_void_p), ("filt", FilterSpec), ("k", C.c_int32), ("TH", C.c_int32),
                ("TWB", C.c_int32), ("HB", C.c_int32), ("grid", C.c_int32), ("R", C.c_int32),
                ("sync", Sync)]


class PersistParams(C.Structure):
    _fields_ = [("g", Geometry), ("src", C.c_void_p * 9), ("dst", C.c_void_p), ("filt", FilterSpec),
                ("loops", C.c_int32), ("k", C.c_int32), ("TH", C.c_int32), ("TWB", C.c_int32),
                ("HB", C.c_int32), ("R", C.c_int32), ("W", C.c_int32), ("nty", C.c_int32),
                ("ntx", C.c_int32), ("arena", C.c_void_p * 9), ("arena_bytes", C.c_int64),
                ("epoch_base", C.c_uint32), ("peer_scope", C.c_uint32), ("debug_skip", C.c_uint32),
                ("pad_", C.c_uint32), ("sync", Sync)]


class Rect(C.Structure):
    _fields_ = [("tile", C.c_void_p), ("pitch", C.c_int64), ("row0", C.c_int64), ("byte0", C.c_int64),
                ("rows", C.c_int64), ("nbytes", C.c_int64), ("dense", C.c_void_p)]


class CudaError(RuntimeError):
    pass


@lru_cache(maxsize=1)
def lib() -> C.CDLL:
    path = _build.ensure_cuda()
    L = C.CDLL(os.fspath(path))
    vp, i64 = C.c_void_p, C.c_int64
    L.pconv_stream_launch.argtypes = [C.POINTER(StreamParams), vp]
    L.pconv_stream_run.argtypes = [C.POINTER(StreamParams), C.c_int, C.POINTER(C.c_int32), C.c_int,
                                   C.c_uint32, C.c_int, C.c_uint32, C.POINTER(C.c_void_p), vp]
    L.pconv_sync_launch.argtypes = [C.POINTER(Sync), vp]
    L.pconv_kernel_prepare.argtypes = [C.c_int, C.c_int, C.c_int]
    L.pconv_stream_band_heights.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int]
    L.pconv_persist_launch.argtypes = [C.POINTER(PersistParams), vp]
    L.pconv_persist_arena_bytes.argtypes = [C.POINTER(PersistParams)]
    L.pconv_persist_arena_bytes.restype = i64
    L.pconv_persist_capacity.argtypes = [C.POINTER(PersistParams)]
    L.pconv_persist_band_heights.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int]
    L.pconv_cuda_set_device.argtypes = [C.c_int]
    L.pconv_cuda_l2_bytes.restype = i64
    L.pconv_cuda_last_error.restype = C.c_char_p
    L.pconv_cuda_version.restype = C.c_char_p
    L.pconv_cuda_malloc.argtypes = [C.POINTER(vp), i64]
    L.pconv_cuda_free.argtypes = [vp]
    L.pconv_cuda_ipc_export.argtypes = [vp, C.c_char_p]
    L.pconv_cuda_ipc_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.pconv_cuda_ipc_close.argtypes = [vp]
    L.pconv_cuda_enable_peer.argtypes = [C.c_int]
    L.pconv_cuda_memset_async.argtypes = [vp, C.c_int, i64, vp]
    L.pconv_cuda_memcpy2d_async.argtypes = [vp, i64, vp, i64, i64, i64, C.c_int, vp]
    L.pconv_cuda_stream_sync.argtypes = [vp]
    L.pconv_cuda_memcpy_async.argtypes = [vp, vp, i64, C.c_int, vp]
    L.pconv_cuda_stream_create.argtypes = [C.POINTER(vp)]
    L.pconv_cuda_stream_destroy.argtypes = [vp]
    L.pconv_cuda_event_create.argtypes = [C.POINTER(vp)]
    L.pconv_cuda_event_destroy.argtypes = [vp]
    L.pconv_cuda_event_record.argtypes = [vp, vp]
    L.pconv_cuda_event_sync.argtypes = [vp]
    L.pconv_cuda_stream_wait_event.argtypes = [vp, vp]
    L.pconv_cuda_event_elapsed_ms.argtypes = [vp, vp, C.POINTER(C.c_float)]
    L.pconv_cuda_host_alloc.argtypes = [C.POINTER(vp), i64]
    L.pconv_cuda_host_free.argtypes = [vp]
    L.pconv_cuda_synth_rect.argtypes = [C.c_uint64, i64, i64, i64, i64, i64, vp, i64, vp]
    L.pconv_cuda_flush_l2.argtypes = [vp, i64, vp]
    L.pconv_cuda_pack_rect.argtypes = [vp, i64, i64, i64, i64, i64, vp, vp]
    L.pconv_cuda_unpack_rect.argtypes = [vp, i64, i64, i64, i64, i64, vp, vp]
    L.pconv_cuda_rect_batch.argtypes = [C.POINTER(Rect), C.c_int, C.c_int, vp]
    L.pconv_cuda_launch_count.restype = i64
    return L


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().pconv_cuda_last_error().decode(errors="replace")
        raise CudaError(f"{what} failed ({rc}): {msg}")


def available() -> bool:
    """True when the CUDA library loads and sees at least one device."""
    try:
        return lib().pconv_cuda_device_count() > 0
    except Exception:
        return False


def device_count() -> int:
    return int(lib().pconv_cuda_device_count())


def set_device(dev: int) -> None:
    _check(lib().pconv_cuda_set_device(dev), "set_device")


def sm_count() -> int:
    return int(lib().pconv_cuda_sm_count())


def l2_bytes() -> int:
    return int(lib().pconv_cuda_l2_bytes())


def launch_count() -> int:
    return int(lib().pconv_cuda_launch_count())


def band_heights(kind: int) -> list[int]:
    """Band heights R (rows per warp; region = 16*R rows) compiled for a filter kind."""
    buf = (C.c_int * 16)()
    n = int(lib().pconv_stream_band_heights(kind, buf, 16))
    return [int(buf[i]) for i in range(n)]


def persist_band_heights(kind: int) -> list[int]:
    """Band heights compiled for the persistent kernel (the same list for 8 and 16 warps per CTA)."""
    buf = (C.c_int * 16)()
    n = int(lib().pconv_persist_band_heights(kind, buf, 16))
    return [int(buf[i]) for i in range(n)]


def malloc(nbytes: int) -> int:
    p = C.c_void_p()
    _check(lib().pconv_cuda_malloc(C.byref(p), nbytes), "cudaMalloc")
    return int(p.value)


def free(ptr: int) -> None:
    _check(lib().pconv_cuda_free(ptr), "cudaFree")


def ipc_export(ptr: int) -> bytes:
    buf = C.create_string_buffer(64)
    _check(lib().pconv_cuda_ipc_export(ptr, buf), "ipc_export")
    return buf.raw


def ipc_open(handle: bytes) -> int:
    p = C.c_void_p()
    _check(lib().pconv_cuda_ipc_open(handle, C.byref(p)), "ipc_open")
    return int(p.value)


def ipc_close(ptr: int) -> None:
    _check(lib().pconv_cuda_ipc_close(ptr), "ipc_close")


def enable_peer(dev: int) -> None:
    _check(lib().pconv_cuda_enable_peer(dev), "enable_peer")


def memset_async(ptr: int, value: int, nbytes: int, stream: int) -> None:
    _check(lib().pconv_cuda_memset_async(ptr, value, nbytes, stream), "memset")


H2D, D2H, D2D = 1, 2, 3


def memcpy2d_async(dst: int, dpitch: int, src: int, spitch: int, width: int, rows: int,
                   kind: int, stream: int) -> None:
    _check(lib().pconv_cuda_memcpy2d_async(dst, dpitch, src, spitch, width, rows, kind, stream),
           "memcpy2d")


def stream_sync(stream: int) -> None:
    _check(lib().pconv_cuda_stream_sync(stream), "stream_sync")


def memcpy_async(dst: int, src: int, nbytes: int, kind: int, stream: int) -> None:
    _check(lib().pconv_cuda_memcpy_async(dst, src, nbytes, kind, stream), "memcpy")


def device_sync() -> None:
    _check(lib().pconv_cuda_device_sync(), "device_sync")


def stream_create() -> int:
    p = C.c_void_p()
    _check(lib().pconv_cuda_stream_create(C.byref(p)), "stream_create")
    return int(p.value)


def stream_destroy(s: int) -> None:
    _check(lib().pconv_cuda_stream_destroy(s), "stream_destroy")


class Event:
    """CUDA event on the library's own runtime (timing is always device-side)."""

    def __init__(self) -> None:
        p = C.c_void_p()
        _check(lib().pconv_cuda_event_create(C.byref(p)), "event_create")
        self.handle = int(p.value)

    def record(self, stream: int) -> "Event":
        _check(lib().pconv_cuda_event_record(self.handle, stream), "event_record")
        return self

    def synchronize(self) -> None:
        _check(lib().pconv_cuda_event_sync(self.handle), "event_sync")

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_float()
        _check(lib().pconv_cuda_event_elapsed_ms(self.handle, stop.handle, C.byref(ms)), "elapsed")
        return float(ms.value)

    def __del__(self) -> None:
        try:
            lib().pconv_cuda_event_destroy(self.handle)
        except Exception:
            pass


def stream_wait_event(stream: int, ev: Event) -> None:
    _check(lib().pconv_cuda_stream_wait_event(stream, ev.handle), "stream_wait_event")


def host_alloc(nbytes: int) -> int:
    p = C.c_void_p()
    _check(lib().pconv_cuda_host_alloc(C.byref(p), nbytes), "cudaHostAlloc")
    return int(p.value)


def host_free(ptr: int) -> None:
    _check(lib().pconv_cuda_host_free(ptr), "cudaFreeHost")


def synth_rect(seed: int, row_bytes: int, row0: int, nrows: int, byte0: int, nbytes: int,
               dst: int, pitch: int, stream: int) -> None:
    _check(lib().pconv_cuda_synth_rect(seed & (2**64 - 1), row_bytes, row0, nrows, byte0, nbytes,
                                       dst, pitch, stream), "synth_rect")


def flush_l2(scratch: int, nbytes: int, stream: int) -> None:
    _check(lib().pconv_cuda_flush_l2(scratch, nbytes, stream), "flush_l2")


def pack_rect(tile: int, pitch: int, row0: int, byte0: int, rows: int, nbytes: int, dense: int,
              stream: int) -> None:
    _check(lib().pconv_cuda_pack_rect(tile, pitch, row0, byte0, rows, nbytes, dense, stream),
           "pack_rect")


def unpack_rect(tile: int, pitch: int, row0: int, byte0: int, rows: int, nbytes: int, dense: int,
                stream: int) -> None:
    _check(lib().pconv_cuda_unpack_rect(tile, pitch, row0, byte0, rows, nbytes, dense, stream),
           "unpack_rect")


def rect_batch(rects, n: int, pack: bool, stream: int) -> None:
    """``rects``: ctypes array of ``Rect``; one launch copies all of them (pack: tile -> dense)."""
    _check(lib().pconv_cuda_rect_batch(rects, n, 1 if pack else 0, stream), "rect_batch")


def kernel_prepare(kind: int, channels: int, band: int) -> None:
    """Force the module load of a stream-kernel instance (the persistent kernel: ``persist_capacity``)."""
    _check(lib().pconv_kernel_prepare(kind, channels, band), "kernel_prepare")


def stream_launch(p: StreamParams, stream: int) -> None:
    _check(lib().pconv_stream_launch(C.byref(p), stream), "stream_launch")


def stream_run(blocks, nranks: int, ks: list[int], epoch0: int, use_flags: bool, stream: int,
               first_wait: int = 0, first_flags: list[int] | None = None) -> None:
    """``blocks``: ctypes array of 2*nranks StreamParams (see pconv_stream_run)."""
    arr = (C.c_int32 * len(ks))(*ks)
    ff = (C.c_void_p * nranks)(*first_flags) if first_flags else None
    _check(lib().pconv_stream_run(blocks, nranks, arr, len(ks), epoch0, 1 if use_flags else 0,
                                  first_wait, ff, stream), "stream_run")


def sync_launch(s: Sync, stream: int) -> None:
    _check(lib().pconv_sync_launch(C.byref(s), stream), "sync_launch")


def persist_launch(p: PersistParams, stream: int) -> None:
    _check(lib().pconv_persist_launch(C.byref(p), stream), "persist_launch")


def persist_arena_bytes(p: PersistParams) -> int:
    """Exchange-arena bytes the plan in ``p`` needs."""
    n = int(lib().pconv_persist_arena_bytes(C.byref(p)))
    if n < 0:
        _check(n, "persist_arena_bytes")
    return n


def persist_capacity(p: PersistParams) -> int:
    """CTAs of the plan's kernel instance the device keeps resident (loads the module as a side effect)."""
    n = int(lib().pconv_persist_capacity(C.byref(p)))
    if n < 0:
        _check(n, "persist_capacity")
    return n
