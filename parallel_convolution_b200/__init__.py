"""parallel_convolution_b200 — a B200-native iterated 3x3 image-convolution engine.

Capabilities of jimouris/parallel-convolution (MPI / MPI+OpenMP / CUDA programs applying a 3x3
filter ``loops`` times to raw grey / interleaved-RGB images), rebuilt for one 8xB200 NVSwitch box:
hand-written sm_100a kernels with temporal blocking in registers, halo exchange fused into
the stencil kernel through NVLink peer memory, a NCCL baseline, and a native CPU engine.

Public API::

    from parallel_convolution_b200 import ConvEngine, Filter
    eng = ConvEngine(width=1920, height=2520, channels=3, filter="gaussian")
    out = eng.run(image_uint8, loops=100)
"""
import os as _os

# Must be in the environment before ANY library creates the CUDA context (see ops/cuda.py): persistent
# kernels of several virtual ranks need one hardware queue each.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .models.filters import Filter, available as available_filters  # noqa: E402,F401

__version__ = "0.1.0"


def __getattr__(name):  # lazy: the engine pulls in torch
    if name in ("ConvEngine", "EngineConfig"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
