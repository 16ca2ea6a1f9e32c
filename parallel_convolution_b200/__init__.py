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
import sys as _sys

# Persistent kernels of several *virtual ranks* (N tiles on one GPU: tests, ConvEngine(workers=N)) need one
# hardware queue each; CUDA reads CUDA_DEVICE_MAX_CONNECTIONS when the context is created. Ask for 32 only
# if nobody chose a value and no CUDA context can exist yet; otherwise leave the environment alone —
# ops.cuda.usable_connections() reads what is really in effect and the runtime falls back to the streaming
# kernel when there are more virtual ranks than queues (never a silent deadlock).
_torch = _sys.modules.get("torch")
if "CUDA_DEVICE_MAX_CONNECTIONS" not in _os.environ and not (
        _torch is not None and getattr(_torch, "cuda", None) is not None and _torch.cuda.is_initialized()):
    _os.environ["CUDA_DEVICE_MAX_CONNECTIONS"] = "32"

from .models.filters import Filter, available as available_filters  # noqa: E402,F401

__version__ = "0.1.0"


def __getattr__(name):  # lazy: the engine pulls in torch
    if name in ("ConvEngine", "EngineConfig"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
