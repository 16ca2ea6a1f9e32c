"""NVTX ranges for the phases of a run (setup / load / iterate / store).

The reference has no tracing at all. Ranges are emitted through
``torch.cuda.nvtx`` when torch + CUDA are present and degrade to no-ops otherwise, so CPU-only
runs and tests never import torch for this.
"""
from __future__ import annotations

import contextlib
import os


@contextlib.contextmanager
def nvtx_range(name: str):
    pushed = False
    if os.environ.get("PCONV_NVTX", "1") != "0":
        try:
            import sys
            if "torch" in sys.modules:     # never pay the torch import just for a trace marker
                import torch
                if torch.cuda.is_available():
                    torch.cuda.nvtx.range_push(name)
                    pushed = True
        except Exception:
            pushed = False
    try:
        yield
    finally:
        if pushed:
            import torch
            torch.cuda.nvtx.range_pop()
