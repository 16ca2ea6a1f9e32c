"""The filter zoo — the "model family" of this framework.

The reference ships three integer tables and selects one by commenting source lines
(``mpi/mpi_conv.c:95-107``; the CUDA variant hard-codes the Gaussian,
``cuda/cuda_convolute.cu:12``). Here every filter is a runtime value:

* ``box``, ``gaussian`` (default), ``edge`` — the reference's three tables, same numerators and
  divisors (``edge`` is the reference's misnamed low-pass ``1 4 1 / 4 8 4 / 1 4 1`` over 28).
* ``sharpen``, ``laplacian``, ``emboss``, ``sobel_x``, ``sobel_y``, ``identity`` — filters with
  negative taps, which the reference cannot express (its ``float -> uint8_t`` conversion is
  undefined out of range). Our defined behaviour: truncate toward zero, then saturate to
  ``[0, 255]``.
* ``Filter.custom(...)`` / ``Filter.parse("a,b,c,...[/div]")`` / ``Filter.load(path)`` — any 3x3.

Each filter is classified (natively, ``csrc/host/filters.cpp``) into the arithmetic path that is
provably bit-identical to the fp32 oracle: ``sep121-swar``, ``int-swar`` or ``f32-exact``.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np

from ..ops import host as _host

# name -> (integer numerators row-major, divisor)
BUILTIN: dict[str, tuple[tuple[int, ...], float]] = {
    "box": ((1, 1, 1, 1, 1, 1, 1, 1, 1), 9.0),            # mpi/mpi_conv.c:95,103
    "gaussian": ((1, 2, 1, 2, 4, 2, 1, 2, 1), 16.0),      # mpi/mpi_conv.c:96,104 (default)
    "edge": ((1, 4, 1, 4, 8, 4, 1, 4, 1), 28.0),          # mpi/mpi_conv.c:97,105
    "sharpen": ((0, -1, 0, -1, 5, -1, 0, -1, 0), 1.0),
    "laplacian": ((-1, -1, -1, -1, 8, -1, -1, -1, -1), 1.0),
    "emboss": ((-2, -1, 0, -1, 1, 1, 0, 1, 2), 1.0),
    "sobel_x": ((-1, 0, 1, -2, 0, 2, -1, 0, 1), 1.0),
    "sobel_y": ((-1, -2, -1, 0, 0, 0, 1, 2, 1), 1.0),
    "identity": ((0, 0, 0, 0, 1, 0, 0, 0, 0), 1.0),
    "gaussian_soft": ((1, 2, 1, 2, 20, 2, 1, 2, 1), 32.0),
}
ALIASES = {"blur": "gaussian", "gauss": "gaussian", "edge-detect": "edge", "edge_detection": "edge",
           "box_blur": "box", "gaussian_blur": "gaussian", "edge_detect": "edge"}


@dataclass(frozen=True)
class Filter:
    """A 3x3 filter plus its native classification."""

    name: str
    spec: _host.FilterSpec

    # -- constructors ---------------------------------------------------------------------
    @staticmethod
    def builtin(name: str) -> "Filter":
        key = ALIASES.get(name.lower(), name.lower())
        if key not in BUILTIN:
            raise KeyError(f"unknown filter {name!r}; known: {sorted(BUILTIN)}")
        w, div = BUILTIN[key]
        return Filter(key, _host.filter_from_int(w, div))

    @staticmethod
    def from_int(weights, div: float = 1.0, name: str = "custom") -> "Filter":
        w = np.asarray(weights, dtype=np.int64).reshape(-1)
        if w.size != 9:
            raise ValueError("a 3x3 filter needs 9 integer weights")
        return Filter(name, _host.filter_from_int(w, div))

    @staticmethod
    def custom(taps, name: str = "custom") -> "Filter":
        h = np.asarray(taps, dtype=np.float32).reshape(-1)
        if h.size != 9:
            raise ValueError("a 3x3 filter needs 9 taps")
        return Filter(name, _host.filter_from_float(h))

    @staticmethod
    def parse(text: str) -> "Filter":
        """``"gaussian"`` | ``"1,2,1,2,4,2,1,2,1/16"`` | ``"0.1 0.1 ... (9 floats)"`` | a path."""
        t = text.strip()
        key = ALIASES.get(t.lower(), t.lower())
        if key in BUILTIN:
            return Filter.builtin(key)
        if Path(t).is_file():
            return Filter.load(t)
        body, _, div = t.partition("/")
        vals = [v for v in body.replace(",", " ").split() if v]
        if len(vals) != 9:
            raise ValueError(f"cannot parse filter {text!r}")
        if div or all(_is_int(v) for v in vals):
            return Filter.from_int([int(v) for v in vals], float(div) if div else 1.0, name=t)
        return Filter.custom([float(v) for v in vals], name=t)

    @staticmethod
    def load(path: str) -> "Filter":
        """Text file: 9 numbers (whitespace/comma separated), optional last line ``/ div``."""
        toks = Path(path).read_text().replace(",", " ").split()
        if "/" in toks:
            i = toks.index("/")
            return Filter.from_int([int(v) for v in toks[:i]], float(toks[i + 1]), name=Path(path).stem)
        return Filter.parse(" ".join(toks))

    # -- views ----------------------------------------------------------------------------
    @property
    def taps(self) -> np.ndarray:
        return self.spec.taps()

    @property
    def kind(self) -> int:
        return int(self.spec.kind)

    @property
    def kind_name(self) -> str:
        return _host.KIND_NAMES[self.kind]

    def describe(self) -> dict:
        return {"name": self.name, "path": self.kind_name, "taps": self.taps.tolist(),
                "int_weights": self.spec.weights().tolist() if self.kind != _host.KIND_F32 else None,
                "shift": int(self.spec.shift)}


def _is_int(s: str) -> bool:
    try:
        int(s)
        return True
    except ValueError:
        return False


def available() -> list[str]:
    return sorted(BUILTIN)
