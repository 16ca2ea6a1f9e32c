"""Raw <-> binary PGM/PPM (netpbm P5/P6) conversion, so results can be opened in any image viewer.

The reference works on headerless raw files only and documents its results as screenshots of a raw
viewer (``README.md:11-14``, ``results/*.png``); a PGM/PPM file is the same bytes behind a three-line
text header. Copies in 64 MiB blocks, so it works for images larger than memory.

    python -m parallel_convolution_b200.utils.netpbm raw2pnm blur_x.raw 1920 2520 rgb blur_x.ppm
    python -m parallel_convolution_b200.utils.netpbm pnm2raw picture.ppm picture.raw   # prints W H rgb|grey
"""
from __future__ import annotations

import os
import sys

_BLOCK = 64 << 20


def _copy(src, dst, nbytes: int) -> None:
    left = nbytes
    while left > 0:
        buf = src.read(min(left, _BLOCK))
        if not buf:
            raise OSError(f"input ends {left} bytes early")
        dst.write(buf)
        left -= len(buf)


def raw_to_pnm(raw_path: str, width: int, height: int, channels: int, out_path: str) -> None:
    if channels not in (1, 3):
        raise ValueError("netpbm holds grey (1) or RGB (3) images")
    need = width * height * channels
    if os.path.getsize(raw_path) < need:
        raise ValueError(f"{raw_path}: {os.path.getsize(raw_path)} bytes, need {need}")
    with open(raw_path, "rb") as src, open(out_path, "wb") as dst:
        dst.write(f"{'P5' if channels == 1 else 'P6'}\n{width} {height}\n255\n".encode("ascii"))
        _copy(src, dst, need)


def read_header(f) -> tuple[int, int, int]:
    """Parse a binary PGM/PPM header from an open file; returns (width, height, channels) with the
    file positioned at the first pixel byte."""
    tokens: list[bytes] = []
    while len(tokens) < 4:
        line = f.readline()
        if not line:
            raise ValueError("truncated netpbm header")
        line = line.split(b"#", 1)[0]
        tokens += line.split()
        if len(tokens) > 4:
            raise ValueError("unsupported netpbm header layout (pixels on the header line)")
    magic, w, h, maxval = tokens
    if magic not in (b"P5", b"P6") or int(maxval) != 255:
        raise ValueError("only binary 8-bit PGM (P5) / PPM (P6) are supported")
    return int(w), int(h), 1 if magic == b"P5" else 3


def pnm_to_raw(pnm_path: str, out_path: str) -> tuple[int, int, int]:
    with open(pnm_path, "rb") as src:
        w, h, c = read_header(src)
        with open(out_path, "wb") as dst:
            _copy(src, dst, w * h * c)
    return w, h, c


def main(argv: list[str] | None = None) -> int:
    a = list(sys.argv[1:] if argv is None else argv)
    try:
        if len(a) == 6 and a[0] == "raw2pnm" and a[4] in ("rgb", "grey", "gray"):
            raw_to_pnm(a[1], int(a[2]), int(a[3]), 3 if a[4] == "rgb" else 1, a[5])
            return 0
        if len(a) == 3 and a[0] == "pnm2raw":
            w, h, c = pnm_to_raw(a[1], a[2])
            print(w, h, "rgb" if c == 3 else "grey")
            return 0
    except (OSError, ValueError) as e:
        sys.stderr.write(f"netpbm: {e}\n")
        return 1
    sys.stderr.write(__doc__ or "")
    return 2


if __name__ == "__main__":
    sys.exit(main())
