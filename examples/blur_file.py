#!/usr/bin/env python
"""`./cuda_conv image width height loops {grey|rgb}` (cuda/main.c:10-55) through the Python API.

    python examples/blur_file.py waterfall_grey_1920_2520.raw 1920 2520 50 grey
"""
import sys

from parallel_convolution_b200 import ConvEngine


def main(argv):
    if len(argv) != 6 or argv[5] not in ("grey", "rgb"):
        sys.exit("usage: blur_file.py image width height loops {grey|rgb}")
    image, width, height, loops = argv[1], int(argv[2]), int(argv[3]), int(argv[4])
    with ConvEngine(width=width, height=height, channels=1 if argv[5] == "grey" else 3, filter="gaussian") as eng:
        report = eng.run_file(image, loops)          # reads the raw file, iterates, writes blur_<image>
    print(f"{report.elapsed_ms / 1e3:.6f}")           # like the reference: the seconds of the iteration loop
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
