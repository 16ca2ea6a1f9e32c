"""Checkpoint / resume of an iteration run.

The state of this workload is exactly one image, so a checkpoint is the raw state image
(``<output>.ckpt.raw``, written rank-parallel like any output) plus a JSON sidecar
(``<output>.ckpt.json``: iterations done, shape, filter taps). ``--resume`` reloads the image and
continues with the remaining iterations; results are bit-identical to an uninterrupted run because
every iteration is a pure function of the previous image. The reference has no such facility
(its users re-feed ``blur_x.raw`` by hand).
"""
from __future__ import annotations

import glob
import json
import os
import zlib


def image_path(out_path: str, state: dict | int | None = None) -> str:
    """Path of the state image: of a loaded checkpoint (``state`` = its sidecar), of the checkpoint
    taken after ``state`` iterations (int), or the pre-versioning name (None)."""
    if isinstance(state, dict):
        name = state.get("image")
        return os.path.join(os.path.dirname(out_path), name) if name else out_path + ".ckpt.raw"
    if isinstance(state, int):
        return f"{out_path}.ckpt.{state}.raw"
    return out_path + ".ckpt.raw"


def sidecar_path(out_path: str) -> str:
    return out_path + ".ckpt.json"


def original_input(image: str) -> str:
    return image


def _crc32(path: str) -> int:
    c = 0
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                return c & 0xFFFFFFFF
            c = zlib.crc32(b, c)


def save_state(engine, out_path: str, iterations_done: int, args) -> None:
    """Crash-safe at every checkpoint, not only the first: the state goes to a NEW file
    (``<output>.ckpt.<iterations>.raw``), the sidecar is switched over to it atomically only after the
    image is complete (size and CRC recorded), and the previous image is deleted last. A crash at any
    point leaves a sidecar that points at a complete image (or no sidecar at all)."""
    new = image_path(out_path, int(iterations_done))
    engine.save_file(new)                        # create, barrier, rank-parallel tile writes, barrier
    if engine.rank == 0:
        meta = {"iterations_done": int(iterations_done), "width": engine.grid.width,
                "height": engine.grid.height, "channels": engine.grid.channels,
                "filter": engine.filter.describe(), "target_loops": int(args.loops),
                "image": os.path.basename(new), "bytes": os.path.getsize(new), "crc32": _crc32(new)}
        tmp = sidecar_path(out_path) + ".tmp"
        with open(tmp, "w") as f:
            json.dump(meta, f)
            f.flush()
            os.fsync(f.fileno())
        os.replace(tmp, sidecar_path(out_path))   # from here on a resume uses the new image
        for old in glob.glob(glob.escape(out_path) + ".ckpt.*raw"):
            if os.path.abspath(old) != os.path.abspath(new):
                try:
                    os.remove(old)
                except FileNotFoundError:
                    pass
    if engine.distributed:
        engine.dist.barrier()


def load_state(out_path: str) -> dict | None:
    """The sidecar of a usable checkpoint, None when there is none. A sidecar whose image is missing,
    truncated or fails its checksum raises: resuming from it would silently produce a wrong image."""
    try:
        with open(sidecar_path(out_path)) as f:
            state = json.load(f)
    except FileNotFoundError:
        return None
    img = image_path(out_path, state)
    if not os.path.exists(img):
        if "image" in state:
            raise ValueError(f"checkpoint image {img} is missing")
        return None
    if "bytes" in state and os.path.getsize(img) != int(state["bytes"]):
        raise ValueError(f"checkpoint image {img} has the wrong size (interrupted write?)")
    if "crc32" in state and _crc32(img) != int(state["crc32"]):
        raise ValueError(f"checkpoint image {img} fails its checksum (interrupted write?)")
    return state


def check_compatible(state: dict, width: int, height: int, channels: int, filt, loops: int) -> None:
    if (state["width"], state["height"], state["channels"]) != (width, height, channels):
        raise ValueError("checkpoint shape does not match the command line")
    if state["filter"]["taps"] != filt.describe()["taps"]:
        raise ValueError("checkpoint was written with a different filter")
    if state["iterations_done"] > loops:
        raise ValueError("checkpoint is already past the requested iteration count")


def clear_state(out_path: str) -> None:
    for p in glob.glob(glob.escape(out_path) + ".ckpt.*raw") + [image_path(out_path), sidecar_path(out_path)]:
        try:
            os.remove(p)
        except FileNotFoundError:
            pass
