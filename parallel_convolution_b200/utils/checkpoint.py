"""Checkpoint / resume of an iteration run.

The state of this workload is exactly one image, so a checkpoint is the raw state image
(``<output>.ckpt.raw``, written rank-parallel like any output) plus a JSON sidecar
(``<output>.ckpt.json``: iterations done, shape, filter taps). ``--resume`` reloads the image and
continues with the remaining iterations; results are bit-identical to an uninterrupted run because
every iteration is a pure function of the previous image. The reference has no such facility
(its users re-feed ``blur_x.raw`` by hand).
"""
from __future__ import annotations

import json
import os


def image_path(out_path: str) -> str:
    return out_path + ".ckpt.raw"


def sidecar_path(out_path: str) -> str:
    return out_path + ".ckpt.json"


def original_input(image: str) -> str:
    return image


def save_state(engine, out_path: str, iterations_done: int, args) -> None:
    engine.save_file(image_path(out_path))
    if engine.rank == 0:
        meta = {"iterations_done": int(iterations_done), "width": engine.grid.width,
                "height": engine.grid.height, "channels": engine.grid.channels,
                "filter": engine.filter.describe(), "target_loops": int(args.loops)}
        tmp = sidecar_path(out_path) + ".tmp"
        with open(tmp, "w") as f:
            json.dump(meta, f)
        os.replace(tmp, sidecar_path(out_path))   # the sidecar appears only after the image is complete
    if engine.distributed:
        engine.dist.barrier()


def load_state(out_path: str) -> dict | None:
    try:
        with open(sidecar_path(out_path)) as f:
            state = json.load(f)
    except FileNotFoundError:
        return None
    if not os.path.exists(image_path(out_path)):
        return None
    return state


def check_compatible(state: dict, width: int, height: int, channels: int, filt, loops: int) -> None:
    if (state["width"], state["height"], state["channels"]) != (width, height, channels):
        raise ValueError("checkpoint shape does not match the command line")
    if state["filter"]["taps"] != filt.describe()["taps"]:
        raise ValueError("checkpoint was written with a different filter")
    if state["iterations_done"] > loops:
        raise ValueError("checkpoint is already past the requested iteration count")


def clear_state(out_path: str) -> None:
    for p in (image_path(out_path), sidecar_path(out_path)):
        try:
            os.remove(p)
        except FileNotFoundError:
            pass
