"""Roofline terms for a run: which bound is the relevant one and how far the run is from it.

Denominators are MEASURED: ``MEASURED_PEAKS.json`` when present (HBM copy bandwidth), the profiling
recipe's 770 GB/s NVLink peer-copy figure, and — for the Gaussian SWAR path, the headline — the issue
peak of the kernel's own instruction mix (``scripts/microbench/issue_peak.cu``: the same ``band_sep121``
code iterated on registers only, 16 warps per SM; 28.8 B/clk/SM for RGB, 29.8 for grey on B200,
``profiles/microbench.txt``), at the SM clock sampled during the run. The integer-IMAD and exact-fp32
paths still use a modelled bound (ops per output byte over the two integer-capable issue pipes).
Every term is labelled accordingly.
"""
from __future__ import annotations

import json
import os

NVLINK_GBS = 770.0            # measured peer copy, per direction per GPU (B200_PROFILING.md)
HBM_GBS_FALLBACK = 6650.0     # the recipe's fallback when MEASURED_PEAKS.json is absent
# region bytes per clock per SM: measured issue peak (sep121) / modelled dual-pipe bound (~32 / ~88 lane-ops per 4 B)
ALU_BYTES_PER_CLK = {"sep121-swar": 28.8, "int-swar": 16.0, "f32-exact": 6.0}
ALU_PEAK_SOURCE = {"sep121-swar": "measured: scripts/microbench/issue_peak.cu (profiles/microbench.txt)",
                   "int-swar": "modelled bound, not a measured peak", "f32-exact": "modelled bound, not a measured peak"}
SM_CLOCK_HZ = 1.9e9


def measured_hbm_gbs() -> tuple[float, str]:
    here = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    try:
        with open(os.path.join(here, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return HBM_GBS_FALLBACK, "fallback"


def terms(width: int, height: int, channels: int, loops: int, gpus: int, k: int, path: str,
          halo_bytes_per_exchange: int, sms: int = 148, sm_clock_hz: float = 0.0) -> dict:
    """Seconds each resource needs at its peak for the whole run (per GPU). ``sm_clock_hz``: the SM
    clock sampled during the run (0: the nominal 1.9 GHz)."""
    hbm, src = measured_hbm_gbs()
    clock = sm_clock_hz if sm_clock_hz > 1e8 else SM_CLOCK_HZ
    img = width * height * channels
    k = max(k, 1)
    return {
        "hbm_s": 2.0 * img / k * loops / gpus / (hbm * 1e9),
        "hbm_peak": f"{hbm:.0f} GB/s ({src})",
        "nvlink_halo_s": halo_bytes_per_exchange * (loops / k) / (NVLINK_GBS * 1e9),
        "int_issue_s": img * loops / gpus / (ALU_BYTES_PER_CLK[path] * sms * clock),
        "int_issue_note": ALU_PEAK_SOURCE[path],
    }


def fraction(seconds: float, t: dict) -> float:
    """Achieved fraction of the slowest (binding) term."""
    bound = max(t["hbm_s"], t["nvlink_halo_s"], t["int_issue_s"])
    return bound / max(seconds, 1e-12)
