"""GPU clock / throttle-reason sampling during a timed region (B200_PROFILING.md timing rules).

A background thread polls NVML (``pynvml``; ``nvidia-smi`` as a fallback) every few milliseconds
and reports the median SM clock under load, the maximum SM clock and every throttle reason that
was active at least once. ``bench.py`` attaches the summary to its JSON line.
"""
from __future__ import annotations

import statistics
import subprocess
import threading

_REASONS = {
    0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
    0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
    0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
}


class ClockSampler:
    def __init__(self, gpu_indices: list[int], interval_s: float = 0.01):
        self.gpus, self.interval = list(gpu_indices), interval_s
        self.samples: list[int] = []
        self.reason_bits = 0
        self.max_mhz = 0
        self._stop = threading.Event()
        self._thr: threading.Thread | None = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in self.gpus]
            for h in self._handles:
                self.max_mhz = max(self.max_mhz, pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _poll_nvml(self) -> None:
        n = self._nvml
        while not self._stop.is_set():
            for h in self._handles:
                try:
                    self.samples.append(int(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                    self.reason_bits |= int(n.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    try:
                        self.reason_bits |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    except Exception:
                        pass
            self._stop.wait(self.interval)

    def _poll_smi(self) -> None:
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", ",".join(map(str, self.gpus))],
                                     capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    a, b, c = [x.strip() for x in line.split(",")]
                    self.samples.append(int(a))
                    self.max_mhz = max(self.max_mhz, int(b))
                    self.reason_bits |= int(c, 16)
            except Exception:
                pass
            self._stop.wait(max(self.interval, 0.2))

    def start(self) -> "ClockSampler":
        self._thr = threading.Thread(target=self._poll_nvml if self._nvml else self._poll_smi, daemon=True)
        self._thr.start()
        return self

    def stop(self) -> dict:
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        top = max(self.samples) if self.samples else 0
        loaded = [s for s in self.samples if s >= 0.5 * top] or self.samples
        reasons = [name for bit, name in _REASONS.items() if self.reason_bits & bit and name != "gpu_idle"]
        return {"sm_mhz": int(statistics.median(loaded)) if loaded else 0, "sm_max_mhz": int(self.max_mhz),
                "reasons": reasons, "samples": len(self.samples)}
