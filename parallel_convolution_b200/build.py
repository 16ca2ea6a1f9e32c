"""In-tree native build: g++ for the host library, nvcc (sm_100a only) for the CUDA library.

The shared objects land in ``parallel_convolution_b200/_lib/`` (git-ignored, but they travel to
the GPU box with the gpurun snapshot). ``python -m parallel_convolution_b200.build`` builds
everything; ``ensure_host()`` / ``ensure_cuda()`` rebuild lazily when a source is newer.

There is exactly one GPU target: ``-gencode arch=compute_100a,code=sm_100a`` (B200). No
multi-arch fatbin, no fallback path.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "_lib"

HOST_LIB = LIB / "libpconv_host.so"
CUDA_LIB = LIB / "libpconv_cuda.so"
NATIVE_BIN = LIB / "pconv_native"

NVCC = os.environ.get("PCONV_NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
def _pick_cxx() -> str:
    # /opt/gcc/bin/g++ (first on PATH in this image) is a wrapper without libgomp.spec; the
    # distro compiler links OpenMP fine. CXX in the environment only wins when it is a path we
    # can actually run.
    for cand in ("/usr/bin/g++", os.environ.get("PCONV_CXX"), shutil.which("g++")):
        if cand and Path(cand).exists():
            return cand
    return "g++"


CXX = _pick_cxx()

HOST_SOURCES = ["host/filters.cpp", "host/decomp.cpp", "host/cpu_engine.cpp", "host/rawio.cpp",
                "host/planner.cpp", "host/xplan_emu.cpp", "host/shm_ranks.cpp", "host/net_ranks.cpp"]
HOST_ORACLE = "host/oracle.cpp"  # compiled separately at -O1 with contraction off
CUDA_SOURCES = ["cuda/runtime.cu", "cuda/pack_kernels.cu", "cuda/stream_kernel.cu", "cuda/persist_kernel.cu",
                "cuda/stream_kind_sep.cu", "cuda/stream_kind_int.cu", "cuda/stream_kind_f32.cu",
                "cuda/persist_sep_w8.cu", "cuda/persist_sep_w16.cu", "cuda/persist_int_w8.cu",
                "cuda/persist_int_w16.cu", "cuda/persist_f32_w8.cu", "cuda/persist_f32_w16.cu"]

CXXFLAGS = ["-std=c++17", "-fPIC", "-O3", "-ffp-contract=off", "-fopenmp", "-Wall",
            "-Wno-unknown-pragmas"]
NVCCFLAGS = ["-ccbin", CXX, "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
             "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--fmad=false"]


class _BuildLock:
    """Exclusive lock around every build: N torchrun ranks importing the package with a stale content hash
    must not compile into the same _lib/ paths at once (a rank could dlopen a half-written library)."""

    def __enter__(self):
        import fcntl
        LIB.mkdir(exist_ok=True)
        self.f = open(LIB / ".build.lock", "w")
        fcntl.flock(self.f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self.f, fcntl.LOCK_UN)
        self.f.close()


def _link(cmd_prefix: list[str], target: Path, inputs: list[str], libs: list[str]) -> None:
    """Link to a temporary name and rename: readers never see a partially written artefact."""
    tmp = target.with_name(target.name + f".tmp{os.getpid()}")
    _run([*cmd_prefix, "-o", str(tmp), *inputs, *libs])
    os.replace(tmp, target)


def _run(cmd: list[str], log: Path | None = None) -> None:
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        log.write_text(" ".join(cmd) + "\n" + proc.stdout)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError(f"build step failed ({proc.returncode}): {' '.join(cmd)}")


def _digest(sources: list[Path], flags: list[str]) -> str:
    """Content hash of a target's inputs. File times are useless here: the tree is snapshotted onto
    fresh GPU boxes and checked out by the driver, which rewrites every mtime."""
    import hashlib
    h = hashlib.sha256(" ".join(flags).encode())
    for s in sorted(set(sources), key=lambda q: str(q)):
        if s.exists():
            h.update(str(s.relative_to(PKG)).encode())
            h.update(s.read_bytes())
    return h.hexdigest()


def _stale(target: Path, sources: list[Path], flags: list[str] | None = None) -> bool:
    stamp = target.with_name(target.name + ".hash")
    if not target.exists() or not stamp.exists():
        return True
    return stamp.read_text().strip() != _digest(sources, flags or [])


def _stamp(target: Path, sources: list[Path], flags: list[str] | None = None) -> None:
    target.with_name(target.name + ".hash").write_text(_digest(sources, flags or []))


def _all_headers() -> list[Path]:
    return list(CSRC.rglob("*.h")) + list(CSRC.rglob("*.cuh"))


def _headers(*dirs: str) -> list[Path]:
    """Headers a target can include: its own directory plus common/ (a host-only edit must not invalidate
    the CUDA objects, which take minutes to rebuild)."""
    out: list[Path] = []
    for d in dirs:
        out += list((CSRC / d).rglob("*.h")) + list((CSRC / d).rglob("*.cuh"))
    return out


def build_host(force: bool = False) -> Path:
    LIB.mkdir(exist_ok=True)
    srcs = [CSRC / s for s in HOST_SOURCES] + [CSRC / HOST_ORACLE]
    if not force and not _stale(HOST_LIB, srcs + _headers("host", "common"), CXXFLAGS):
        return HOST_LIB
    with _BuildLock():
        if not force and not _stale(HOST_LIB, srcs + _headers("host", "common"), CXXFLAGS):
            return HOST_LIB      # another rank built it while we waited for the lock
        return _build_host_locked(srcs)


def _build_host_locked(srcs: list[Path]) -> Path:
    objs = []
    for s in HOST_SOURCES:
        o = LIB / (Path(s).stem + ".o")
        _run([CXX, *CXXFLAGS, "-c", str(CSRC / s), "-o", str(o)])
        objs.append(str(o))
    o = LIB / "oracle.o"
    _run([CXX, "-std=c++17", "-fPIC", "-O1", "-ffp-contract=off", "-fno-tree-vectorize",
          "-c", str(CSRC / HOST_ORACLE), "-o", str(o)])
    objs.append(str(o))
    _link([CXX, "-shared", "-fopenmp"], HOST_LIB, objs, [])
    _stamp(HOST_LIB, srcs + _headers("host", "common"), CXXFLAGS)
    return HOST_LIB


def build_host_sanitized() -> Path:
    """Host library instrumented with AddressSanitizer + UndefinedBehaviorSanitizer (debug aid;
    scripts/host_sanitize.sh runs the CPU tests against it with LD_PRELOAD=libasan)."""
    LIB.mkdir(exist_ok=True)
    out = LIB / "libpconv_host_asan.so"
    srcs = [str(CSRC / s) for s in HOST_SOURCES] + [str(CSRC / HOST_ORACLE)]
    _run([CXX, "-std=c++17", "-fPIC", "-O1", "-g", "-ffp-contract=off", "-fopenmp", "-fno-omit-frame-pointer",
          "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-shared", "-o", str(out), *srcs])
    return out


def build_cuda(force: bool = False) -> Path:
    LIB.mkdir(exist_ok=True)
    srcs = [CSRC / s for s in CUDA_SOURCES if (CSRC / s).exists()]
    if not force and not _stale(CUDA_LIB, srcs + _headers("cuda", "common"), NVCCFLAGS):
        return CUDA_LIB
    with _BuildLock():
        if not force and not _stale(CUDA_LIB, srcs + _headers("cuda", "common"), NVCCFLAGS):
            return CUDA_LIB
        return _build_cuda_locked(srcs, force)


def _build_cuda_locked(srcs: list[Path], force: bool) -> Path:
    if not Path(NVCC).exists():
        raise RuntimeError(f"nvcc not found at {NVCC}; the CUDA library cannot be built")
    from concurrent.futures import ThreadPoolExecutor
    hdrs = _headers("cuda", "common")

    def compile_one(s: Path) -> str:
        o = LIB / (s.stem + ".cu.o")
        if force or _stale(o, [s] + hdrs, NVCCFLAGS):
            _run([NVCC, *NVCCFLAGS, "-c", str(s), "-o", str(o)], log=LIB / (s.stem + ".ptxas.log"))
            _stamp(o, [s] + hdrs, NVCCFLAGS)
        return str(o)

    with ThreadPoolExecutor(max_workers=min(os.cpu_count() or 4, len(srcs))) as ex:   # the kernel TUs take minutes
        objs = list(ex.map(compile_one, srcs))
    _link([NVCC, "-ccbin", CXX, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"], CUDA_LIB, objs,
          ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    _stamp(CUDA_LIB, srcs + _headers("cuda", "common"), NVCCFLAGS)
    return CUDA_LIB


def build_native(force: bool = False) -> Path | None:
    """Stand-alone C++ CLI (drop-in for the reference's ./cuda_conv on one GPU, CPU fallback)."""
    main = CSRC / "native_main.cpp"
    if not main.exists():
        return None
    build_host(force)
    build_cuda(force)
    nsrc = [main] + _all_headers()
    if not force and NATIVE_BIN.exists() and not _stale(NATIVE_BIN, nsrc):
        return NATIVE_BIN
    _run([CXX, "-std=c++17", "-O2", f"-I{CSRC}", str(main), "-o", str(NATIVE_BIN), f"-L{LIB}",
          "-lpconv_host", "-lpconv_cuda", "-Wl,-rpath,$ORIGIN", "-fopenmp"])
    _stamp(NATIVE_BIN, nsrc)
    return NATIVE_BIN


def ensure_host() -> Path:
    return build_host(False)


def ensure_cuda() -> Path:
    return build_cuda(False)


def build_all(force: bool = False) -> None:
    build_host(force)
    build_cuda(force)
    build_native(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", *[p for p in (HOST_LIB, CUDA_LIB, NATIVE_BIN) if p.exists()], sep="\n  ")
