"""Build libdgp_b200.so in-tree with nvcc for sm_100a (no torch extension machinery needed:
the library is a plain C-ABI shared object, see include/dgp_b200.h)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "lib"
OBJDIR = ROOT / "build"
LIB = LIBDIR / "libdgp_b200.so"
SOURCES = ["api.cu", "gemm.cu", "assemble.cu", "potrf.cu", "solve.cu", "grad.cu", "peaks.cu", "dist.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--extended-lambda", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
]


ASAN = os.environ.get("DGP_ASAN") == "1"  # debug builds: host code under AddressSanitizer
if ASAN:
    NVCC_FLAGS += ["-g", "-Xcompiler", "-fsanitize=address", "-Xcompiler", "-fno-omit-frame-pointer"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (cand == "nvcc" or Path(cand).exists()):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(out: Path, deps: list[Path]) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + [ROOT.parent / "include" / "dgp_b200.h"]
    nvcc = _nvcc()

    def compile_one(src: str) -> Path:
        s, o = CSRC / src, OBJDIR / (src + ".o")
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(s), "-o", str(o)]
            if verbose:
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return o

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", str(LIB), *map(str, objs), "-ldl"]
        if ASAN:
            cmd += ["-Xcompiler", "-fsanitize=address"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    build_cpp_mirror_test(force)
    return LIB


def build_cpp_mirror_test(force: bool = False) -> Path:
    """The compiled-language side of the C ABI: tests/cpp/host_mirror_test.cpp (the C++ host mirror,
    include/dynaml_b200.hpp) and tests/cpp/multi_gpu_batch_test.cpp (candidates over several GPUs from one process)."""
    root = ROOT.parent
    out = LIBDIR / "host_mirror_test"
    for name in ("host_mirror_test", "multi_gpu_batch_test"):
        src = root / "tests" / "cpp" / f"{name}.cpp"
        exe = LIBDIR / name
        deps = [src, root / "include" / "dynaml_b200.hpp", root / "include" / "dgp_b200.h", LIB]
        if src.exists() and (force or _stale(exe, deps)):
            cmd = ["g++", "-std=c++17", "-O2", "-pthread", "-I", str(root / "include"), str(src), "-o", str(exe),
                   "-L", str(LIBDIR), "-ldgp_b200", "-Wl,-rpath,$ORIGIN"]
            if ASAN:
                cmd += ["-g", "-fsanitize=address", "-fno-omit-frame-pointer"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"g++ failed for {src}:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
