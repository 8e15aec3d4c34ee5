"""Builds ipt_b200/lib/libipt_b200.so (CUDA kernels + C ABI + C++ host classes) for sm_100a with nvcc.

In-tree on purpose: the built library travels to the GPU box with the repository snapshot.
`python -m ipt_b200.build [--force] [--verbose]`
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "ipt_b200"
LIB = PKG / "lib" / "libipt_b200.so"
SOURCES = [PKG / "csrc" / "ipt_capi.cu", PKG / "host" / "sample_scenes.cpp", PKG / "host" / "device_plugins.cpp"]
DEPS = (
    list((PKG / "csrc").glob("*.cu*"))
    + list((PKG / "host").glob("*.[ch]pp"))
    + list((PKG / "host" / "compat").glob("*.h"))
    + [ROOT / "include" / "ipt_b200.h"]
)
NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-ffp-contract=off",
    "-shared",
    "-ldl",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")
    return nvcc


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.exists() and p.stat().st_mtime > t for p in DEPS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    srcs = [str(s) for s in SOURCES if s.exists()]
    cmd = [find_nvcc(), *NVCC_FLAGS, f"-I{ROOT / 'include'}", f"-I{PKG / 'host'}", f"-I{PKG / 'host' / 'compat'}", "-o", str(LIB), *srcs]
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB


def build_variant(name: str, defines: list[str]) -> Path:
    """A tuning variant of the same sources: ipt_b200/lib/variants/<name>.so built with extra -D flags."""
    out = PKG / "lib" / "variants" / f"{name}.so"
    out.parent.mkdir(parents=True, exist_ok=True)
    srcs = [str(s) for s in SOURCES if s.exists()]
    cmd = [find_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], f"-I{ROOT / 'include'}", f"-I{PKG / 'host'}", f"-I{PKG / 'host' / 'compat'}", "-o", str(out), *srcs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
