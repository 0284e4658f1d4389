"""In-tree build of libvh_b200.so (sm_100a) with plain nvcc -- no JIT cache, no torch extension.

The built library lands next to this file (git-ignored, but it travels to the GPU box with the
gpurun snapshot).  `python -m voxelhashing_demo_b200._build` rebuilds; `build()` is what
`__graft_entry__.build()` calls.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = PKG / "_obj"
LIB = PKG / "libvh_b200.so"
HOST_DEMO = PKG / "vh_headless_app"

NVCC = os.environ.get("VH_NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

# -fmad=false: the reference's own flag (CMakeLists.txt:23); fused multiply-adds only where the
# source says fmaf().  -lineinfo so ncu's source page maps back to these files.
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-fmad=false", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-ccbin", HOSTCXX,
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=default",
    "-I", str(ROOT / "include"), "-I", str(CSRC),
] + os.environ.get("VH_EXTRA_NVCC_FLAGS", "").split()      # e.g. -DVH_ICP_TRACE for tools/icp_trace.py


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu")) + sorted((CSRC / "host").glob("*.cpp"))


def _stamp(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").rglob("*.h")):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def _compile(src: Path, verbose: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    stamp = OBJ / (src.stem + ".stamp")
    want = _stamp(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == want:
        return obj
    cmd = [NVCC, *NVCC_FLAGS, "-x", "cu", "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    stamp.write_text(want)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link libvh_b200.so in-tree."""
    if force and OBJ.exists():
        shutil.rmtree(OBJ)
    OBJ.mkdir(exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", HOSTCXX,
               "-o", str(LIB), *map(str, objs), "-lcudart_static", "-lrt", "-ldl", "-lpthread", "-lz"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    demo_src = ROOT / "examples" / "headless_app.cpp"
    if demo_src.exists() and (force or not HOST_DEMO.exists() or HOST_DEMO.stat().st_mtime < max(LIB.stat().st_mtime, demo_src.stat().st_mtime)):
        cmd = [NVCC, "-std=c++17", "-O2", "-ccbin", HOSTCXX, "-I", str(ROOT / "include"), str(demo_src), "-o", str(HOST_DEMO),
               "-L", str(PKG), "-lvh_b200", "-Xlinker", f"-rpath={PKG}", "-Xlinker", "-rpath=$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"headless_app build failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(path)
