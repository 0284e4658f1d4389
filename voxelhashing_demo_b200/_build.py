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
DIST_DEMO = PKG / "vh_dist_app"

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
    # flags with the checkout path factored out: the snapshot on the GPU box lives under another root, and an
    # identical tree must not look stale there
    h.update(" ".join(NVCC_FLAGS).replace(str(ROOT), "<root>").encode())
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


def _link_if_stale(out: Path, stamp_file: Path, want: str, cmd: list[str], what: str, force: bool) -> None:
    """Link decisions go by CONTENT stamps, never by mtimes: the gpurun snapshot does not preserve file times, and a
    box that re-linked because 'the objects look newer' did so from all ranks at once."""
    if not force and out.exists() and stamp_file.exists() and stamp_file.read_text() == want:
        return
    tmp = out.with_name(out.name + f".tmp{os.getpid()}")
    r = subprocess.run([c if c != str(out) else str(tmp) for c in cmd], capture_output=True, text=True)
    if r.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError(f"{what} failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, out)                         # atomic: a concurrent reader sees the old or the new file, never half of one
    stamp_file.write_text(want)


def build(verbose: bool = False, force: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link libvh_b200.so in-tree.  Safe to call from several processes at
    once (one rank per GPU does): the whole step runs under an exclusive file lock and is a no-op when the stamps match."""
    import fcntl

    PKG.mkdir(exist_ok=True)
    with open(PKG / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(verbose, force)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool, force: bool) -> Path:
    if force and OBJ.exists():
        shutil.rmtree(OBJ)
    OBJ.mkdir(exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    lib_want = hashlib.sha256("".join((OBJ / (s.stem + ".stamp")).read_text() for s in srcs).encode()).hexdigest()
    _link_if_stale(LIB, OBJ / "libvh_b200.stamp", lib_want,
                   [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", HOSTCXX,
                    "-o", str(LIB), *map(str, objs), "-lcudart_static", "-lrt", "-ldl", "-lpthread", "-lz"], "link", force)
    demo_src = ROOT / "examples" / "headless_app.cpp"
    if demo_src.exists():
        demo_want = hashlib.sha256((lib_want + hashlib.sha256(demo_src.read_bytes()).hexdigest()).encode()).hexdigest()
        _link_if_stale(HOST_DEMO, OBJ / "vh_headless_app.stamp", demo_want,
                       [NVCC, "-std=c++17", "-O2", "-ccbin", HOSTCXX, "-I", str(ROOT / "include"), str(demo_src), "-o", str(HOST_DEMO),
                        "-L", str(PKG), "-lvh_b200", "-Xlinker", f"-rpath={PKG}", "-Xlinker", "-rpath=$ORIGIN"], "headless_app build", force)
    dist_src = ROOT / "examples" / "dist_app.cpp"
    if dist_src.exists():                          # the multi-GPU frame loop from C++ (vh_dist_*): one process per GPU
        dist_want = hashlib.sha256((lib_want + hashlib.sha256(dist_src.read_bytes()).hexdigest()).encode()).hexdigest()
        _link_if_stale(DIST_DEMO, OBJ / "vh_dist_app.stamp", dist_want,
                       [NVCC, "-std=c++17", "-O2", "-ccbin", HOSTCXX, "-I", str(ROOT / "include"), str(dist_src), "-o", str(DIST_DEMO),
                        "-L", str(PKG), "-lvh_b200", "-Xlinker", f"-rpath={PKG}", "-Xlinker", "-rpath=$ORIGIN"], "dist_app build", force)
    return LIB


if __name__ == "__main__":
    path = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(path)
