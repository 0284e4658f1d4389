"""CPU suite: the C-ABI library builds for sm_100a, loads, exports every symbol include/vh/abi.h
declares, agrees with the reference's struct layouts, and refuses to run without a GPU (no fallback)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from voxelhashing_demo_b200 import Config, Context, VHError
from voxelhashing_demo_b200 import lib as L

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(built_library):
    lib = L.load_library()
    header = (ROOT / "include" / "vh" / "abi.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", header))
    declared -= {"VH_REF", "defined", "sizeof"}
    declared = {d for d in declared if not d.isupper()}
    assert declared, "no prototypes parsed"
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, f"declared in abi.h but not exported: {missing}"
    # and the python-side lists cover the header
    assert declared <= set(L.LEGACY_SYMBOLS) | set(L.HANDLE_SYMBOLS)
    nm = subprocess.run(["nm", "-D", "--defined-only", str(L.LIB_PATH)], capture_output=True, text=True).stdout
    for name in L.LEGACY_SYMBOLS:                                  # unmangled: C linkage, as the reference's callers bind them
        assert re.search(rf"\sT {name}$", nm, flags=re.M), name
    # the C++ host classes are exported too (mangled)
    for cls in ("SDF_Hashtable", "CameraTracking", "Solver"):
        assert re.search(rf"\sT _ZN\d+{cls}", nm), cls


def test_struct_layouts_match_reference():
    """SURVEY.md Appendix A.1 (measured from the reference headers with nvcc 12.9 / gcc 13.3)."""
    assert C.sizeof(L.HashTableParams) == 176
    off = {n: getattr(L.HashTableParams, n).offset for n, _ in L.HashTableParams._fields_}
    assert (off["numBuckets"], off["voxelBlockSize"], off["truncation"], off["integrationWeightMax"]) == (128, 144, 164, 172)
    assert L.VOXEL_ENTRY_DTYPE.itemsize == 20 and L.VOXEL_ENTRY_DTYPE.fields["ptr"][1] == 12 and L.VOXEL_ENTRY_DTYPE.fields["offset"][1] == 16
    assert L.VOXEL_DTYPE.itemsize == 8 and C.sizeof(L.VhIcpSystem) == 128


def test_header_compiles_as_c_and_cxx(tmp_path):
    """abi.h is a C header (plain pointers and sizes) and a C++ header; static_asserts pin the layouts."""
    (tmp_path / "c.c").write_text('#include "vh/abi.h"\nint main(void){ vh_config c; vh_default_config(&c); return (int)sizeof(HashTableParams) - 176; }\n')
    (tmp_path / "cc.cpp").write_text(
        '#include "SDF_Hashtable.h"\n#include "CameraTracking.h"\n'
        "static_assert(sizeof(float4x4) == 64, \"\");\n"
        "int main(){ float4x4 m; m.setIdentity(); float4x4 i = m.getInverse(); return i.m11 == 1.0f ? 0 : 1; }\n")
    inc = ["-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include"]
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-fsyntax-only", *inc, str(tmp_path / "c.c")], check=True)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", *inc, str(tmp_path / "cc.cpp")], check=True)


def test_default_config_is_the_reference_defaults(built_library):
    lib = L.load_library()
    c = L.VhConfig()
    lib.vh_default_config(C.byref(c))
    t = c.table                                                      # common.h:39-50
    assert (t.numBuckets, t.bucketSize, t.attachedLinkedListSize, t.numVoxelBlocks, t.voxelBlockSize) == (5000, 5, 4, 1000, 8)
    assert (t.voxelSize, t.truncation, t.truncScale, t.integrationWeightMax) == (np.float32(0.02), 1.0, np.float32(0.01), 255.0)
    assert (c.width, c.height, c.depthScale, c.icpIterations) == (640, 480, 5000.0, 20)
    assert np.allclose([c.fx, c.fy, c.cx, c.cy], [517.3, 516.5, 318.6, 255.3]) and c.icpDistThres == np.float32(0.08)
    assert list(t.global_transform.entries) == list(np.eye(4).reshape(-1))


def test_no_gpu_means_loud_failure(built_library):
    """The product path has no CPU fallback: without a device, creating a context raises."""
    lib = L.load_library()
    if lib.vh_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(VHError, match="no CUDA device"):
        Context(Config())
    h = C.c_void_p()
    c = Config().to_c()
    assert lib.vh_create(C.byref(c), C.byref(h)) == L.VH_ERR_NO_DEVICE and not h


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may reach into oracle/."""
    pkg = ROOT / "voxelhashing_demo_b200"
    offenders = []
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")) + \
            list(pkg.rglob("*.cpp")) + list((ROOT / "include").rglob("*.h")):
        txt = f.read_text(errors="replace")
        if re.search(r"vh_oracle|libvh_oracle|from oracle|import oracle|oracle/|libvh_ref", txt):
            offenders.append(str(f.relative_to(ROOT)))
    assert not offenders, offenders
    nm = subprocess.run(["nm", "-D", str(L.LIB_PATH)], capture_output=True, text=True).stdout
    assert "vo_" not in nm


def test_sass_has_the_blackwell_paths(built_library):
    """The sm_100a instructions the design depends on are in the shipped SASS (excerpt: profiles/r2_sass_excerpt.txt):
    allocation claims a slot with ONE 128-bit CAS; integration moves a thread's four voxels with ONE 256-bit access
    (sm_100+) and evaluates them with packed fp32x2 arithmetic (FFMA2, sm_100+); allocation de-duplicates with MATCH.ANY;
    the persistent Align kernel exchanges {value, sequence} words with 64-bit strong (relaxed-scope) loads / stores --
    .GPU inside the grid, .SYS across NVLink -- and contains no ticket atomics."""
    import re

    sass = subprocess.run(["cuobjdump", "-sass", str(L.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass

    def count(pat):
        return len(re.findall(pat, sass))

    assert count(r"ATOMG\.E\.CAS\.128") >= 2
    assert count(r"LDG\.E\.[A-Z0-9.]*256") >= 8 and count(r"STG\.E\.[A-Z0-9.]*256") >= 8
    assert count(r"\bFFMA2\b") >= 60
    assert count(r"MATCH\.ANY") >= 4 and count(r"\bREDUX\b") >= 1
    assert count(r"LDG\.E\.64\.STRONG\.GPU") >= 8 and count(r"STG\.E\.64\.STRONG\.GPU") >= 4
    assert count(r"LDG\.E\.64\.STRONG\.SYS") >= 8 and count(r"STG\.E\.64\.STRONG\.SYS") >= 8
    # the Align kernels: the exchange is store + poll -- the only global atomic is the time-out counter of the cross-GPU wait
    align = re.findall(r"Function : \S*k_icp_align\S*\n(.*?)(?=\n\s*Function : |\Z)", sass, flags=re.S)
    assert len(align) == 4 and all(a.count("ATOMG") + a.count("RED.E") <= 1 for a in align)


def test_invalid_configs_are_rejected_before_touching_a_device(built_library):
    lib = L.load_library()
    for kw in (dict(numBuckets=0), dict(numVoxelBlocks=0), dict(voxelSize=0.0), dict(width=0), dict(numVoxelBlocks=5_000_000),
               dict(partCount=4, partRank=4), dict(partCount=2, partRank=-1)):
        h = C.c_void_p()
        c = Config(**kw).to_c()
        assert lib.vh_create(C.byref(c), C.byref(h)) == L.VH_ERR_INVALID and not h, kw
        assert lib.vh_last_error()
    c = Config().to_c()
    c.table.voxelBlockSize = 4
    h = C.c_void_p()
    assert lib.vh_create(C.byref(c), C.byref(h)) == L.VH_ERR_INVALID
    assert lib.vh_create(None, C.byref(h)) == L.VH_ERR_INVALID
    # every handle call refuses a null context instead of crashing
    null = C.c_void_p()
    assert lib.vh_compact(null, None) == L.VH_ERR_INVALID
    assert lib.vh_alloc_blocks(null, None, None, None) == L.VH_ERR_INVALID
    assert lib.vh_icp_align(null, None, None, None, None, 1, None) == L.VH_ERR_INVALID
    assert lib.vh_raycast(null, None, None, None) == L.VH_ERR_INVALID
    assert lib.vh_set_peers(null, 0, 1, None) == L.VH_ERR_INVALID


def test_build_is_safe_from_several_processes_and_ignores_mtimes(built_library, tmp_path):
    """One rank per GPU calls build() at the same moment on a box whose snapshot does not preserve file times:
    nothing may be re-linked (content stamps), and concurrent callers must not trip over each other (file lock)."""
    import os
    import sys
    import time

    from voxelhashing_demo_b200 import _build

    before = (_build.LIB.stat().st_ino, _build.LIB.stat().st_size)
    for o in _build.OBJ.glob("*.o"):                      # make every object look newer than the library
        os.utime(o, (time.time() + 100, time.time() + 100))
    code = "import sys; sys.path.insert(0, %r); from voxelhashing_demo_b200 import _build; _build.build()" % str(ROOT)
    procs = [subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE) for _ in range(4)]
    for p in procs:
        out, err = p.communicate(timeout=600)
        assert p.returncode == 0, err.decode()[-2000:]
    assert (_build.LIB.stat().st_ino, _build.LIB.stat().st_size) == before, "the library was re-linked although nothing changed"
    L.load_library()


def test_host_se3_functions_match_the_oracle(built_library, oracle, tmp_path):
    """SE3Exp / SE3Log / updateTransform of the C++ host surface (ref SE3.h:6-8, SE3.cpp:4-26) are pure host code:
    a small C++ program linked against the library, run here without a GPU, must agree with the oracle's closed forms."""
    from voxelhashing_demo_b200 import _build

    src = tmp_path / "se3.cpp"
    src.write_text(r'''
#include <cstdio>
#include "SE3.h"
int main() {
    const float tw[3][6] = {{0, 0, 0, 0, 0, 1.57079632679f}, {0.01f, -0.004f, 0.002f, 0.003f, -0.002f, 0.001f}, {0.3f, -0.2f, 0.5f, 0.4f, 0.1f, -0.7f}};
    for (int k = 0; k < 3; ++k) {
        Vector6f t; for (int i = 0; i < 6; ++i) t(i) = tw[k][i];
        Matrix4x4f M = SE3Exp(t);
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) std::printf("%.9g ", M(r, c));
        Vector6f back = SE3Log(M);
        for (int i = 0; i < 6; ++i) std::printf("%.9g ", back(i));
        std::printf("\n");
    }
    Vector6f a, b; for (int i = 0; i < 6; ++i) { a(i) = tw[1][i]; b(i) = tw[2][i]; }
    Vector6f u = updateTransform(a, b);
    for (int i = 0; i < 6; ++i) std::printf("%.9g ", u(i));
    std::printf("\n");
    return 0;
}
''')
    exe = tmp_path / "se3"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include", str(src), "-o", str(exe),
                    "-L", str(_build.PKG), "-lvh_b200", f"-Wl,-rpath,{_build.PKG}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    twists = [[0, 0, 0, 0, 0, np.pi / 2], [0.01, -0.004, 0.002, 0.003, -0.002, 0.001], [0.3, -0.2, 0.5, 0.4, 0.1, -0.7]]
    for line, tw in zip(out[:3], twists):
        vals = np.array([float(x) for x in line.split()])
        M, back = vals[:16].reshape(4, 4), vals[16:]
        assert np.allclose(M, oracle.se3_exp(tw), atol=2e-6)
        assert np.allclose(back, tw, atol=2e-6)                                   # log(exp(t)) = t
    u = np.array([float(x) for x in out[3].split()])
    want = oracle.se3_log(oracle.se3_exp(twists[1]).astype(np.float64) @ oracle.se3_exp(twists[2]).astype(np.float64))
    assert np.allclose(u, want, atol=5e-6)                                        # estimate <- log(exp(x) exp(estimate)), Solver.cpp:110-111
