"""ctypes binding of oracle/libvh_oracle.so and (GPU box only) oracle/_ref/libvh_ref.so.

TEST INFRASTRUCTURE, NOT PRODUCT.  Imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing in voxelhashing_demo_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_LIB = HERE / "libvh_oracle.so"
REF_LIB = HERE / "_ref" / "libvh_ref.so"

REF_EXACT, FIXED = 0, 1


class VoConfig(C.Structure):
    _fields_ = [
        ("policy", C.c_int), ("width", C.c_int), ("height", C.c_int),
        ("K", C.c_float * 9), ("Kinv", C.c_float * 9),
        ("depthScale", C.c_float), ("depthMin", C.c_float), ("depthMax", C.c_float),
        ("numBuckets", C.c_uint), ("bucketSize", C.c_uint), ("attachedLinkedListSize", C.c_uint), ("numVoxelBlocks", C.c_uint),
        ("overflowSlots", C.c_uint),
        ("voxelSize", C.c_float), ("truncation", C.c_float), ("truncScale", C.c_float), ("maxIntegrationDistance", C.c_float),
        ("integrationWeightSample", C.c_uint), ("integrationWeightMax", C.c_float),
        ("icpDistThres", C.c_float), ("icpNormalThres", C.c_float),
        ("partCount", C.c_int), ("partRank", C.c_int),
        ("bilateralSigmaSpace", C.c_float), ("bilateralSigmaRange", C.c_float),
    ]


class VoAllocReport(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("requestedPixels", "requestedBlocks", "requestedNew", "inserted", "bucketsTouched",
                                      "bucketsContended", "maxNewPerBucket", "dropped")]


class VoIcpSystem(C.Structure):
    _fields_ = [("JtJ", C.c_float * 21), ("Jtr", C.c_float * 6), ("error", C.c_float), ("count", C.c_float), ("pad", C.c_float * 3)]


def build_oracle(force: bool = False) -> Path:
    """Rebuild decisions go by a content stamp (the snapshot sent to the GPU box does not preserve file times), under a
    file lock (several processes may import the checker at once)."""
    import fcntl
    import hashlib

    want = hashlib.sha256((HERE / "vh_oracle.cpp").read_bytes() + (HERE / "vh_oracle.h").read_bytes()
                          + (HERE / "Makefile").read_bytes()).hexdigest()
    stamp = HERE / ".libvh_oracle.stamp"
    with open(HERE / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if force or not ORACLE_LIB.exists() or not stamp.exists() or stamp.read_text() != want:
            subprocess.run(["make", "-B", "-C", str(HERE), "oracle"], check=True, capture_output=True)
            stamp.write_text(want)
    return ORACLE_LIB


_olib = None


def oracle_lib() -> C.CDLL:
    global _olib
    if _olib is not None:
        return _olib
    build_oracle()
    lib = C.CDLL(str(ORACLE_LIB), mode=C.RTLD_LOCAL)
    P, I, F = C.c_void_p, C.c_int, C.c_float
    CP = C.POINTER(VoConfig)
    lib.vo_num_threads.restype = I
    lib.vo_set_num_threads.argtypes = [I]
    lib.vo_create.argtypes = [CP]
    lib.vo_create.restype = P
    lib.vo_destroy.argtypes = [P]
    lib.vo_destroy.restype = None
    lib.vo_reset.argtypes = [P]
    lib.vo_reset.restype = None
    lib.vo_preprocess.argtypes = [CP, P, P, P, P]
    lib.vo_preprocess.restype = None
    lib.vo_alloc.argtypes = [P, P, P, C.POINTER(VoAllocReport)]
    lib.vo_alloc.restype = None
    lib.vo_last_requested_new.argtypes = [P, P, I]
    lib.vo_last_requested_new.restype = I
    lib.vo_compact.argtypes = [P, P]
    lib.vo_compact.restype = I
    lib.vo_garbage_collect.argtypes = [P, I, C.c_float, C.c_float]
    lib.vo_garbage_collect.restype = I
    lib.vo_extract_mesh.argtypes = [P, P, I]
    lib.vo_extract_mesh.restype = I
    lib.vo_stream_out.argtypes = [P, P, C.c_float, P, P, I]
    lib.vo_stream_out.restype = I
    lib.vo_stream_in.argtypes = [P, P, P, I]
    lib.vo_stream_in.restype = I
    lib.vo_integrate.argtypes = [P, P, P]
    lib.vo_integrate.restype = C.c_longlong
    lib.vo_integrate_depthf.argtypes = [P, P, P]
    lib.vo_integrate_depthf.restype = C.c_longlong
    lib.vo_num_allocated.argtypes = [P]
    lib.vo_num_allocated.restype = I
    lib.vo_export_entries.argtypes = [P, P, I]
    lib.vo_export_entries.restype = I
    lib.vo_export_compact.argtypes = [P, P, I]
    lib.vo_export_compact.restype = I
    lib.vo_get_block.argtypes = [P, I, I, I, P]
    lib.vo_get_block.restype = I
    lib.vo_heap_counter.argtypes = [P]
    lib.vo_heap_counter.restype = I
    lib.vo_hash.argtypes = [CP, I, I, I]
    lib.vo_hash.restype = C.c_uint
    lib.vo_world2block.argtypes = [CP, P, P]
    lib.vo_world2block.restype = None
    lib.vo_block_in_frustum.argtypes = [CP, P, I, I, I]
    lib.vo_block_in_frustum.restype = I
    lib.vo_find_correspondences.argtypes = [CP, P, P, P, P, P, P, P, P]
    lib.vo_find_correspondences.restype = F
    lib.vo_jacobians.argtypes = [CP, P, P, P]
    lib.vo_jacobians.restype = None
    lib.vo_icp_system_build.argtypes = [CP, P, P, P, P, P, I, I, C.POINTER(VoIcpSystem)]
    lib.vo_icp_system_build.restype = None
    lib.vo_icp_solve.argtypes = [P, P, P]
    lib.vo_icp_solve.restype = I
    lib.vo_icp_align.argtypes = [CP, P, P, P, P, I, P, P]
    lib.vo_icp_align.restype = I
    lib.vo_se3_exp.argtypes = [P, P]
    lib.vo_se3_exp.restype = None
    lib.vo_se3_log.argtypes = [P, P]
    lib.vo_se3_log.restype = None
    lib.vo_mat4_inverse.argtypes = [P, P]
    lib.vo_mat4_inverse.restype = None
    lib.vo_raycast.argtypes = [P, P, P, P]
    lib.vo_raycast.restype = None
    _olib = lib
    return lib


def f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def make_config(cfg) -> VoConfig:
    """VoConfig from a voxelhashing_demo_b200.fusion.Config-like object (attribute names shared)."""
    c = VoConfig()
    c.policy, c.width, c.height = cfg.policy, cfg.width, cfg.height
    K = np.array([cfg.fx, 0, cfg.cx, 0, cfg.fy, cfg.cy, 0, 0, 1], dtype=np.float32)
    f = np.float32
    fx, fy, cx, cy = f(cfg.fx), f(cfg.fy), f(cfg.cx), f(cfg.cy)
    Kinv = np.array([f(1) / fx, 0, -cx / fx, 0, f(1) / fy, -cy / fy, 0, 0, 1], dtype=np.float32)
    for i in range(9):
        c.K[i] = float(K[i])
        c.Kinv[i] = float(Kinv[i])
    c.depthScale, c.depthMin, c.depthMax = cfg.depthScale, cfg.depthMin, cfg.depthMax
    c.numBuckets, c.bucketSize = cfg.numBuckets, cfg.bucketSize
    c.attachedLinkedListSize, c.numVoxelBlocks = cfg.attachedLinkedListSize, cfg.numVoxelBlocks
    c.overflowSlots = (cfg.overflowSlots or cfg.numBuckets) if cfg.policy == FIXED else 0
    c.voxelSize, c.truncation, c.truncScale = cfg.voxelSize, cfg.truncation, cfg.truncScale
    c.maxIntegrationDistance = cfg.maxIntegrationDistance
    c.integrationWeightSample, c.integrationWeightMax = cfg.integrationWeightSample, cfg.integrationWeightMax
    c.icpDistThres, c.icpNormalThres = cfg.icpDistThres, cfg.icpNormalThres
    c.partCount, c.partRank = max(1, cfg.partCount), cfg.partRank
    c.bilateralSigmaSpace, c.bilateralSigmaRange = cfg.bilateralSigmaSpace, cfg.bilateralSigmaRange
    return c


class OracleTable:
    """CPU hash table + voxel heap following the reference's algorithm (vh_oracle.cpp)."""

    def __init__(self, cfg):
        self.lib = oracle_lib()
        self.ccfg = make_config(cfg)
        self.cfg = cfg
        self.h = self.lib.vo_create(C.byref(self.ccfg))
        if not self.h:
            raise MemoryError("vo_create failed")

    def close(self):
        if self.h:
            self.lib.vo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self.lib.vo_reset(self.h)

    def preprocess(self, depth_u16: np.ndarray):
        n = self.cfg.width * self.cfg.height
        d = np.ascontiguousarray(depth_u16, dtype=np.uint16).reshape(-1)
        verts = np.zeros((n, 4), np.float32)
        normals = np.zeros((n, 4), np.float32)
        depthf = np.zeros(n, np.float32)
        self.lib.vo_preprocess(C.byref(self.ccfg), d.ctypes.data, verts.ctypes.data, normals.ctypes.data, depthf.ctypes.data)
        return verts, normals, depthf

    def alloc(self, pose, verts) -> VoAllocReport:
        rep = VoAllocReport()
        p, v = f32(pose).reshape(16), f32(verts)
        self.lib.vo_alloc(self.h, p.ctypes.data, v.ctypes.data, C.byref(rep))
        return rep

    def last_requested_new(self) -> set:
        n = self.lib.vo_last_requested_new(self.h, 0, 0)
        out = np.zeros((max(n, 1), 3), np.int32)
        self.lib.vo_last_requested_new(self.h, out.ctypes.data, n)
        return {tuple(int(x) for x in r) for r in out[:n]}

    def compact(self, pose) -> int:
        p = f32(pose).reshape(16)
        return int(self.lib.vo_compact(self.h, p.ctypes.data))

    def integrate(self, pose, verts) -> int:
        p, v = f32(pose).reshape(16), f32(verts)
        return int(self.lib.vo_integrate(self.h, p.ctypes.data, v.ctypes.data))

    def integrate_depthf(self, pose, depthf) -> int:
        p, v = f32(pose).reshape(16), f32(depthf)
        return int(self.lib.vo_integrate_depthf(self.h, p.ctypes.data, v.ctypes.data))

    def garbage_collect(self, scope=0, sdf_threshold=0.0, weight_decay=0.0) -> int:
        return int(self.lib.vo_garbage_collect(self.h, int(scope), float(sdf_threshold), float(weight_decay)))

    def extract_mesh(self) -> np.ndarray:
        """-> float32 [n, 3, 3]: triangles of the zero level set (marching tetrahedra, world metres)."""
        n = int(self.lib.vo_extract_mesh(self.h, None, 0))
        tris = np.zeros((max(n, 1), 3, 3), np.float32)
        n = int(self.lib.vo_extract_mesh(self.h, tris.ctypes.data, n))
        return tris[:n]

    def stream_out(self, center, radius, capacity):
        """-> (entries [n,5] int32, voxels [n,512,2] float32) of the blocks farther than radius from center; they leave the table."""
        ent = np.zeros((max(capacity, 1), 5), np.int32)
        vox = np.zeros((max(capacity, 1), 512, 2), np.float32)
        c = f32(center).reshape(3)
        n = int(self.lib.vo_stream_out(self.h, c.ctypes.data, float(radius), ent.ctypes.data, vox.ctypes.data, int(capacity)))
        return ent[:n].copy(), vox[:n].copy()

    def stream_in(self, entries, voxels) -> int:
        ent = np.ascontiguousarray(entries, np.int32).reshape(-1, 5)
        vox = np.ascontiguousarray(voxels, np.float32)
        return int(self.lib.vo_stream_in(self.h, ent.ctypes.data, vox.ctypes.data, len(ent)))

    def fuse_frame(self, pose, verts, depthf=None):
        rep = self.alloc(pose, verts)
        nvis = self.compact(pose)
        nupd = self.integrate_depthf(pose, depthf) if depthf is not None else self.integrate(pose, verts)
        return rep, nvis, nupd

    def entries(self) -> np.ndarray:
        n = self.lib.vo_num_allocated(self.h)
        out = np.zeros((max(n, 1), 5), np.int32)
        self.lib.vo_export_entries(self.h, out.ctypes.data, n)
        return out[:n]

    def compact_entries(self) -> np.ndarray:
        n = self.lib.vo_export_compact(self.h, 0, 0)
        out = np.zeros((max(n, 1), 5), np.int32)
        self.lib.vo_export_compact(self.h, out.ctypes.data, n)
        return out[:n]

    def block(self, x, y, z):
        out = np.zeros((512, 2), np.float32)
        ok = self.lib.vo_get_block(self.h, int(x), int(y), int(z), out.ctypes.data)
        return out if ok else None

    def block_dict(self) -> dict:
        return {(int(e[0]), int(e[1]), int(e[2])): self.block(e[0], e[1], e[2]) for e in self.entries()}

    def heap_counter(self) -> int:
        return int(self.lib.vo_heap_counter(self.h))

    def raycast(self, pose):
        n = self.cfg.width * self.cfg.height
        verts = np.zeros((n, 4), np.float32)
        normals = np.zeros((n, 4), np.float32)
        p = f32(pose).reshape(16)
        self.lib.vo_raycast(self.h, p.ctypes.data, verts.ctypes.data, normals.ctypes.data)
        return verts, normals


# ---- free functions --------------------------------------------------------------------------------
def find_correspondences(cfg, inp, inpN, tgt, tgtN, delta):
    lib = oracle_lib()
    cc = make_config(cfg)
    n = cfg.width * cfg.height
    corr, corrN, res = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32), np.zeros(n, np.float32)
    a, b, c_, d = f32(inp), (None if inpN is None else f32(inpN)), f32(tgt), f32(tgtN)
    dl = f32(delta).reshape(16)
    err = lib.vo_find_correspondences(C.byref(cc), a.ctypes.data, 0 if b is None else b.ctypes.data, c_.ctypes.data, d.ctypes.data,
                                      dl.ctypes.data, corr.ctypes.data, corrN.ctypes.data, res.ctypes.data)
    return float(err), corr, corrN, res


def jacobians(cfg, corr, corrN):
    lib = oracle_lib()
    cc = make_config(cfg)
    J = np.zeros((cfg.width * cfg.height, 6), np.float32)
    a, b = f32(corr), f32(corrN)
    lib.vo_jacobians(C.byref(cc), a.ctypes.data, b.ctypes.data, J.ctypes.data)
    return J


def icp_system(cfg, inp, inpN, tgt, tgtN, delta, row0=0, row1=None) -> np.ndarray:
    lib = oracle_lib()
    cc = make_config(cfg)
    s = VoIcpSystem()
    a, b, c_, d = f32(inp), (None if inpN is None else f32(inpN)), f32(tgt), f32(tgtN)
    dl = f32(delta).reshape(16)
    lib.vo_icp_system_build(C.byref(cc), a.ctypes.data, 0 if b is None else b.ctypes.data, c_.ctypes.data, d.ctypes.data,
                            dl.ctypes.data, row0, cfg.height if row1 is None else row1, C.byref(s))
    return np.frombuffer(bytes(s), dtype=np.float32).copy()


def icp_solve(system32, estimate6, delta16=None):
    lib = oracle_lib()
    s = f32(system32).copy()
    est = f32(estimate6).copy()
    dl = np.eye(4, dtype=np.float32).reshape(16) if delta16 is None else f32(delta16).reshape(16).copy()
    ok = lib.vo_icp_solve(s.ctypes.data, est.ctypes.data, dl.ctypes.data)
    return bool(ok), est, dl.reshape(4, 4)


def icp_align(cfg, inp, inpN, tgt, tgtN, iterations, estimate6=None):
    lib = oracle_lib()
    cc = make_config(cfg)
    est = np.zeros(6, np.float32) if estimate6 is None else f32(estimate6).copy()
    dl = se3_exp(est).reshape(16).copy()
    a, b, c_, d = f32(inp), (None if inpN is None else f32(inpN)), f32(tgt), f32(tgtN)
    n = lib.vo_icp_align(C.byref(cc), a.ctypes.data, 0 if b is None else b.ctypes.data, c_.ctypes.data, d.ctypes.data,
                         iterations, est.ctypes.data, dl.ctypes.data)
    return int(n), est, dl.reshape(4, 4)


def se3_exp(twist6) -> np.ndarray:
    lib = oracle_lib()
    t = f32(twist6)
    m = np.zeros(16, np.float32)
    lib.vo_se3_exp(t.ctypes.data, m.ctypes.data)
    return m.reshape(4, 4)


def se3_log(m44) -> np.ndarray:
    lib = oracle_lib()
    m = f32(m44).reshape(16)
    t = np.zeros(6, np.float32)
    lib.vo_se3_log(m.ctypes.data, t.ctypes.data)
    return t


def mat4_inverse(m44) -> np.ndarray:
    lib = oracle_lib()
    m = f32(m44).reshape(16)
    o = np.zeros(16, np.float32)
    lib.vo_mat4_inverse(m.ctypes.data, o.ctypes.data)
    return o.reshape(4, 4)


def hash_block(cfg, x, y, z) -> int:
    cc = make_config(cfg)
    return int(oracle_lib().vo_hash(C.byref(cc), int(x), int(y), int(z)))


def world2block(cfg, p3):
    cc = make_config(cfg)
    p = f32(p3)
    b = np.zeros(3, np.int32)
    oracle_lib().vo_world2block(C.byref(cc), p.ctypes.data, b.ctypes.data)
    return tuple(int(v) for v in b)


def block_in_frustum(cfg, pose, x, y, z) -> bool:
    cc = make_config(cfg)
    p = f32(pose).reshape(16)
    return bool(oracle_lib().vo_block_in_frustum(C.byref(cc), p.ctypes.data, int(x), int(y), int(z)))


# ---- reference CUDA harness (GPU box only) --------------------------------------------------------------
_rlib = {}


def ref_lib(variant: str = "") -> C.CDLL:
    """oracle/_ref/libvh_ref[_<variant>].so: the reference's own .cu files + ref_harness.cu. Needs a GPU to run.
    variant "" = -O3 as BASELINE.md 3.1 states; "noprintf" = the two device printf calls of insertVoxelEntry deleted;
    "G" = the shipped -G (device debug) flags.  One variant per process: they share the reference's global symbols."""
    if variant in _rlib:
        return _rlib[variant]
    path = REF_LIB if not variant else REF_LIB.with_name(f"libvh_ref_{variant}.so")
    if not path.exists():
        raise FileNotFoundError(f"{path} not built (make -C oracle ref; needs /root/reference)")
    lib = C.CDLL(str(path), mode=C.RTLD_LOCAL)
    P, I, F = C.c_void_p, C.c_int, C.c_float
    lib.ref_init.argtypes = [I, I, I, F, F]
    lib.ref_init.restype = I
    lib.ref_set_intrinsic.argtypes = [P, P]
    lib.ref_set_intrinsic.restype = None
    lib.ref_preprocess.argtypes = [P, P, P]
    lib.ref_preprocess.restype = None
    lib.ref_integrate.argtypes = [P, P, P]
    lib.ref_integrate.restype = I
    lib.ref_stage_begin.argtypes = [P]
    lib.ref_stage_begin.restype = None
    lib.ref_stage_alloc.argtypes = [P, P]
    lib.ref_stage_alloc.restype = None
    lib.ref_stage_compact.restype = I
    lib.ref_stage_integrate.argtypes = [P]
    lib.ref_stage_integrate.restype = None
    lib.ref_num_slots.restype = I
    lib.ref_export_table.argtypes = [P, I]
    lib.ref_export_table.restype = I
    lib.ref_export_compact.argtypes = [P, I]
    lib.ref_export_compact.restype = I
    lib.ref_export_block.argtypes = [I, P]
    lib.ref_export_block.restype = None
    lib.ref_heap_counter.restype = I
    lib.ref_correspond.argtypes = [P, P, P, P, P, P, P]
    lib.ref_correspond.restype = F
    lib.ref_build_system.argtypes = [P, P, P, P]
    lib.ref_build_system.restype = None
    lib.ref_align.argtypes = [P, P, P, I, P, P, P]
    lib.ref_align.restype = I
    _rlib[variant] = lib
    return lib


def ref_solve_callback():
    """C function pointer to the oracle's vo_icp_solve, handed to ref_align for the Eigen half."""
    return C.cast(oracle_lib().vo_icp_solve, C.c_void_p)
