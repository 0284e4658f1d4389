"""Host-side mirror of the reference's operator interface, over the C ABI.

`Context` wraps one `vh_context` (handle API).  `SDF_Hashtable` and `CameraTracking` keep the
reference's class and method names (SDF_Hashtable.h:24-42, CameraTracking.h:34-59) so parity tests
read like the reference's own call sites (Application.cpp:73-84).  `FramePipeline` wraps the native
CUDA-graph frame loop.  torch is used for device memory and streams only; every computation is a
call into libvh_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import lib as L


def _torch():
    import torch

    return torch


def _ptr(t) -> int:
    """Device (or host) address of a torch tensor / numpy array / int / None."""
    if t is None:
        return 0
    if isinstance(t, int):
        return t
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


def _stream(stream=None) -> int:
    if stream is None:
        return _torch().cuda.current_stream().cuda_stream
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream


def read_depth(path) -> np.ndarray:
    """16-bit depth image (PNG as the reference loads it with stbi_load_16, Application.cpp:28-29, or binary PGM)
    -> uint16 array [H, W].  Host only."""
    lib = L.load_library()
    p = C.POINTER(C.c_uint16)()
    w, h = C.c_int(), C.c_int()
    if lib.vh_depth_read(str(path).encode(), C.byref(p), C.byref(w), C.byref(h)) != L.VH_OK:
        raise L.VHError(f"vh_depth_read({path}): {lib.vh_depth_last_error().decode()}")
    try:
        return np.ctypeslib.as_array(p, shape=(h.value, w.value)).copy()
    finally:
        lib.vh_depth_free(p)


def write_depth_png(path, depth_u16: np.ndarray, filter_type: int = 4) -> None:
    a = np.ascontiguousarray(depth_u16, dtype=np.uint16)
    lib = L.load_library()
    if lib.vh_depth_write_png(str(path).encode(), a.ctypes.data, a.shape[1], a.shape[0], int(filter_type)) != L.VH_OK:
        raise L.VHError(f"vh_depth_write_png({path}): {lib.vh_depth_last_error().decode()}")


@dataclass
class Config:
    """vh_config with the reference's defaults (common.h:7-50)."""

    policy: int = L.POLICY_REF_EXACT
    width: int = 640
    height: int = 480
    fx: float = 517.3
    fy: float = 516.5
    cx: float = 318.6
    cy: float = 255.3
    depthScale: float = 5000.0
    depthMin: float = 0.1
    depthMax: float = 4.0
    numBuckets: int = 5000
    bucketSize: int = 5
    attachedLinkedListSize: int = 4
    numVoxelBlocks: int = 1000
    voxelSize: float = 0.02
    maxIntegrationDistance: float = 4.0
    truncScale: float = 0.01
    truncation: float = 1.0
    integrationWeightSample: int = 10
    integrationWeightMax: float = 255.0
    overflowSlots: int = 0
    icpDistThres: float = 0.08
    icpNormalThres: float = -1.0
    icpIterations: int = 20
    partCount: int = 1
    partRank: int = 0
    bilateralSigmaSpace: float = 0.0      # pixels; Fixed policy, 0 = off
    bilateralSigmaRange: float = 0.0      # metres
    extra: dict = field(default_factory=dict)

    def to_c(self) -> L.VhConfig:
        c = L.VhConfig()
        L.load_library().vh_default_config(C.byref(c))
        t = c.table
        t.numBuckets, t.bucketSize = self.numBuckets, self.bucketSize
        t.attachedLinkedListSize, t.numVoxelBlocks = self.attachedLinkedListSize, self.numVoxelBlocks
        t.voxelBlockSize, t.voxelSize = 8, self.voxelSize
        t.maxIntegrationDistance, t.truncScale, t.truncation = self.maxIntegrationDistance, self.truncScale, self.truncation
        t.integrationWeightSample, t.integrationWeightMax = self.integrationWeightSample, self.integrationWeightMax
        c.policy, c.width, c.height = self.policy, self.width, self.height
        c.fx, c.fy, c.cx, c.cy = self.fx, self.fy, self.cx, self.cy
        c.depthScale, c.depthMin, c.depthMax = self.depthScale, self.depthMin, self.depthMax
        c.overflowSlots = self.overflowSlots
        c.icpDistThres, c.icpNormalThres, c.icpIterations = self.icpDistThres, self.icpNormalThres, self.icpIterations
        c.partCount, c.partRank = self.partCount, self.partRank
        c.bilateralSigmaSpace, c.bilateralSigmaRange = self.bilateralSigmaSpace, self.bilateralSigmaRange
        return c

    def K(self) -> np.ndarray:
        return np.array([self.fx, 0, self.cx, 0, self.fy, self.cy, 0, 0, 1], dtype=np.float32)

    def Kinv(self) -> np.ndarray:
        """Closed-form inverse in fp32, the same expression the library uses (vh_abi.cu fillView)."""
        f = np.float32
        fx, fy, cx, cy = f(self.fx), f(self.fy), f(self.cx), f(self.cy)
        return np.array([f(1) / fx, 0, -cx / fx, 0, f(1) / fy, -cy / fy, 0, 0, 1], dtype=np.float32)


def _pose_arg(pose) -> np.ndarray:
    p = np.ascontiguousarray(np.asarray(pose, dtype=np.float32).reshape(16))
    return p


class Context:
    """One hash table + tracker on the current CUDA device (handle API of include/vh/abi.h)."""

    def __init__(self, cfg: Config):
        self.lib = L.load_library()
        self.cfg = cfg
        self._h = C.c_void_p()
        ccfg = cfg.to_c()
        L.check(self.lib.vh_create(C.byref(ccfg), C.byref(self._h)), "vh_create")

    # -- lifetime -----------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.vh_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def reset(self, stream=None):
        L.check(self.lib.vh_reset(self._h, _stream(stream)), "vh_reset")

    def bytes_allocated(self) -> int:
        return int(self.lib.vh_bytes_allocated(self._h))

    def set_intrinsic_matrices(self, K9, Kinv9):
        K9 = np.ascontiguousarray(K9, dtype=np.float32)
        Kinv9 = np.ascontiguousarray(Kinv9, dtype=np.float32)
        L.check(self.lib.vh_set_intrinsic_matrices(self._h, K9.ctypes.data, Kinv9.ctypes.data))

    # -- tensors ------------------------------------------------------------------------------------
    def new_maps(self):
        """(verts, normals, depthf) device tensors of the context's image size."""
        torch = _torch()
        n = self.cfg.width * self.cfg.height
        return (
            torch.zeros((n, 4), dtype=torch.float32, device="cuda"),
            torch.zeros((n, 4), dtype=torch.float32, device="cuda"),
            torch.zeros((n,), dtype=torch.float32, device="cuda"),
        )

    # -- stages -------------------------------------------------------------------------------------
    def preprocess(self, depth_u16, verts, normals, depthf=None, stream=None):
        L.check(self.lib.vh_preprocess(self._h, _ptr(depth_u16), _ptr(verts), _ptr(normals), _ptr(depthf), _stream(stream)), "vh_preprocess")

    def set_pose(self, pose, stream=None):
        p = _pose_arg(pose)
        L.check(self.lib.vh_set_pose(self._h, p.ctypes.data, _stream(stream)), "vh_set_pose")

    def set_pose_device(self, d_pose, stream=None):
        L.check(self.lib.vh_set_pose_device(self._h, _ptr(d_pose), _stream(stream)), "vh_set_pose_device")

    def alloc_blocks(self, verts, normals=None, stream=None):
        L.check(self.lib.vh_alloc_blocks(self._h, _ptr(verts), _ptr(normals), _stream(stream)), "vh_alloc_blocks")

    def set_tuning(self, align_ctas: int = 0, fusion_reserved_sms: int = -1):
        L.check(self.lib.vh_set_tuning(self._h, align_ctas, fusion_reserved_sms), "vh_set_tuning")

    def alloc_blocks_depth(self, depth_u16, stream=None):
        """Allocation straight from the raw u16 depth image (2 B / pixel; same blocks as from the vertex map)."""
        L.check(self.lib.vh_alloc_blocks_depth(self._h, _ptr(depth_u16), _stream(stream)), "vh_alloc_blocks_depth")

    def compact(self, stream=None):
        L.check(self.lib.vh_compact(self._h, _stream(stream)), "vh_compact")

    def integrate(self, verts, stream=None):
        L.check(self.lib.vh_integrate(self._h, _ptr(verts), _stream(stream)), "vh_integrate")

    def integrate_depthf(self, depthf, stream=None):
        L.check(self.lib.vh_integrate_depthf(self._h, _ptr(depthf), _stream(stream)), "vh_integrate_depthf")

    def fuse_frame(self, pose, verts, normals=None, depthf=None, stream=None):
        """SDF_Hashtable::integrate without its host syncs: set pose, alloc, compact, integrate."""
        self.set_pose(pose, stream)
        L.check(self.lib.vh_fuse_frame(self._h, _ptr(verts), _ptr(normals), _ptr(depthf), _stream(stream)), "vh_fuse_frame")

    def stats(self, stream=None) -> L.VhStats:
        s = L.VhStats()
        L.check(self.lib.vh_get_stats(self._h, C.byref(s), _stream(stream)), "vh_get_stats")
        return s

    def garbage_collect(self, scope=L.VH_GC_VISIBLE, sdf_threshold=0.0, weight_decay=0.0, stream=None):
        """Starve + release blocks (Niessner 2013, 4.4); empties the visible list, so compact again before integrating."""
        L.check(self.lib.vh_garbage_collect(self._h, int(scope), float(sdf_threshold), float(weight_decay), _stream(stream)),
                "vh_garbage_collect")

    def stream_out(self, center, radius, capacity, pinned_host=True, stream=None):
        """Move the blocks farther than `radius` metres from `center` out of the table (Niessner 2013, 4.5).
        -> (entries [n,5] int32 tensor, voxels [n,512,2] float32 tensor); in pinned host memory by default, which the
        kernel writes straight over the host link."""
        import torch

        cap = max(int(capacity), 1)
        kw = dict(pin_memory=True) if pinned_host else dict(device="cuda")
        ent = torch.zeros((cap, 5), dtype=torch.int32, **kw)
        vox = torch.zeros((cap, 512, 2), dtype=torch.float32, **kw)
        c3 = (C.c_float * 3)(*[float(x) for x in center])
        n = C.c_int(0)
        L.check(self.lib.vh_stream_out(self._h, c3, float(radius), _ptr(ent), _ptr(vox), int(capacity), C.byref(n), _stream(stream)),
                "vh_stream_out")
        return ent[: n.value], vox[: n.value]

    def stream_in(self, entries, voxels, stream=None) -> int:
        """Bring blocks back (device or pinned-host tensors as returned by stream_out); -> number accepted."""
        n = C.c_int(0)
        L.check(self.lib.vh_stream_in(self._h, _ptr(entries), _ptr(voxels), int(entries.shape[0]), C.byref(n), _stream(stream)),
                "vh_stream_in")
        return n.value

    # -- tracking -----------------------------------------------------------------------------------
    def icp_reset(self, reset_estimate=True, stream=None):
        L.check(self.lib.vh_icp_reset(self._h, int(reset_estimate), _stream(stream)))

    def icp_iterate(self, inp, inpN, tgt, tgtN, stream=None):
        L.check(self.lib.vh_icp_iterate(self._h, _ptr(inp), _ptr(inpN), _ptr(tgt), _ptr(tgtN), _stream(stream)))

    def icp_align(self, inp, inpN, tgt, tgtN, iterations=0, stream=None):
        L.check(self.lib.vh_icp_align(self._h, _ptr(inp), _ptr(inpN), _ptr(tgt), _ptr(tgtN), iterations, _stream(stream)))

    def icp_reduce(self, inp, inpN, tgt, tgtN, row0, row1, d_system, stream=None):
        L.check(self.lib.vh_icp_reduce(self._h, _ptr(inp), _ptr(inpN), _ptr(tgt), _ptr(tgtN), row0, row1, _ptr(d_system), _stream(stream)))

    def set_peers(self, rank: int, world: int, peer_ptrs):
        """Register the peer-mapped exchange regions for the fused cross-GPU all-reduce of the ICP tail."""
        arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        L.check(self.lib.vh_set_peers(self._h, rank, world, arr), "vh_set_peers")

    def peer_bytes(self) -> int:
        return int(self.lib.vh_peer_bytes())

    def track_frame(self, depth_u16, verts, normals, depthf, tgt, tgtN, iterations=0, pose_in=None, pose_out=None, stream=None):
        """Pre-processing + Align (+ pose chain) of one frame in one persistent launch; verts / normals / depthf are outputs."""
        L.check(self.lib.vh_track_frame(self._h, _ptr(depth_u16), _ptr(verts), _ptr(normals), _ptr(depthf), _ptr(tgt), _ptr(tgtN),
                                        iterations, _ptr(pose_in), _ptr(pose_out), _stream(stream)), "vh_track_frame")

    def icp_align_rows(self, inp, inpN, tgt, tgtN, row0, row1, iterations=0, stream=None):
        L.check(self.lib.vh_icp_align_rows(self._h, _ptr(inp), _ptr(inpN), _ptr(tgt), _ptr(tgtN), row0, row1, iterations,
                                           _stream(stream)), "vh_icp_align_rows")

    def icp_solve(self, d_system, stream=None):
        L.check(self.lib.vh_icp_solve(self._h, _ptr(d_system), _stream(stream)))

    def icp_get(self, stream=None):
        """(delta 4x4 row-major, twist (v, omega), last system as float32[32])."""
        delta = np.zeros(16, np.float32)
        twist = np.zeros(6, np.float32)
        sysv = np.zeros(32, np.float32)
        L.check(self.lib.vh_icp_get(self._h, delta.ctypes.data, twist.ctypes.data, sysv.ctypes.data, _stream(stream)))
        return delta.reshape(4, 4), twist, sysv

    def icp_set_twist(self, twist6, stream=None):
        t = np.ascontiguousarray(twist6, dtype=np.float32)
        L.check(self.lib.vh_icp_set_delta(self._h, t.ctypes.data, _stream(stream)))

    def icp_delta_device(self) -> int:
        return int(self.lib.vh_icp_delta_device(self._h))

    def pose_compose(self, d_pose_in, d_pose_out, stream=None):
        L.check(self.lib.vh_pose_compose(self._h, _ptr(d_pose_in), _ptr(d_pose_out), _stream(stream)))

    def find_correspondences(self, inp, inpN, tgt, tgtN, delta, corr, corrN, res, d_err, stream=None):
        d = _pose_arg(delta)
        L.check(self.lib.vh_find_correspondences(self._h, _ptr(inp), _ptr(inpN), _ptr(tgt), _ptr(tgtN), d.ctypes.data,
                                                 _ptr(corr), _ptr(corrN), _ptr(res), _ptr(d_err), _stream(stream)))

    def jacobians(self, corr, corrN, J, stream=None):
        L.check(self.lib.vh_jacobians(self._h, _ptr(corr), _ptr(corrN), _ptr(J), _stream(stream)))

    def icp_reduce_corr(self, corr, corrN, res, d_system, stream=None):
        L.check(self.lib.vh_icp_reduce_corr(self._h, _ptr(corr), _ptr(corrN), _ptr(res), _ptr(d_system), _stream(stream)))

    def raycast(self, verts, normals, stream=None):
        L.check(self.lib.vh_raycast(self._h, _ptr(verts), _ptr(normals), _stream(stream)), "vh_raycast")

    # -- export -------------------------------------------------------------------------------------
    def export_entries(self) -> np.ndarray:
        n = C.c_int(0)
        L.check(self.lib.vh_export_entries(self._h, 0, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=L.VOXEL_ENTRY_DTYPE)
        L.check(self.lib.vh_export_entries(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out[: n.value]

    def export_compact(self) -> np.ndarray:
        n = C.c_int(0)
        L.check(self.lib.vh_export_compact(self._h, 0, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=L.VOXEL_ENTRY_DTYPE)
        L.check(self.lib.vh_export_compact(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out[: n.value]

    def export_block(self, ptr: int) -> np.ndarray:
        out = np.zeros(512, dtype=L.VOXEL_DTYPE)
        L.check(self.lib.vh_export_block(self._h, int(ptr), out.ctypes.data))
        return out

    def block_dict(self) -> dict:
        """{(x, y, z): float32[512, 2]} for every allocated block (test helper; slow)."""
        return {(int(e["x"]), int(e["y"]), int(e["z"])): self.export_block(int(e["ptr"])).view(np.float32).reshape(512, 2)
                for e in self.export_entries() if e["ptr"] >= 0}

    def extract_mesh(self, stream=None):
        """Zero level set as a triangle soup: float32 CUDA tensor [n, 3, 3] (world metres, normals towards free space)."""
        import torch

        n = C.c_int(0)
        L.check(self.lib.vh_extract_mesh(self._h, None, 0, C.byref(n), _stream(stream)), "vh_extract_mesh")
        tris = torch.empty((max(n.value, 1), 3, 3), dtype=torch.float32, device="cuda")
        L.check(self.lib.vh_extract_mesh(self._h, _ptr(tris), n.value, C.byref(n), _stream(stream)), "vh_extract_mesh")
        return tris[: n.value]

    def save_mesh_ply(self, path, tris=None):
        t = (self.extract_mesh() if tris is None else tris).detach().cpu().contiguous().numpy()
        L.check(self.lib.vh_save_mesh_ply(str(path).encode(), t.ctypes.data, int(t.shape[0])), "vh_save_mesh_ply")
        return int(t.shape[0])

    def save(self, path: str):
        L.check(self.lib.vh_save(self._h, str(path).encode()), "vh_save")

    def load(self, path: str):
        L.check(self.lib.vh_load(self._h, str(path).encode()), "vh_load")

    def dump_text(self, path: str):
        L.check(self.lib.vh_dump_text(self._h, str(path).encode()), "vh_dump_text")


class SDF_Hashtable:
    """Reference-facing fusion class (SDF_Hashtable.h:24-42), headless."""

    def __init__(self, cfg: Config | None = None):
        self.cfg = cfg or Config()
        self.ctx = Context(self.cfg)

    def integrate(self, viewMat, verts, normals, sync: bool = True):
        """ref SDF_Hashtable.cpp:11-40: viewMat is camera->world, row-major 4x4."""
        self.ctx.fuse_frame(viewMat, verts, normals)
        if sync:
            _torch().cuda.current_stream().synchronize()

    def registerGLtoCUDA(self, renderer=None):  # ref :42-50 -- nothing to register headless
        return None

    def unmapCUDApointers(self):  # ref :52-58
        return None

    def occupiedBlockCount(self) -> int:
        return self.ctx.stats().numVisible


class CameraTracking:
    """Reference-facing tracking class (CameraTracking.h:34-59)."""

    def __init__(self, width: int, height: int, cfg: Config | None = None, ctx: Context | None = None):
        if ctx is None:
            cfg = cfg or Config(width=width, height=height, numBuckets=16, numVoxelBlocks=16)
            ctx = Context(cfg)
        self.ctx = ctx
        self.maxIters = ctx.cfg.icpIterations  # ref CameraTracking.h:40

    def Align(self, d_input, d_inputNormals, d_target, d_targetNormals, d_depthInput=None, d_depthTarget=None):
        """ref CameraTracking.cpp:26-69: maxIters Gauss-Newton iterations; the estimate accumulates (Q24)."""
        self.ctx.icp_align(d_input, d_inputNormals, d_target, d_targetNormals, self.maxIters)

    def getTransform(self) -> np.ndarray:
        """4x4, input frame -> target frame.  The reference returns an Eigen column-major matrix; this
        is the same matrix as a row-major numpy array (M[r, c])."""
        return self.ctx.icp_get()[0]


class FramePipeline:
    """Native frame loop (vh_pipeline_*): preprocess -> ICP -> pose chain -> alloc -> compact -> integrate."""

    FRAME_TO_FRAME, FRAME_TO_MODEL, NONE = 0, 1, 2

    def __init__(self, ctx: Context, iterations: int = 0, mode: int = 0, use_graph: bool = True, overlap: bool = False):
        """overlap=True (frame-to-frame + graphs only): the fusion of frame k runs beside the tracking of frame k+1;
        the model then lags the pose until flush() / pose()."""
        self.ctx = ctx
        self.lib = ctx.lib
        self._p = C.c_void_p()
        flags = (L.VH_PIPE_GRAPH if use_graph else 0) | (L.VH_PIPE_OVERLAP if overlap else 0)
        L.check(self.lib.vh_pipeline_create(ctx.handle, iterations, mode, flags, C.byref(self._p)), "vh_pipeline_create")

    def flush(self, stream=None):
        L.check(self.lib.vh_pipeline_flush(self._p, _stream(stream)), "vh_pipeline_flush")

    def close(self):
        if getattr(self, "_p", None) is not None and self._p:
            self.lib.vh_pipeline_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, pose=None, stream=None):
        p = None if pose is None else _pose_arg(pose)
        L.check(self.lib.vh_pipeline_reset(self._p, 0 if p is None else p.ctypes.data, _stream(stream)), "vh_pipeline_reset")

    def push_device(self, d_depth, stream=None):
        L.check(self.lib.vh_pipeline_push_device(self._p, _ptr(d_depth), _stream(stream)), "vh_pipeline_push_device")

    def push_device_ready(self, d_depth, ready_event=None, stream=None):
        """One frame whose depth image is not produced on `stream`: complete once `ready_event` (a torch.cuda.Event that has
        been recorded, or None = complete already) has fired.  Overlapped schedule: its pre-processing runs beside the
        tracking of the previous frame."""
        ev = 0 if ready_event is None else (int(ready_event) if isinstance(ready_event, int) else int(ready_event.cuda_event))
        L.check(self.lib.vh_pipeline_push_device_ready(self._p, _ptr(d_depth), ev, _stream(stream)), "vh_pipeline_push_device_ready")

    def push_host(self, h_depth, h_pose_out=None, stream=None):
        L.check(self.lib.vh_pipeline_push_host(self._p, _ptr(h_depth), _ptr(h_pose_out), _stream(stream)), "vh_pipeline_push_host")

    def pose(self, stream=None) -> np.ndarray:
        out = np.zeros(16, np.float32)
        L.check(self.lib.vh_pipeline_pose(self._p, out.ctypes.data, _stream(stream)), "vh_pipeline_pose")
        return out.reshape(4, 4)

    def pose_async(self, h_pose_pinned, stream=None):
        """Stream-ordered copy of the latest pose into a pinned host tensor (16 floats); no synchronisation."""
        L.check(self.lib.vh_pipeline_pose_async(self._p, _ptr(h_pose_pinned), _stream(stream)), "vh_pipeline_pose_async")

    def depthf_ptr(self) -> int:
        """Device address of the dense metric depth (W x H floats) of the latest pushed frame."""
        p = C.c_void_p()
        L.check(self.lib.vh_pipeline_depthf(self._p, C.byref(p)), "vh_pipeline_depthf")
        return int(p.value)

    def launches(self) -> int:
        return int(self.lib.vh_pipeline_launches(self._p))
