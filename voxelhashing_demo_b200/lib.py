"""ctypes binding of libvh_b200.so (include/vh/abi.h).

The shared library is the product; this module only declares its entry points.  There is no
Python/numpy/torch fallback for any of them: if the library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libvh_b200.so"

VH_OK, VH_ERR_INVALID, VH_ERR_CUDA, VH_ERR_NO_DEVICE, VH_ERR_CAPACITY, VH_ERR_UNSUPPORTED = range(6)
VH_GC_VISIBLE, VH_GC_ALL = 0, 1
VH_PIPE_GRAPH, VH_PIPE_OVERLAP = 1, 2
POLICY_REF_EXACT, POLICY_FIXED = 0, 1


class VHError(RuntimeError):
    pass


class Float4x4(C.Structure):
    _fields_ = [("entries", C.c_float * 16)]


class HashTableParams(C.Structure):  # ref VoxelDataStructures.h:28-52, 176 bytes
    _fields_ = [
        ("global_transform", Float4x4),
        ("inv_global_transform", Float4x4),
        ("numBuckets", C.c_uint),
        ("bucketSize", C.c_uint),
        ("attachedLinkedListSize", C.c_uint),
        ("numVoxelBlocks", C.c_uint),
        ("voxelBlockSize", C.c_int),
        ("voxelSize", C.c_float),
        ("numOccupiedBlocks", C.c_uint),
        ("maxIntegrationDistance", C.c_float),
        ("truncScale", C.c_float),
        ("truncation", C.c_float),
        ("integrationWeightSample", C.c_uint),
        ("integrationWeightMax", C.c_float),
    ]


class VhConfig(C.Structure):
    _fields_ = [
        ("table", HashTableParams),
        ("policy", C.c_int),
        ("width", C.c_int),
        ("height", C.c_int),
        ("fx", C.c_float),
        ("fy", C.c_float),
        ("cx", C.c_float),
        ("cy", C.c_float),
        ("depthScale", C.c_float),
        ("depthMin", C.c_float),
        ("depthMax", C.c_float),
        ("overflowSlots", C.c_uint),
        ("icpDistThres", C.c_float),
        ("icpNormalThres", C.c_float),
        ("icpIterations", C.c_int),
        ("partCount", C.c_int),
        ("partRank", C.c_int),
        ("bilateralSigmaSpace", C.c_float),
        ("bilateralSigmaRange", C.c_float),
    ]


class VhStats(C.Structure):
    _fields_ = [
        ("heapCounter", C.c_int),
        ("numAllocated", C.c_int),
        ("numVisible", C.c_int),
        ("overflowUsed", C.c_int),
        ("dropped", C.c_int),
        ("numUpdated", C.c_ulonglong),
        ("lastInserted", C.c_int),
        ("lastFreed", C.c_int),
        ("overflowLeaked", C.c_int),
        ("exchangeTimeouts", C.c_int),
    ]


class VhIcpSystem(C.Structure):
    _fields_ = [("JtJ", C.c_float * 21), ("Jtr", C.c_float * 6), ("error", C.c_float), ("count", C.c_float), ("pad", C.c_float * 3)]


VOXEL_ENTRY_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("z", "<i4"), ("ptr", "<i4"), ("offset", "<i4")])
VOXEL_DTYPE = np.dtype([("sdf", "<f4"), ("weight", "<f4")])
assert C.sizeof(HashTableParams) == 176 and VOXEL_ENTRY_DTYPE.itemsize == 20 and VOXEL_DTYPE.itemsize == 8

# every symbol include/vh/abi.h declares (tests check the library exports all of them)
LEGACY_SYMBOLS = [
    "updateConstantHashTableParams", "deviceAllocate", "deviceFree", "resetHashTableMutexes", "allocBlocks",
    "flattenIntoBuffer", "calculateKinectProjectionMatrix", "integrateDepthMap", "mapGLobjectsToCUDApointers",
    "preProcess", "SetCameraIntrinsic", "computeCorrespondences", "CalculateJacobiansAndResiduals",
    "buildLinearSystemOnDevice", "vhLegacyCompactTable", "vhLegacyVoxelBlocks", "vhLegacyCompactCounter",
    "vhLegacyContext",
]
HANDLE_SYMBOLS = [
    "vh_last_error", "vh_default_config", "vh_device_count", "vh_create", "vh_destroy", "vh_reset", "vh_get_config",
    "vh_set_intrinsics", "vh_set_intrinsic_matrices", "vh_set_tuning", "vh_bytes_allocated", "vh_preprocess", "vh_set_pose",
    "vh_set_pose_device", "vh_alloc_blocks", "vh_alloc_blocks_depth", "vh_compact", "vh_integrate", "vh_integrate_depthf", "vh_fuse_frame",
    "vh_get_stats", "vh_garbage_collect", "vh_stream_out", "vh_stream_in", "vh_icp_reset", "vh_icp_iterate", "vh_icp_align", "vh_track_frame", "vh_icp_reduce", "vh_icp_solve", "vh_icp_get",
    "vh_icp_set_delta", "vh_set_peers", "vh_peer_bytes", "vh_icp_align_rows", "vh_icp_delta_device", "vh_pose_compose", "vh_icp_reduce_corr", "vh_find_correspondences",
    "vh_jacobians", "vh_raycast", "vh_export_entries", "vh_export_compact", "vh_export_block",
    "vh_compact_table_device", "vh_compact_counter_device", "vh_voxel_blocks_device", "vh_save", "vh_load",
    "vh_dump_text", "vh_extract_mesh", "vh_save_mesh_ply", "vh_depth_read", "vh_depth_free", "vh_depth_write_png", "vh_depth_last_error",
    "vh_pipeline_create", "vh_pipeline_flush", "vh_pipeline_destroy", "vh_pipeline_reset", "vh_pipeline_push_device", "vh_pipeline_push_device_ready",
    "vh_pipeline_push_host", "vh_pipeline_pose", "vh_pipeline_pose_device", "vh_pipeline_pose_async", "vh_pipeline_depthf", "vh_pipeline_maps", "vh_pipeline_launches",
    "vh_dist_create", "vh_dist_handle_bytes", "vh_dist_export", "vh_dist_connect", "vh_dist_broadcast_frame", "vh_dist_frame_consumed", "vh_dist_timeouts", "vh_dist_destroy",
]

_lib = None


def load_library() -> C.CDLL:
    """Load libvh_b200.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise VHError(
            f"{LIB_PATH} is missing: build it with `python -m voxelhashing_demo_b200._build` "
            "(nvcc, sm_100a). This package has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_LOCAL)
    P, I, F = C.c_void_p, C.c_int, C.c_float
    lib.vh_last_error.restype = C.c_char_p
    lib.vh_default_config.argtypes = [C.POINTER(VhConfig)]
    lib.vh_default_config.restype = None
    lib.vh_device_count.restype = I
    lib.vh_create.argtypes = [C.POINTER(VhConfig), C.POINTER(P)]
    lib.vh_destroy.argtypes = [P]
    lib.vh_destroy.restype = None
    lib.vh_reset.argtypes = [P, P]
    lib.vh_get_config.argtypes = [P, C.POINTER(VhConfig)]
    lib.vh_set_intrinsics.argtypes = [P, F, F, F, F]
    lib.vh_set_intrinsic_matrices.argtypes = [P, P, P]
    lib.vh_bytes_allocated.argtypes = [P]
    lib.vh_set_tuning.argtypes = [P, I, I]
    lib.vh_bytes_allocated.restype = C.c_ulonglong
    lib.vh_preprocess.argtypes = [P, P, P, P, P, P]
    lib.vh_set_pose.argtypes = [P, P, P]
    lib.vh_set_pose_device.argtypes = [P, P, P]
    lib.vh_alloc_blocks.argtypes = [P, P, P, P]
    lib.vh_alloc_blocks_depth.argtypes = [P, P, P]
    lib.vh_compact.argtypes = [P, P]
    lib.vh_integrate.argtypes = [P, P, P]
    lib.vh_integrate_depthf.argtypes = [P, P, P]
    lib.vh_fuse_frame.argtypes = [P, P, P, P, P]
    lib.vh_get_stats.argtypes = [P, C.POINTER(VhStats), P]
    lib.vh_garbage_collect.argtypes = [P, C.c_int, C.c_float, C.c_float, P]
    lib.vh_stream_out.argtypes = [P, C.POINTER(C.c_float), C.c_float, P, P, C.c_int, C.POINTER(C.c_int), P]
    lib.vh_stream_in.argtypes = [P, P, P, C.c_int, C.POINTER(C.c_int), P]
    lib.vh_icp_reset.argtypes = [P, I, P]
    lib.vh_icp_iterate.argtypes = [P, P, P, P, P, P]
    lib.vh_icp_align.argtypes = [P, P, P, P, P, I, P]
    lib.vh_track_frame.argtypes = [P, P, P, P, P, P, P, I, P, P, P]
    lib.vh_icp_reduce.argtypes = [P, P, P, P, P, I, I, P, P]
    lib.vh_icp_solve.argtypes = [P, P, P]
    lib.vh_set_peers.argtypes = [P, I, I, C.POINTER(C.c_void_p)]
    lib.vh_peer_bytes.restype = C.c_ulonglong
    lib.vh_icp_align_rows.argtypes = [P, P, P, P, P, I, I, I, P]
    lib.vh_icp_get.argtypes = [P, P, P, P, P]
    lib.vh_icp_set_delta.argtypes = [P, P, P]
    lib.vh_icp_delta_device.argtypes = [P]
    lib.vh_icp_delta_device.restype = P
    lib.vh_pose_compose.argtypes = [P, P, P, P]
    lib.vh_icp_reduce_corr.argtypes = [P, P, P, P, P, P]
    lib.vh_find_correspondences.argtypes = [P, P, P, P, P, P, P, P, P, P, P]
    lib.vh_jacobians.argtypes = [P, P, P, P, P]
    lib.vh_raycast.argtypes = [P, P, P, P]
    lib.vh_export_entries.argtypes = [P, P, I, C.POINTER(I)]
    lib.vh_export_compact.argtypes = [P, P, I, C.POINTER(I)]
    lib.vh_export_block.argtypes = [P, I, P]
    for n in ("vh_compact_table_device", "vh_compact_counter_device", "vh_voxel_blocks_device"):
        getattr(lib, n).argtypes = [P]
        getattr(lib, n).restype = P
    lib.vh_save.argtypes = [P, C.c_char_p]
    lib.vh_load.argtypes = [P, C.c_char_p]
    lib.vh_dump_text.argtypes = [P, C.c_char_p]
    lib.vh_extract_mesh.argtypes = [P, P, C.c_int, C.POINTER(C.c_int), P]
    lib.vh_save_mesh_ply.argtypes = [C.c_char_p, P, C.c_int]
    lib.vh_depth_read.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_uint16)), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.vh_depth_free.argtypes = [C.POINTER(C.c_uint16)]
    lib.vh_depth_free.restype = None
    lib.vh_depth_write_png.argtypes = [C.c_char_p, P, C.c_int, C.c_int, C.c_int]
    lib.vh_depth_last_error.restype = C.c_char_p
    # native frame pipeline
    lib.vh_pipeline_create.argtypes = [P, I, I, I, C.POINTER(P)]
    lib.vh_pipeline_flush.argtypes = [P, P]
    lib.vh_pipeline_destroy.argtypes = [P]
    lib.vh_pipeline_destroy.restype = None
    lib.vh_pipeline_reset.argtypes = [P, P, P]
    lib.vh_pipeline_push_device.argtypes = [P, P, P]
    lib.vh_pipeline_push_device_ready.argtypes = [P, P, P, P]
    lib.vh_dist_create.argtypes = [P, I, I, P]
    lib.vh_dist_handle_bytes.restype = C.c_ulonglong
    lib.vh_dist_export.argtypes = [P, P]
    lib.vh_dist_connect.argtypes = [P, P]
    lib.vh_dist_broadcast_frame.argtypes = [P, P, P, P, P]
    lib.vh_dist_frame_consumed.argtypes = [P, P]
    lib.vh_dist_timeouts.argtypes = [P]
    lib.vh_dist_destroy.argtypes = [P]
    lib.vh_dist_destroy.restype = None
    lib.vh_pipeline_push_host.argtypes = [P, P, P, P]
    lib.vh_pipeline_pose.argtypes = [P, P, P]
    lib.vh_pipeline_pose_async.argtypes = [P, P, P]
    lib.vh_pipeline_depthf.argtypes = [P, C.POINTER(P)]
    lib.vh_pipeline_pose_device.argtypes = [P]
    lib.vh_pipeline_pose_device.restype = P
    lib.vh_pipeline_maps.argtypes = [P, I, C.POINTER(P), C.POINTER(P)]
    lib.vh_pipeline_launches.argtypes = [P]
    lib.vh_pipeline_launches.restype = C.c_longlong
    # legacy (reference names)
    HP = C.POINTER(HashTableParams)
    lib.updateConstantHashTableParams.argtypes = [HP]
    lib.updateConstantHashTableParams.restype = None
    lib.deviceAllocate.argtypes = [HP]
    lib.deviceAllocate.restype = None
    lib.deviceFree.restype = None
    lib.resetHashTableMutexes.argtypes = [HP]
    lib.resetHashTableMutexes.restype = None
    lib.allocBlocks.argtypes = [P, P]
    lib.allocBlocks.restype = None
    lib.flattenIntoBuffer.argtypes = [HP]
    lib.flattenIntoBuffer.restype = I
    lib.calculateKinectProjectionMatrix.restype = None
    lib.integrateDepthMap.argtypes = [HP, P]
    lib.integrateDepthMap.restype = None
    lib.mapGLobjectsToCUDApointers.argtypes = [P, P, P]
    lib.mapGLobjectsToCUDApointers.restype = None
    lib.preProcess.argtypes = [P, P, P]
    lib.preProcess.restype = None
    lib.SetCameraIntrinsic.argtypes = [P, P]
    lib.SetCameraIntrinsic.restype = C.c_bool
    lib.computeCorrespondences.argtypes = [P, P, P, P, P, P, C.POINTER(Float4x4), I, I]   # by-value float4x4 = hidden reference
    lib.computeCorrespondences.restype = F
    lib.CalculateJacobiansAndResiduals.argtypes = [P, P, P, P]
    lib.CalculateJacobiansAndResiduals.restype = None
    lib.buildLinearSystemOnDevice.argtypes = [P, P, P, P, P]
    lib.buildLinearSystemOnDevice.restype = None
    for n in ("vhLegacyCompactTable", "vhLegacyVoxelBlocks", "vhLegacyCompactCounter", "vhLegacyContext"):
        getattr(lib, n).restype = P
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != VH_OK:
        msg = load_library().vh_last_error().decode(errors="replace")
        raise VHError(f"{what or 'libvh_b200'} failed (status {rc}): {msg}")
