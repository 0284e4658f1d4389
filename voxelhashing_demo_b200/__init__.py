"""voxelhashing_demo_b200 -- B200-native fusion-and-tracking hot path of nilspin/VoxelHashing_demo.

The product is the CUDA shared library `libvh_b200.so` (C ABI in include/vh/abi.h) plus the C++
host classes in include/; this package is its Python face: a ctypes binding (`lib`), the mirror of
the reference's host classes (`fusion`), synthetic scenes (`scenes`) and the multi-GPU plumbing
(`dist`).  Importing the package does not load the library; constructing a `Context` does, and
raises if the library has not been built -- there is no CPU or PyTorch fallback.
"""
from .lib import POLICY_FIXED, POLICY_REF_EXACT, VHError, load_library  # noqa: F401
from .fusion import CameraTracking, Config, Context, FramePipeline, SDF_Hashtable, read_depth, write_depth_png  # noqa: F401

__all__ = ["Config", "Context", "SDF_Hashtable", "CameraTracking", "FramePipeline", "VHError", "load_library", "read_depth", "write_depth_png",
           "POLICY_FIXED", "POLICY_REF_EXACT"]
