"""Shared fixtures.  `-m "not gpu"` runs here (no GPU); `-m gpu` runs on the B200 box through the C ABI."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_library():
    """libvh_b200.so, (re)built in-tree if stale.  nvcc cross-compiles without a GPU."""
    from voxelhashing_demo_b200 import _build

    return _build.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding

    binding.build_oracle()
    return binding


def small_cfg(**kw):
    """160x120 camera with the reference intrinsics scaled by 1/4: oracle runs in milliseconds."""
    from voxelhashing_demo_b200 import Config

    base = dict(width=160, height=120, fx=517.3 / 4, fy=516.5 / 4, cx=318.6 / 4, cy=255.3 / 4)
    base.update(kw)
    return Config(**base)


def render(cfg, scene, pose):
    from voxelhashing_demo_b200 import scenes

    return scenes.render_depth(scene, pose, cfg.width, cfg.height, cfg.fx, cfg.fy, cfg.cx, cfg.cy, cfg.depthScale)


def entries_to_set(entries) -> set:
    """{(x, y, z)} from a structured VoxelEntry array or an (n,5) int array."""
    if getattr(entries, "dtype", None) is not None and entries.dtype.names:
        return {(int(e["x"]), int(e["y"]), int(e["z"])) for e in entries}
    return {(int(e[0]), int(e[1]), int(e[2])) for e in entries}


def rot_err(Ra, Rb) -> float:
    """Frobenius norm of Ra^T Rb - I (the pose metric of SURVEY.md section 8c)."""
    return float(np.linalg.norm(np.asarray(Ra, np.float64).T @ np.asarray(Rb, np.float64) - np.eye(3)))
