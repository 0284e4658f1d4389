"""GPU-box: timeline of one k_icp_iter launch from %globaltimer stamps (library built with -DVH_ICP_TRACE)."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["VH_EXTRA_NVCC_FLAGS"] = "-DVH_ICP_TRACE" + (" -DVH_ICP_TRACE_TWICE" if "twice" in sys.argv else "")
import torch  # noqa: E402

from voxelhashing_demo_b200 import _build  # noqa: E402

_build.build(force=True)
import bench  # noqa: E402
from voxelhashing_demo_b200 import Context  # noqa: E402
from voxelhashing_demo_b200 import lib as L  # noqa: E402

cfg, scene, traj, _ = bench.workload_config("C3" if "c3" in sys.argv else "C2")
ctx = Context(cfg)
frames, poses = bench.render_frames(cfg, scene, traj, 2)
d = torch.from_numpy(frames).cuda()
a, b = ctx.new_maps(), ctx.new_maps()
ctx.preprocess(d[0], *a)
ctx.preprocess(d[1], *b)
s = torch.cuda.Stream()
lib = L.load_library()
with torch.cuda.stream(s):
    ctx.icp_reset(True, s)
    for _ in range(10):
        ctx.icp_iterate(b[0], b[1], a[0], a[1], s)
    s.synchronize()
    tr = np.zeros(1024 * 8, np.uint64)
    lib.vh_icp_trace_read.argtypes = [C.c_void_p, C.c_int]
    lib.vh_icp_trace_read(tr.ctypes.data, tr.size)
print('solve cycles: gauss-jordan', tr[1023*8+0], ' exp', tr[1023*8+1], ' product+orthonormalise', tr[1023*8+2])
tr = tr.reshape(1024, 8).astype(np.int64)[:148]
tr = tr[tr[:, 0] > 0]
print('CTAs', len(tr))
t0 = tr[:, 0].min()
names = ["start", "after delta load", "after main loop", "after block reduce", "after ticket", "tail: partials summed", "tail: system ready", "tail: solved"]
for k in range(5):
    col = tr[:, k] - t0
    print(f"{names[k]:26s} min {col.min():6d} ns  median {int(np.median(col)):6d}  max {col.max():6d}")
last = np.argmax(tr[:, 4])
for k in (5, 6, 7) if "twice" not in sys.argv else (6, 7, 5):
    label = names[k] if not ("twice" in sys.argv and k == 5) else "tail: solved a 2nd time"
    print(f"{label:26s} {tr[last, k] - t0:6d} ns   (last CTA = {last})")
