"""GPU-box: timeline of the persistent Align kernel (k_track.cu) from %globaltimer stamps (library built with
-DVH_ICP_TRACE).  usage: python tools/align_trace.py [c3] [ctas=N]"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["VH_EXTRA_NVCC_FLAGS"] = "-DVH_ICP_TRACE"
for a in sys.argv[1:]:
    if a.startswith("ctas="):
        os.environ["VH_ICP_CTAS"] = a.split("=")[1]
    if a.startswith("ablate="):
        os.environ["VH_EXTRA_NVCC_FLAGS"] += " -DVH_ALIGN_ABLATE=" + a.split("=")[1]
    if a.startswith("batch="):
        os.environ["VH_EXTRA_NVCC_FLAGS"] += " -DVH_ALIGN_BATCH=" + a.split("=")[1]
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
IT, SL = 24, 8
with torch.cuda.stream(s):
    for _ in range(3):
        ctx.icp_reset(True, s)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        if "pre" in sys.argv:
            ctx.track_frame(d[1], b[0], b[1], b[2], a[0], a[1], 20, None, None, s)
        else:
            ctx.icp_align(b[0], b[1], a[0], a[1], 20, s)
        e1.record(s)
        s.synchronize()
        print("align x20 by events: %.1f us" % (e0.elapsed_time(e1) * 1e3))
    tr = np.zeros(1024 * IT * SL, np.uint64)
    lib.vh_align_trace_read.argtypes = [C.c_void_p, C.c_int]
    lib.vh_align_trace_read(tr.ctypes.data, tr.size)
tr = tr.reshape(1024, IT, SL).astype(np.int64)
if "pre" in sys.argv:
    pr = tr[: int((tr[:, 23, 0] > 0).sum()), 23]
    t00 = pr[:, 0].min()
    print("prologue: start spread %d ns; maps written: median %d max %d ns; after fence + barrier + fence: median %d max %d ns" % (
        pr[:, 0].max() - t00, np.median(pr[:, 1] - t00), (pr[:, 1] - t00).max(), np.median(pr[:, 2] - t00), (pr[:, 2] - t00).max()))
n = int((tr[:, 0, 0] > 0).sum())
tr = tr[:n, :20]
print("CTAs", n)
t0 = tr[:, 0, 0].min()
names = ["iteration start", "main loop done", "block reduce done", "exchange read + barrier", "solve done"]
print("first iteration starts: min 0, median %d, max %d ns after the earliest CTA" % (np.median(tr[:, 0, 0] - t0), (tr[:, 0, 0] - t0).max()))
print("whole kernel (earliest start -> latest solve of iteration 19): %.2f us" % ((tr[:, 19, 4].max() - t0) / 1e3))
for it in (0, 1, 2, 10, 19):
    base = tr[:, it, 0].min()
    print(f"iteration {it}:")
    for k in range(5):
        col = tr[:, it, k] - base
        print(f"   {names[k]:26s} min {col.min():6d} ns  median {int(np.median(col)):6d}  max {col.max():6d}")
per = np.diff(tr[:, :, 0].min(axis=0))
print("iteration period (earliest start to earliest start), ns:", per.tolist())
dur = tr[:, 1:, :]
print("median over CTAs and iterations 1..19 of each phase (ns): main %d, block reduce %d, exchange %d, solve %d" % (
    np.median(dur[:, :, 1] - dur[:, :, 0]), np.median(dur[:, :, 2] - dur[:, :, 1]), np.median(dur[:, :, 3] - dur[:, :, 2]),
    np.median(dur[:, :, 4] - dur[:, :, 3])))
print("max over CTAs, median over iterations 1..19 (ns): main %d, block reduce %d, exchange %d, solve %d" % (
    np.median((dur[:, :, 1] - dur[:, :, 0]).max(axis=0)), np.median((dur[:, :, 2] - dur[:, :, 1]).max(axis=0)),
    np.median((dur[:, :, 3] - dur[:, :, 2]).max(axis=0)), np.median((dur[:, :, 4] - dur[:, :, 3]).max(axis=0))))
print("warp 0 after the exchange (median, ns): final sum %d, solve core %d, publish + barrier %d" % (
    np.median(dur[:, :, 5] - dur[:, :, 3]), np.median(dur[:, :, 6] - dur[:, :, 5]), np.median(dur[:, :, 4] - dur[:, :, 6])))
