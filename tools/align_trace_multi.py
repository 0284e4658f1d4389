"""GPU-box, under torchrun: timeline of the persistent Align kernel with the cross-GPU exchange inside (library built with
-DVH_ICP_TRACE).  python -m torch.distributed.run --nproc-per-node N tools/align_trace_multi.py [c2]"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["VH_EXTRA_NVCC_FLAGS"] = "-DVH_ICP_TRACE" + "".join(" -D" + a for a in sys.argv[1:] if a.startswith("VH_"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from voxelhashing_demo_b200 import _build  # noqa: E402

if rank == 0:
    _build.build(force=True)
dist.barrier()
import bench  # noqa: E402
from voxelhashing_demo_b200 import Context  # noqa: E402
from voxelhashing_demo_b200 import lib as L  # noqa: E402
from voxelhashing_demo_b200.dist import PartitionedTracker, row_range  # noqa: E402

name = "C2" if "c2" in sys.argv else "C4"
cfg, scene, traj, _ = bench.workload_config(name, world, rank)
cfg.numVoxelBlocks = 4096
ctx = Context(cfg)
tr = PartitionedTracker(ctx, rank, world, overlap=False)          # sets the peer mailboxes up
assert tr.fused or world == 1
frames, poses = bench.render_frames(cfg, scene, traj, 2)
d = torch.from_numpy(frames).cuda()
a, b = ctx.new_maps(), ctx.new_maps()
ctx.preprocess(d[0], *a)
ctx.preprocess(d[1], *b)
r0, r1 = row_range(rank, world, cfg.height)
lib = L.load_library()
IT, SL = 24, 8
for rep in range(4):
    ctx.icp_reset(True)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.icp_align_rows(b[0], b[1], a[0], a[1], r0, r1, 20)
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print("align x20 rows [%d,%d) by events: %.1f us" % (r0, r1, e0.elapsed_time(e1) * 1e3))
tr_ = np.zeros(1024 * IT * SL, np.uint64)
lib.vh_align_trace_read.argtypes = [C.c_void_p, C.c_int]
lib.vh_align_trace_read(tr_.ctypes.data, tr_.size)
t = tr_.reshape(1024, IT, SL).astype(np.int64)
n = int((t[:, 0, 0] > 0).sum())
t = t[:n, 1:20]
if rank == 0:
    print("world", world, "CTAs", n)
    print("median over CTAs and iterations 1..19 (ns): main %d, block reduce %d, local exchange %d, cross-GPU exchange (CTA 0 warp 0) %d, solve %d, period %d" % (
        np.median(t[:, :, 1] - t[:, :, 0]), np.median(t[:, :, 2] - t[:, :, 1]), np.median(t[:, :, 3] - t[:, :, 2]),
        np.median(t[0, :, 5] - t[0, :, 3]), np.median(t[0, :, 6] - t[0, :, 5]), np.median(np.diff(t[:, :, 0].min(axis=0)))))
    print("per-iteration cross-GPU exchange on CTA 0 (ns):", (t[0, :, 5] - t[0, :, 3]).tolist())
    print("all CTAs: after local exchange -> solve start (ns): median %d, max %d" % (np.median(t[:, :, 5] - t[:, :, 3]), np.median((t[:, :, 5] - t[:, :, 3]).max(axis=0))))
dist.barrier()
dist.destroy_process_group()
