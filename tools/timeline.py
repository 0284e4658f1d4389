"""GPU-box: interleaving of the frame loop's kernels over the streams, read from %globaltimer stamps (library built with
-DVH_TIMELINE).  Single GPU:  python tools/timeline.py [C2|C3|C4]      multi-GPU: torchrun ... tools/timeline.py C4"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["VH_EXTRA_NVCC_FLAGS"] = "-DVH_TIMELINE"
os.environ["VH_TIMELINE"] = "1"
import torch  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from voxelhashing_demo_b200 import _build  # noqa: E402

if rank == 0:
    _build.build(force=True)
if world > 1:
    dist.barrier()
import bench  # noqa: E402
from voxelhashing_demo_b200 import Context  # noqa: E402
from voxelhashing_demo_b200 import lib as L  # noqa: E402
from voxelhashing_demo_b200.dist import PartitionedTracker  # noqa: E402

name = next((a for a in sys.argv[1:] if a in ("C2", "C3", "C4")), "C2")
cfg, scene, traj, _ = bench.workload_config(name, world, rank)
ctx = Context(cfg)
tr = PartitionedTracker(ctx, rank, world, overlap=True)
n = 12
frames, poses = bench.render_frames(cfg, scene, traj, n) if rank == 0 else (None, [traj(k) for k in range(n)])
d = torch.from_numpy(frames).cuda() if rank == 0 else None
lib = L.load_library()
lib.vh_timeline_read.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
buf = np.zeros(1 << 16, np.uint64)
tr.reset(poses[0].astype(np.float32))
for rep in range(2):
    for i in range(n):
        tr.push(d[i] if rank == 0 else None, input_ready=True)
    tr.flush()
    torch.cuda.synchronize()
    if rep == 0:
        lib.vh_timeline_read(ctx._h, buf.ctypes.data, buf.size, 1)          # discard the warm-up pass
cnt = lib.vh_timeline_read(ctx._h, buf.ctypes.data, buf.size, 1)
if rank == 0:
    ev = sorted((int(w >> 8), int((w >> 1) & 127), int(w & 1)) for w in buf[:cnt].tolist())
    names = {1: "preprocess", 2: "ALIGN", 3: "set_frame", 4: "alloc", 5: "compact", 6: "INTEGRATE"}
    t0 = ev[0][0]
    print(f"{name}, {world} GPU(s): kernel begin (CTA 0) / end (CTA 0 of the persistent kernels), us since the first stamp")
    for t, k, ph in ev:
        if (t - t0) / 1e3 > 6000:
            break
        print(f"{(t - t0) / 1e3:9.1f}  {'  ' if k in (1, 2) else '                    '}{names.get(k, k)} {'end' if ph else 'begin'}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
