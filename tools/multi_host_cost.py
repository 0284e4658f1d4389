"""GPU box, torchrun: where does the HOST time of a partitioned frame go?  (perf_counter around each enqueue, no syncs)"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from voxelhashing_demo_b200 import Context  # noqa: E402
from voxelhashing_demo_b200.dist import PartitionedTracker  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg, scene, traj, _ = bench.workload_config("C4", world, rank)
frames, poses = bench.render_frames(cfg, scene, traj, 4) if rank == 0 else (None, [traj(k) for k in range(4)])
d_frames = torch.from_numpy(frames).cuda() if rank == 0 else None
ctx = Context(cfg)
tr = PartitionedTracker(ctx, rank, world)
tr.reset(poses[0].astype(np.float32))
for i in range(10):
    tr.push(d_frames[i % 4] if rank == 0 else None)
torch.cuda.synchronize()
dist.barrier()
acc = {"copy": 0.0, "bcast": 0.0, "push": 0.0}
N = int(os.environ.get("NFRAMES", "100"))
t_all = time.perf_counter()
for i in range(N):
    t0 = time.perf_counter()
    if rank == 0:
        pass
    t1 = time.perf_counter()
    dist.broadcast(tr.depth.view(torch.uint8), src=0)
    t2 = time.perf_counter()
    tr.pipe.push_device(tr.depth)
    t3 = time.perf_counter()
    acc["copy"] += t1 - t0; acc["bcast"] += t2 - t1; acc["push"] += t3 - t2
host = time.perf_counter() - t_all
torch.cuda.synchronize()
total = time.perf_counter() - t_all
print(f"rank {rank}: host {host / N * 1e6:.0f} us/frame (copy {acc['copy'] / N * 1e6:.0f}, bcast {acc['bcast'] / N * 1e6:.0f}, push {acc['push'] / N * 1e6:.0f}); wall incl. GPU {total / N * 1e6:.0f} us/frame", flush=True)
dist.destroy_process_group()
