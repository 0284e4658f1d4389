"""GPU box: k_alloc time on config C4 for an unpartitioned table and for rank 0 of an 8-way partition (steady state)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from voxelhashing_demo_b200 import Context, scenes  # noqa: E402

for parts in (1, 8):
    cfg, scene, traj, _ = bench.workload_config(sys.argv[1] if len(sys.argv) > 1 else "C4", parts, 0)
    ctx = Context(cfg)
    pose = traj(0).astype(np.float32)
    d = torch.from_numpy(scenes.render_depth(scene, pose, cfg.width, cfg.height, cfg.fx, cfg.fy, cfg.cx, cfg.cy).reshape(-1)).cuda()
    v, n, df = ctx.new_maps()
    ctx.preprocess(d, v, n, df)
    ctx.set_pose(pose)
    ts = []
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.alloc_blocks(v, n)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    st = ctx.stats()
    print(f"parts {parts}: first (cold, inserting) {ts[0]:.1f} us, steady {np.median(ts[2:]):.1f} us, allocated {st.numAllocated}")
    ctx.close()
