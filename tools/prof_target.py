"""Minimal launch sequences for ncu (GPU box):  python tools/prof_target.py integrate|icp|frame|raycast
A number printed by a run under ncu is never a bench value; this only produces kernels to capture."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench  # noqa: E402
from voxelhashing_demo_b200 import Context, FramePipeline, scenes  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "integrate"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3

if what == "integrate":          # large-volume integrate: working set > 2x L2
    cfg, scene, traj, _ = bench.workload_config("C4_4mm")
    ctx = Context(cfg)
    pose = traj(0).astype(np.float32)
    d = torch.from_numpy(scenes.render_depth(scene, pose, cfg.width, cfg.height, cfg.fx, cfg.fy, cfg.cx, cfg.cy).reshape(-1)).cuda()
    v, n, df = ctx.new_maps()
    ctx.preprocess(d, v, n, df)
    ctx.set_pose(pose)
    ctx.alloc_blocks(v, n)
    ctx.compact()
    for _ in range(reps):
        ctx.integrate_depthf(df)
    ctx.garbage_collect(1, 0.0, 0.0)     # k_gc over the same 1.46 GB (read-only scan)
    torch.cuda.synchronize()
    st = ctx.stats()
    print("visible", st.numVisible, "updated", st.numUpdated, "alloc", st.numAllocated, "dropped", st.dropped)
elif what == "alloc":            # insert-heavy allocation: the FIRST frame of the 4 mm large-volume scene (355 k new blocks), then a steady-state pass
    cfg, scene, traj, _ = bench.workload_config("C4_4mm")
    ctx = Context(cfg)
    pose = traj(0).astype(np.float32)
    d = torch.from_numpy(scenes.render_depth(scene, pose, cfg.width, cfg.height, cfg.fx, cfg.fy, cfg.cx, cfg.cy).reshape(-1)).cuda()
    v, n, df = ctx.new_maps()
    ctx.preprocess(d, v, n, df)
    ctx.set_pose(pose)
    ctx.alloc_blocks_depth(d)        # launch 0: every block is new
    ctx.alloc_blocks_depth(d)        # launch 1: every block is present
    torch.cuda.synchronize()
    st = ctx.stats()
    print("alloc", st.numAllocated, "dropped", st.dropped, "overflow", st.overflowUsed)
elif what in ("icp", "frame", "raycast"):
    cfg, scene, traj, _ = bench.workload_config("C2")
    ctx = Context(cfg)
    frames, poses = bench.render_frames(cfg, scene, traj, 8)
    d = torch.from_numpy(frames).cuda()
    mode = FramePipeline.FRAME_TO_MODEL if what == "raycast" else FramePipeline.FRAME_TO_FRAME
    pipe = FramePipeline(ctx, iterations=20, mode=mode, use_graph=(what != "icp"))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        pipe.reset(poses[0].astype(np.float32))
        for r in range(reps):
            for i in range(8):
                pipe.push_device(d[i])
        print(pipe.pose())
