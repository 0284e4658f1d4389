"""GPU-box micro-timing of the ICP iteration variants (CUDA events, warm L2)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
from voxelhashing_demo_b200 import Context  # noqa: E402

cfg, scene, traj, _ = bench.workload_config("C2")
ctx = Context(cfg)
frames, poses = bench.render_frames(cfg, scene, traj, 2)
d = torch.from_numpy(frames).cuda()
a, b = ctx.new_maps(), ctx.new_maps()
ctx.preprocess(d[0], *a)
ctx.preprocess(d[1], *b)
sysbuf = torch.zeros(32, device="cuda")
s = torch.cuda.Stream()


def timeit(fn, n=200):
    with torch.cuda.stream(s):
        for _ in range(20):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(n):
            fn()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n




def graph_time(fn, reps=20):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn()
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    return timeit(lambda: g.replay(), 50) / reps


H = cfg.height
print("GPU time per launch inside a CUDA graph of 20 identical launches (us):")
print("  reduce full image       :", graph_time(lambda: ctx.icp_reduce(b[0], b[1], a[0], a[1], 0, H, sysbuf, s)))
print("  reduce half image       :", graph_time(lambda: ctx.icp_reduce(b[0], b[1], a[0], a[1], 0, H // 2, sysbuf, s)))
print("  reduce 8 rows           :", graph_time(lambda: ctx.icp_reduce(b[0], b[1], a[0], a[1], 0, 8, sysbuf, s)))
print("  solve only              :", graph_time(lambda: ctx.icp_solve(sysbuf, s)))
ctx.icp_reset(True, s)
print("  iterate (reduce+solve)  :", graph_time(lambda: ctx.icp_iterate(b[0], b[1], a[0], a[1], s)))
print("  align x20 (PDL chain)/20:", graph_time(lambda: ctx.icp_align(b[0], b[1], a[0], a[1], 20, s), 1) / 20)
print("  set_pose                :", graph_time(lambda: ctx.set_pose(np.eye(4, dtype=np.float32), s)))
print("  preprocess              :", graph_time(lambda: ctx.preprocess(d[0], *a, s)))
print("  alloc                   :", graph_time(lambda: ctx.alloc_blocks(a[0], a[1], s)))
print("  compact                 :", graph_time(lambda: ctx.compact(s)))
print("  integrate               :", graph_time(lambda: ctx.integrate_depthf(a[2], s)))
print("python launch overhead, no graph (us/launch): solve", timeit(lambda: ctx.icp_solve(sysbuf, s)))
