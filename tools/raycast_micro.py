"""GPU-box: time the raycast stages at VGA (config C2 model after 30 fused frames)."""
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
frames, poses = bench.render_frames(cfg, scene, traj, 30)
d = torch.from_numpy(frames).cuda()
v, n, df = ctx.new_maps()
for i in range(30):
    ctx.preprocess(d[i], v, n, df)
    ctx.fuse_frame(poses[i].astype(np.float32), v, n, df)
rv, rn = torch.zeros_like(v), torch.zeros_like(n)
s = torch.cuda.Stream()


def timeit(fn, reps=50):
    with torch.cuda.stream(s):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            fn()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


ctx.set_pose(poses[29].astype(np.float32), s)
print("raycast (compact + interval splat + march) us:", round(timeit(lambda: ctx.raycast(rv, rn, s)), 1))
print("compact alone us:", round(timeit(lambda: ctx.compact(s)), 1))
hit = (rv[:, 2] > 0).sum().item()
print("hits", hit, "of", cfg.width * cfg.height, " visible blocks", ctx.stats().numVisible)
