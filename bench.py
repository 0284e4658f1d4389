#!/usr/bin/env python
"""bench.py -- headline benchmark of the fusion-and-tracking hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C2|C3|C4|C5]

A "step" is one depth frame through the whole hot path: preprocess -> ICP (20 Gauss-Newton iterations,
1 level) -> pose update -> block allocation -> visible-block compaction -> TSDF integration.

  N = 1 (default): config C2 of BASELINE.json -- 300-frame synthetic VGA sequence with known
        trajectory on one B200.  `value` = frames/s with all frames resident in HBM; `e2e` = the same
        through the public C ABI with HOST buffers (H2D of every depth frame, D2H of every pose).
        The K-step pass is repeated (default 5x, each after its own W warm-up steps) and the MEDIAN is
        reported, so a 3 ms timed region is not a single sample.
  N > 1 (torchrun): config C4 AS BASELINE.json STATES IT -- building-scale scene at 2 mm voxels, block-hash
        space partitioned over the ranks (owner = mix(block) mod N, per-GPU 1 048 576 blocks / 4 000 037
        buckets), depth frame broadcast over NVLink, ICP image rows split with the 32-float all-reduce fused
        into the persistent Align kernel over peer memory.  Strong scaling: every rank works on the same
        frames; the line carries the SAME workload on one GPU (`single_gpu_same_workload`), the per-rank
        integrate roofline, per-stage times and the cross-rank equivalence checks.
        (A torchrun launch with WORLD_SIZE=1, or --workload C4, runs that workload on one GPU.)
  --impl reference: the reference's own CUDA kernels rebuilt for sm_100a (oracle/_ref/libvh_ref.so,
        driven with the reference's own call sequence and host syncs) on the same frames and the same table
        geometry (the reference has no CPU path); its printf-stripped and shipped -G builds and the CPU
        transliteration (oracle/, OpenMP, 1 thread and all cores) are timed next to it.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "VGA frames/s (alloc+integrate+ICP)"
UNIT = "frames/s"

# Libraries print to stdout behind our back (NCCL's banners, the reference's std::cout and device
# printf).  The contract is ONE JSON line on stdout: keep the real stdout aside, point fd 1 at stderr.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj) -> None:
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


# ---------------------------------------------------------------------------------------------------
C4_VOXEL = 0.002                     # BASELINE.json configs[3]: "Building-scale synthetic scene at 2 mm voxels"
C4_BLOCKS_PER_GPU = 1048576          # per-GPU numVoxelBlocks (4 GiB of voxels)
C4_BLOCKS_SINGLE = 4000000           # one GPU holding the whole volume (ptr = id * 512 fits the reference's int)


def workload_config(name: str, part_count: int = 1, part_rank: int = 0, policy=None):
    from voxelhashing_demo_b200 import POLICY_FIXED, Config, scenes

    pol = POLICY_FIXED if policy is None else policy
    if name in ("C2", "C5"):
        cfg = Config(policy=pol, numBuckets=100003, bucketSize=5, numVoxelBlocks=65536, voxelSize=0.02,
                     truncation=0.06, truncScale=0.01, overflowSlots=16384 if pol == POLICY_FIXED else 0,
                     icpNormalThres=0.8 if pol == POLICY_FIXED else -1.0, icpIterations=20,
                     partCount=part_count, partRank=part_rank)
        return cfg, scenes.scene_S1T(), scenes.trajectory_C2, 300
    if name == "C3":
        cfg = Config(policy=pol, width=1280, height=720, fx=1034.6, fy=1033.0, cx=637.2, cy=382.95,
                     numBuckets=1000003, bucketSize=5, numVoxelBlocks=262144, voxelSize=0.005, truncation=0.02,
                     truncScale=0.0025, overflowSlots=131072, depthMax=8.0, maxIntegrationDistance=8.0,
                     icpNormalThres=0.8, icpIterations=20, partCount=part_count, partRank=part_rank)
        return cfg, scenes.scene_S2(), (lambda k: scenes.trans(0, 0, 0.3) @ scenes.trajectory_C3(k)), 1000
    if name in ("C4", "C4_4mm"):
        # C4: 2 mm voxels (1.6 cm blocks), truncation band 4 voxels + 0.1 % of the depth.  C4_4mm: the 4 mm variant the
        # integrate roofline leg has used since r1 (1.46 GB visible set on one GPU).
        vs = C4_VOXEL if name == "C4" else 0.004
        blocks = (C4_BLOCKS_PER_GPU if part_count > 1 else C4_BLOCKS_SINGLE) if name == "C4" else 2097152
        cfg = Config(policy=pol, width=1280, height=720, fx=1034.6, fy=1033.0, cx=637.2, cy=382.95,
                     numBuckets=4000037, bucketSize=5, numVoxelBlocks=blocks, voxelSize=vs, truncation=4 * vs,
                     truncScale=vs / 2, overflowSlots=524288, depthMax=12.5, maxIntegrationDistance=12.5,
                     icpNormalThres=0.8, icpIterations=20, partCount=part_count, partRank=part_rank)
        return cfg, scenes.scene_S3(), (lambda k: scenes.trans(0, 0, 0.5) @ scenes.trajectory_C3(k)), 200
    raise SystemExit(f"unknown workload {name}")


def config_dict(name: str, cfg, n_unique: int) -> dict:
    """The `config` object of the JSON line: the workload only (identical in the own arm and the reference arm)."""
    what = {"C2": "C2: synthetic 640x480 sequence (analytic sphere+plane scene S1T, known trajectory, 300 frames)",
            "C5": "C5: C2 tracked frame-to-MODEL (CUDA hash-table raycast in the loop)",
            "C3": "C3: synthetic 1280x720 sequence (room scene S2, 5 mm voxels, 1000 frames)",
            "C4": "C4: building-scale synthetic 1280x720 sequence (hall scene S3) at 2 mm voxels",
            "C4_4mm": "C4 geometry at 4 mm voxels (integrate roofline leg)"}[name]
    return {"workload": what, "frames_cycled": n_unique, "image": f"{cfg.width}x{cfg.height}", "voxel_m": cfg.voxelSize,
            "table": f"{cfg.numBuckets}x{cfg.bucketSize} buckets", "truncation_m": cfg.truncation,
            "icp": f"{cfg.icpIterations} iterations x 1 level",
            "l2": f"inputs larger than L2: {n_unique} distinct u16 frames cycled (ping-pong); no L2 flush between steps"}


def render_frames(cfg, scene, traj, count):
    from voxelhashing_demo_b200 import scenes

    poses = [traj(k) for k in range(count)]
    frames = np.stack([scenes.render_depth(scene, p, cfg.width, cfg.height, cfg.fx, cfg.fy, cfg.cx, cfg.cy, cfg.depthScale)
                       for p in poses])
    return frames.reshape(count, -1), poses


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe's clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = 0.0

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.idx), "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            # nvidia-smi needs 0.1-1 s to start (longer on an 8-GPU box with 8 ranks doing the same): wait for its first
            # row so that the sampler is already streaming when the timed region begins
            deadline = time.time() + 5.0
            while not self.rows and time.time() < deadline:
                time.sleep(0.01)
        except OSError:
            self.proc = None
        self.t0 = time.time()
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *a):
        self.t1 = time.time()
        if self.proc:
            time.sleep(0.06)                     # one more period: the row covering the end of the region
            self.proc.terminate()
            with contextlib.suppress(Exception):
                self.proc.wait(timeout=2)

    def summary(self) -> dict:
        inside = [r for t, r in self.rows if self.t0 <= t <= self.t1 + 0.06]
        note = None
        if not inside and self.rows:             # region shorter than one sampling period: nearest row
            inside = [min(self.rows, key=lambda tr: abs(tr[0] - self.t1))[1]]
            note = "timed region shorter than the 50 ms sampling period: nearest sample"
        sm, mx, reasons = [], [], set()
        for r in inside:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


def ncu_traffic(kernel_prefix: str):
    """dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full capture (profiles/)."""
    for f in sorted((ROOT / "profiles").glob("r*_traffic.json"), reverse=True):
        if f.name.startswith("r1a"):
            continue
        for k, v in json.loads(f.read_text()).get("bytes", {}).items():
            if k.startswith(kernel_prefix):
                return float(v)
    return None


def peaks() -> tuple[float, str]:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


@contextlib.contextmanager
def silence_stdout():
    """The reference prints from host and device code on every call; keep the JSON line clean."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)


# ---------------------------------------------------------------------------------------------------
def cpu_port_baseline(cfg, frames, pose0, budget_s=12.0, max_frames=40, threads=0):
    """The host transliteration (oracle/, OpenMP) on a bounded sample of the same workload, timed on
    this box's cores: preprocess + 20-iteration Align + alloc + compact + integrate per frame."""
    from oracle import binding as ob

    lib = ob.oracle_lib()
    all_cores = os.cpu_count() or lib.vo_num_threads()      # torchrun exports OMP_NUM_THREADS=1: ask for the cores explicitly
    lib.vo_set_num_threads(threads if threads else all_cores)
    cores = lib.vo_num_threads()
    ot = ob.OracleTable(cfg)
    pose = pose0.astype(np.float32)
    prev = None
    est = np.zeros(6, np.float32)
    done, t0 = 0, time.perf_counter()
    for k in range(min(max_frames, len(frames))):
        v, n, df = ot.preprocess(frames[k])
        if prev is not None:
            _, est, delta = ob.icp_align(cfg, v, n, prev[0], prev[1], cfg.icpIterations, est)
            pose = (pose.astype(np.float64) @ delta.astype(np.float64)).astype(np.float32)
        ot.fuse_frame(pose, v, df)
        prev = (v, n)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    ot.close()
    lib.vo_set_num_threads(all_cores)
    return {"value": done / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {done} frames of the workload ({dt:.1f} s): oracle/ C++ port, OpenMP x{cores}, "
                      f"preprocess + {cfg.icpIterations}-iteration Align + alloc + compact + integrate per frame"}


def cpu_baselines(cfg, frames, pose0, budget_s=12.0):
    """All cores (the `cpu_baseline` object) and one thread (both are reported)."""
    allc = cpu_port_baseline(cfg, frames, pose0, budget_s=budget_s)
    one = cpu_port_baseline(cfg, frames, pose0, budget_s=budget_s * 0.75, threads=1)
    allc["one_thread"] = {"value": one["value"], "unit": UNIT, "cores": 1, "sample": one["sample"]}
    return allc


def median_passes(run_pass, repeats: int):
    """run_pass() -> (ms, extra); returns (median ms, all ms, extra of the median pass)."""
    res = [run_pass() for _ in range(max(1, repeats))]
    order = sorted(range(len(res)), key=lambda i: res[i][0])
    mid = order[len(order) // 2]
    return res[mid][0], [r[0] for r in res], res[mid][1]


# ---------------------------------------------------------------------------------------------------
def run_own(args):
    import torch

    from voxelhashing_demo_b200 import POLICY_REF_EXACT, Context, FramePipeline, _build

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    _build.build()
    under_torchrun = "RANK" in os.environ and "WORLD_SIZE" in os.environ
    if world > 1 or (under_torchrun and args.workload is None) or args.workload == "C4":
        return run_partitioned(args, rank, world, local)

    name = args.workload or "C2"
    cfg, scene, traj, seq_len = workload_config(name)
    K, W = args.steps, args.warmup
    n_unique = min(seq_len, max(K + W, 2))
    frames, poses = render_frames(cfg, scene, traj, n_unique)
    from voxelhashing_demo_b200.scenes import pingpong

    order = [pingpong(i, n_unique) for i in range(W + K)]
    d_frames = torch.from_numpy(frames).cuda()                       # all frames resident in HBM
    h_frames = torch.from_numpy(frames).pin_memory()                 # e2e: pinned host buffers
    h_pose = torch.zeros((W + K, 16), dtype=torch.float32).pin_memory()
    frame_bytes = frames.shape[1] * 2
    stream = torch.cuda.Stream()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def push_dev(pipe, frame):
        # the frames are resident in HBM and nothing on `stream` produces them: the "ready" entry point lets the pipeline
        # pre-process frame k+1 beside the Align of frame k (--ready 0: vh_pipeline_push_device, ordered behind the stream)
        return pipe.push_device_ready(frame, None, stream) if args.ready else pipe.push_device(frame, stream)

    def sequence_runner(ctx, cfg_):
        def run_sequence(host: bool, sample_clocks: bool = False):
            ctx.reset(stream)
            mode = FramePipeline.FRAME_TO_MODEL if name == "C5" else FramePipeline.FRAME_TO_FRAME
            pipe = FramePipeline(ctx, iterations=cfg_.icpIterations, mode=mode, use_graph=True, overlap=bool(args.overlap))
            with torch.cuda.stream(stream):
                pipe.reset(poses[order[0]].astype(np.float32), stream)
                for i in range(W):
                    (pipe.push_host(h_frames[order[i]], h_pose[i], stream) if host else push_dev(pipe, d_frames[order[i]]))
                pipe.flush(stream)
                stream.synchronize()
                l0 = pipe.launches()
                with (ClockSampler(local) if sample_clocks else contextlib.nullcontext()) as cs:
                    torch.cuda.synchronize()
                    ev0.record(stream)
                    for i in range(W, W + K):
                        (pipe.push_host(h_frames[order[i]], h_pose[i], stream) if host else push_dev(pipe, d_frames[order[i]]))
                    pipe.flush(stream)                      # overlapped schedule: the last frame's fusion belongs to the timed region
                    ev1.record(stream)
                    stream.synchronize()
                ms = ev0.elapsed_time(ev1)
                launches = pipe.launches() - l0
                pose = pipe.pose(stream)
            pipe.close()
            return ms, (launches, pose, cs.summary() if sample_clocks else None)
        return run_sequence

    ctx = Context(cfg)
    run_sequence = sequence_runner(ctx, cfg)
    # clocks are sampled over ALL timed passes of the device-resident leg (one nvidia-smi stream around them)
    with ClockSampler(local) as cs_all:
        ms, passes, (launches, pose, _) = median_passes(lambda: run_sequence(False), args.repeats)
    clocks = cs_all.summary()
    clocks["note"] = f"sampled every 50 ms across the {len(passes)} timed passes (warm-ups between them included)"
    ms_e2e, passes_e2e, (_, pose_e2e, _) = median_passes(lambda: run_sequence(True), args.repeats)
    truth = poses[order[-1]]
    pose_err = float(np.max(np.abs(pose[:3, 3] - truth[:3, 3])))
    st = ctx.stats()

    stages, roof = stage_timings(ctx, cfg, d_frames, poses, order, stream)
    hbm = None
    if not args.no_hbm:
        # the large-volume config as BASELINE.json states it (2 mm: 4.9 GB visible set) and the smaller 4 mm set r1 reported
        hbm = integrate_hbm_roofline(stream, "C4")
        hbm["set_4mm"] = {k: v for k, v in integrate_hbm_roofline(stream, "C4_4mm").items() if k != "gc_scan"}
    ctx.close()

    # the repo's OWN kernels in the reference's arithmetic (RefExact policy) on the same frames and table: the
    # like-for-like leg next to the reference arm (same config, same policy)
    refexact = None
    if name == "C2" and not args.no_refexact:
        cfg_r, _, _, _ = workload_config("C2", policy=POLICY_REF_EXACT)
        ctx_r = Context(cfg_r)
        run_r = sequence_runner(ctx_r, cfg_r)
        ms_r, passes_r, (launches_r, _, _) = median_passes(lambda: run_r(False), min(3, args.repeats))
        ms_re, _, _ = median_passes(lambda: run_r(True), min(3, args.repeats))
        refexact = {"policy": "RefExact (the reference's as-built arithmetic, SURVEY.md Appendix A; parity-pinned to the reference's CUDA outputs)",
                    "value": K / (ms_r / 1e3), "unit": UNIT, "ms_per_step": ms_r / K, "passes_ms": passes_r,
                    "e2e": {"value": K / (ms_re / 1e3), "unit": UNIT}, "gpu_launches": int(launches_r)}
        ctx_r.close()
    cpu = None if args.no_cpu else cpu_baselines(cfg, frames, poses[0])

    out = {
        "metric": METRIC, "value": K / (ms / 1e3), "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_dict(name, cfg, n_unique),
        "policy": "Fixed (corrected pipeline, DESIGN.md section 4; bit-exact against its own oracle, anchored to the analytic scene)",
        "passes": {"repeats": len(passes), "timed_ms": passes, "e2e_timed_ms": passes_e2e, "reported": "median"},
        "results": {"final_pose_translation_error_m": pose_err, "visible_blocks": st.numVisible, "allocated_blocks": st.numAllocated,
                    "voxel_updates_per_frame": int(st.numUpdated),
                    "e2e_pose_equals_device_resident_pose": bool(np.array_equal(pose.view(np.uint32), pose_e2e.view(np.uint32)))},
        "e2e": {"value": K / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": frame_bytes, "d2h_bytes_per_step": 64,
                "ms_per_step": ms_e2e / K},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "stages_us": stages,
        "voxel_updates_per_s": int(st.numUpdated) / (stages["integrate"] * 1e-6) if stages.get("integrate") else None,
    }
    if hbm:
        out["roofline_integrate_hbm"] = hbm
    if refexact:
        out["refexact_leg"] = refexact
    if cpu:
        out["cpu_baseline"] = cpu
    emit(out)


def stage_timings(ctx, cfg, d_frames, poses, order, stream, reps: int = 8):
    """Instrumented pass: each stage timed with CUDA events on the launching stream around `reps` back-to-back launches
    (median over frames of the steady-state model), so launch latency is not counted as kernel time.  Gives the
    per-kernel durations the roofline object needs."""
    import torch

    n = cfg.width * cfg.height
    maps = [ctx.new_maps(), ctx.new_maps()]
    acc = {k: [] for k in ("preprocess", "icp_align", "alloc", "compact", "integrate")}
    peak, which = peaks()

    def timed(fn, r=reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()                                               # one untimed launch: the GPU is awake, caches as in steady state
        e0.record(stream)
        for _ in range(r):
            fn()
        e1.record(stream)
        stream.synchronize()
        return e0.elapsed_time(e1) * 1e3 / r

    with torch.cuda.stream(stream):
        ids = order[-min(len(order), 12):]
        for j, fi in enumerate(ids):
            cur, prev = maps[j & 1], maps[1 - (j & 1)]
            acc["preprocess"].append(timed(lambda: ctx.preprocess(d_frames[fi], cur[0], cur[1], cur[2], stream)))
            if j > 0:
                def align():
                    ctx.icp_reset(True, stream)
                    ctx.icp_align(cur[0], cur[1], prev[0], prev[1], cfg.icpIterations, stream)
                acc["icp_align"].append(timed(align, 4))
            ctx.set_pose(poses[fi].astype(np.float32), stream)
            acc["alloc"].append(timed(lambda: ctx.alloc_blocks(cur[0], cur[1], stream)))
            acc["compact"].append(timed(lambda: ctx.compact(stream)))
            acc["integrate"].append(timed(lambda: ctx.integrate_depthf(cur[2], stream)))
        st = ctx.stats(stream)
    stages = {k: float(np.median(v)) for k, v in acc.items() if v}
    stages["icp_iteration"] = stages["icp_align"] / cfg.icpIterations
    stages["note"] = (f"CUDA events around {reps} back-to-back launches per sample (4 for the Align, each preceded by the 2 us k_icp_reset); "
                      "inside the frame loop the fusion stages run on a second stream beside the Align of the next frame")
    # dominant kernel of the step: the persistent Align kernel (one launch per frame, 20 Gauss-Newton iterations)
    per_px = 64 if cfg.icpNormalThres > -1.0 else 48       # source vertex + gathered target vertex + normal (+ source normal)
    icp_bytes = per_px * n * cfg.icpIterations
    t_icp = stages["icp_align"] * 1e-6
    ach = icp_bytes / t_icp / 1e9
    integ_bytes = 16 * int(st.numUpdated) + 16 * st.numVisible + 4 * n
    roof = {"kernel": "k_icp_align<Fixed> (persistent: one launch = one Align = 20 iterations)", "bound": "hbm", "achieved": ach, "peak": peak,
            "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic("k_icp_align"), "peak_source": which,
            "algorithmic_bytes": icp_bytes, "us": stages["icp_align"],
            "note": f"dominant kernel of the step (~85 % of the frame): {per_px} B x W x H x {cfg.icpIterations} iterations of algorithmic bytes per launch / "
                    "launch time (CUDA events on the launching stream).  The 20 MB working set is L1/L2-resident at VGA -- `traffic` (ncu, "
                    "cold L2) is ~1/16 of the algorithmic bytes -- so this is NOT an HBM utilisation: the kernel is bound by instruction issue "
                    "in the association loop (ncu: 64 % issue utilisation there) and by the per-iteration exchange + 6x6 solve "
                    "(DESIGN.md 3.2).  The HBM-bound kernel of the path is k_integrate: roofline_integrate_hbm",
            "integrate_c2": {"bytes": integ_bytes, "us": stages["integrate"], "GB/s": integ_bytes / (stages["integrate"] * 1e-6) / 1e9,
                             "note": "L2-resident working set: not an HBM fraction"}}
    return stages, roof


def integrate_hbm_roofline(stream, name: str = "C4_4mm"):
    """Integration over a visible set larger than 2x L2 (config C4 geometry on one GPU): the HBM-bound regime
    north_star quotes its 60 % target on.  achieved = (16 N_upd + 16 N_vis + 4 W H) / t."""
    import torch

    from voxelhashing_demo_b200 import Context, scenes

    cfg, scene, traj, _ = workload_config(name)
    ctx = Context(cfg)
    pose = traj(0).astype(np.float32)
    depth = scenes.render_depth(scene, pose, cfg.width, cfg.height, cfg.fx, cfg.fy, cfg.cx, cfg.cy, cfg.depthScale)
    d = torch.from_numpy(depth.reshape(-1)).cuda()
    v, nm, df = ctx.new_maps()
    times = []
    with torch.cuda.stream(stream):
        ctx.preprocess(d, v, nm, df, stream)
        ctx.set_pose(pose, stream)
        ctx.alloc_blocks(v, nm, stream)
        ctx.compact(stream)
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.integrate_depthf(df, stream)
            e1.record(stream)
            stream.synchronize()
            if i >= 3:
                times.append(e0.elapsed_time(e1) * 1e-3)
        st = ctx.stats(stream)
        # garbage-collection scan of the same model (read-only here: no ageing, nothing qualifies for release)
        gc_times = []
        for i in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.garbage_collect(1, 0.0, 0.0, stream)
            e1.record(stream)
            stream.synchronize()
            if i >= 2:
                gc_times.append(e0.elapsed_time(e1) * 1e-3)
        freed = ctx.stats(stream).lastFreed
    peak, which = peaks()
    n = cfg.width * cfg.height
    nbytes = 16 * int(st.numUpdated) + 16 * st.numVisible + 4 * n
    t = float(np.mean(times))
    ctx.close()
    return {"kernel": "k_integrate<Fixed,dense>", "bound": "hbm", "achieved": nbytes / t / 1e9, "peak": peak, "unit": "GB/s",
            "frac": nbytes / t / 1e9 / peak, "traffic": ncu_traffic("k_integrate"), "algorithmic_bytes": nbytes,
            "peak_source": which, "us": t * 1e6, "voxel_m": cfg.voxelSize,
            "visible_blocks": st.numVisible, "voxel_working_set_MB": st.numVisible * 4096 / 1e6, "voxels_updated": int(st.numUpdated),
            "voxel_updates_per_s": int(st.numUpdated) / t, "dropped": st.dropped,
            "gc_scan": {"kernel": "k_gc (scope ALL, no ageing)", "bytes": 4096 * (st.numAllocated - freed), "us": float(np.mean(gc_times)) * 1e6,
                        "GB/s": 4096 * (st.numAllocated - freed) / float(np.mean(gc_times)) / 1e9,
                        "frac": 4096 * (st.numAllocated - freed) / float(np.mean(gc_times)) / 1e9 / peak, "released": int(freed)},
            "note": "working set > 2x L2 (126 MB), 5 timed launches after 3 warm-ups, no L2 flush needed"}


# ---------------------------------------------------------------------------------------------------
def run_partitioned(args, rank, world, local):
    """Config C4 (2 mm) with the block-hash space partitioned over `world` GPUs (world = 1: the whole volume on one GPU)."""
    import torch
    import torch.distributed as dist

    from voxelhashing_demo_b200 import Context
    from voxelhashing_demo_b200.dist import PartitionedTracker

    multi = world > 1
    if multi:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name = args.workload or "C4"
    cfg, scene, traj, seq_len = workload_config(name, world, rank)
    K, W = args.steps, args.warmup
    n_unique = min(seq_len, max(K + W, 2), 64)
    from voxelhashing_demo_b200.scenes import pingpong

    order = [pingpong(i, n_unique) for i in range(W + K)]
    frames = None
    if rank == 0:
        frames, poses = render_frames(cfg, scene, traj, n_unique)
        d_frames = torch.from_numpy(frames).cuda()
    else:
        poses = [traj(k) for k in range(n_unique)]
        d_frames = None
    stream = torch.cuda.current_stream()

    def timed_run(tracker, src, with_pose_readback=False, h_pose=None):
        """W warm-up pushes, then K timed ones; returns (ms on this rank, host enqueue ms)."""
        tracker.reset(poses[order[0]].astype(np.float32))
        for i in range(W):
            tracker.push(src[order[i]] if src is not None else None, input_ready=True)
        tracker.flush()
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        t_host0 = time.perf_counter()
        for i in range(W, W + K):
            tracker.push(src[order[i]] if src is not None else None, input_ready=True)
            if with_pose_readback:
                tracker.pose_async(h_pose[i - W])
        tracker.flush()                                  # the last frame's fusion belongs to the timed region
        e1.record(stream)
        host_ms = (time.perf_counter() - t_host0) * 1e3
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        return e0.elapsed_time(e1), host_ms

    # ---- the same workload on ONE GPU (rank 0, unpartitioned): the strong-scaling reference point of this very line
    single, single_pose, single_blocks = None, None, None
    if rank == 0 and multi and not args.no_single:
        cfg1, _, _, _ = workload_config(name, 1, 0)
        ctx1 = Context(cfg1)
        t1 = PartitionedTracker(ctx1, 0, 1, overlap=bool(args.overlap))       # default scheduling: what a 1-GPU user gets
        # (timed_run uses dist.barrier when multi: the single-GPU leg runs on rank 0 alone, so it is written out here)
        t1.reset(poses[order[0]].astype(np.float32))
        for i in range(W):
            t1.push(d_frames[order[i]], input_ready=True)
        t1.flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(W, W + K):
            t1.push(d_frames[order[i]], input_ready=True)
        t1.flush()
        e1.record()
        torch.cuda.synchronize()
        st1 = ctx1.stats()
        single_pose = t1.pose()
        single_blocks = st1.numAllocated
        single = {"value": K / (e0.elapsed_time(e1) / 1e3), "unit": UNIT, "ms_per_step": e0.elapsed_time(e1) / K, "n_gpus": 1,
                  "allocated_blocks": st1.numAllocated, "visible_blocks": st1.numVisible, "dropped": st1.dropped,
                  "voxel_updates_per_s": int(st1.numUpdated) / (e0.elapsed_time(e1) / K / 1e3)}
        del t1
        ctx1.close()
        torch.cuda.empty_cache()
    if multi:
        dist.barrier()

    ctx = Context(cfg)
    tuning = (args.align_ctas, args.reserve_sms) if args.align_ctas > 0 else None
    tracker = PartitionedTracker(ctx, rank, world, overlap=bool(args.overlap), tuning=tuning)
    l0 = tracker.launches
    with ClockSampler(local) as cs:
        ms_local, host_enqueue_ms = timed_run(tracker, d_frames if rank == 0 else None)
    launches = tracker.launches - l0

    def allmax(x: float) -> float:
        if not multi:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x: float) -> float:
        if not multi:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    ms = allmax(ms_local)                                             # device time, max over ranks
    st = ctx.stats()
    n_upd, n_vis, n_alloc, n_drop = allsum(st.numUpdated), allsum(st.numVisible), allsum(st.numAllocated), allsum(st.dropped)
    pose = tracker.pose()
    # ---- equivalence evidence carried by the line itself: pose bit-identical on every rank; equal to the 1-GPU run
    checks = {}
    if multi:
        pb = torch.from_numpy(pose.astype(np.float32).view(np.int32).astype(np.int64).reshape(-1)).cuda()
        lo, hi = pb.clone(), pb.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        checks["pose_bits_equal_across_ranks"] = bool(torch.equal(lo, hi))
        checks["fused_peer_exchange"] = bool(tracker.fused)
        if rank == 0 and single_pose is not None:
            checks["max_abs_pose_minus_single_gpu"] = float(np.max(np.abs(pose - single_pose)))
            checks["pose_within_1e-6_of_single_gpu"] = bool(checks["max_abs_pose_minus_single_gpu"] < 1e-6)
            # a pose that differs in its last bits can move a band end point across a block face: a handful of blocks may differ
            checks["allocated_blocks_minus_single_gpu"] = int(n_alloc) - int(single_blocks)
            checks["allocated_blocks_within_1e-5_of_single_gpu"] = bool(abs(int(n_alloc) - int(single_blocks)) <= 1e-5 * int(single_blocks))
    # ---- end to end: the ingest rank holds the frames in pinned HOST memory; every step copies one frame H2D, the
    # frame is broadcast, and every rank reads its pose back D2H
    h_frames = torch.from_numpy(frames).pin_memory() if rank == 0 else None
    h_pose = torch.zeros((K, 16), dtype=torch.float32).pin_memory()
    ms_e2e_local, _ = timed_run(tracker, h_frames, with_pose_readback=True, h_pose=h_pose)
    ms_e2e = allmax(ms_e2e_local)

    # ---- per-stage device times on this rank's partition (max over ranks): `reps` back-to-back launches per stage
    tracker.flush()
    torch.cuda.synchronize()
    stages = {}
    maps_a, maps_b = ctx.new_maps(), ctx.new_maps()
    d_prev, d_last = tracker.last_frames()                           # this rank's landing buffers: the frame before (ICP target) and the latest
    ctx.preprocess(d_last, *maps_a)
    ctx.preprocess(d_prev, *maps_b)
    rows = tracker.rows

    def timed(fn, r=4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        e0.record(stream)
        for _ in range(r):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / r

    def align():
        ctx.icp_reset(True)
        ctx.icp_align_rows(maps_a[0], maps_a[1], maps_b[0], maps_b[1], rows[0], rows[1], cfg.icpIterations)

    stages["preprocess"] = allmax(timed(lambda: ctx.preprocess(d_last, *maps_a)))
    stages["icp_align"] = allmax(timed(align))
    stages["alloc"] = allmax(timed(lambda: ctx.alloc_blocks(maps_a[0], maps_a[1])))
    stages["compact"] = allmax(timed(lambda: ctx.compact()))
    t_int = allmax(timed(lambda: ctx.integrate_depthf(maps_a[2])))
    stages["integrate"] = t_int
    st2 = ctx.stats()
    upd_stage = allsum(st2.numUpdated)
    n = cfg.width * cfg.height
    peak, which = peaks()
    # roofline of the stage that shards: per-rank algorithmic bytes / the SLOWEST rank's time
    bytes_rank = 16 * int(st2.numUpdated) + 16 * st2.numVisible + 4 * n
    frac_rank = bytes_rank / (t_int * 1e-6) / 1e9 / peak
    frac_min = -allmax(-frac_rank)
    bytes_all = allsum(bytes_rank)
    cpu = None
    if rank == 0 and not args.no_cpu:
        cfg1, _, _, _ = workload_config(name, 1, 0)
        cfg1.numVoxelBlocks = min(cfg1.numVoxelBlocks, 1048576)     # host memory: 4 GiB of voxels is plenty for the 2-frame sample
        cpu = cpu_port_baseline(cfg1, frames, poses[0], budget_s=10.0, max_frames=2)
    if rank == 0:
        truth = poses[order[-1]]
        out = {
            "metric": METRIC, "value": K / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": dict(config_dict(name, cfg, n_unique),
                           partition=f"hash space partitioned over {world} GPU(s) (owner = mix(block) mod {world}), {cfg.numVoxelBlocks} blocks per GPU; "
                                     + ("frame stored by rank 0 straight into every rank's landing buffer over NVLink (vh_dist, CUDA IPC, no NCCL on the data path)"
                                        if getattr(tracker, "_dist", None) is not None else "frame broadcast over NVLink (NCCL)") + "; 32-float ICP all-reduce "
                                     + ("FUSED into the persistent Align kernel over NVLink peer memory" if tracker.fused else
                                        ("through NCCL" if multi else "n/a (one GPU)")),
                           l2="per-frame voxel working set exceeds L2 on every rank"),
            "results": {"final_pose_translation_error_m": float(np.max(np.abs(pose[:3, 3] - truth[:3, 3]))),
                        "visible_blocks_all_ranks": int(n_vis), "allocated_blocks_all_ranks": int(n_alloc), "dropped_all_ranks": int(n_drop),
                        "voxel_working_set_MB_all_ranks": n_vis * 4096 / 1e6},
            "checks": checks,
            "voxel_updates_per_s": n_upd / (ms / K / 1e3),
            "stages_us": dict(stages, note="max over ranks of each stage on the rank's own partition, CUDA events around 4 back-to-back launches; "
                                           "in the frame loop the fusion stages of frame k run beside the tracking of frame k+1"),
            "roofline": {"kernel": "k_integrate<Fixed,dense> (the stage that shards; per-rank partition)", "bound": "hbm",
                         "achieved": bytes_all / (t_int * 1e-6) / 1e9 / world, "peak": peak, "unit": "GB/s",
                         "frac": bytes_all / (t_int * 1e-6) / 1e9 / world / peak, "frac_slowest_rank": frac_min,
                         "traffic": ncu_traffic("k_integrate"), "peak_source": which,
                         "algorithmic_bytes_all_ranks": bytes_all, "us_max_over_ranks": t_int,
                         "voxel_updates_per_s_all_ranks": upd_stage / (t_int * 1e-6),
                         "note": "achieved = mean per-GPU algorithmic bytes (16 N_upd + 16 N_vis + 4 W H) / the slowest rank's launch time"},
            "gpu_launches": int(launches),
            "scheduling": {"align_ctas": tracker.tuning[0], "sms_reserved_for_align": tracker.tuning[1]} if tracker.tuning else "default (Align on all SMs but 8)",
            "host_enqueue_ms_per_step": host_enqueue_ms / K, "clocks": cs.summary(),
            "single_gpu_same_workload": single,
            "speedup_vs_single_gpu_same_workload": (K / (ms / 1e3)) / single["value"] if single else None,
            "e2e": {"value": K / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(cfg.width * cfg.height * 2),
                    "d2h_bytes_per_step": 64 * world, "ms_per_step": ms_e2e / K},
        }
        if cpu:
            out["cpu_baseline"] = cpu
        emit(out)
    if multi:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference arm: the reference's own CUDA kernels (rebuilt for sm_100a from /root/reference by
    oracle/Makefile, UNMODIFIED) driven with the reference's own call sequence -- SDF_Hashtable::integrate
    (4 device syncs + 2 blocking D2H per frame) and CameraTracking::Align (5 syncs + 3 D2H per iteration,
    cuBLAS Sgemv/Ssyrk over the 7.4 MB Jacobian) -- on the frames and the table geometry of config C2.  Falls back to
    the CPU transliteration when no GPU / no prebuilt oracle/_ref is available."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    cfg, scene, traj, seq_len = workload_config("C2")
    n_unique = min(seq_len, max(K + W, 2))
    frames, poses = render_frames(cfg, scene, traj, n_unique)
    cpu = cpu_baselines(cfg, frames, poses[0])
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict("C2", cfg, n_unique),
            "policy": "the reference's as-built arithmetic (it has no other); its kernels take the table geometry, voxel size and truncation of the config"}
    if args.gpus > 1:
        base["note_multi_gpu"] = ("the reference is a one-GPU program with the image size baked in at 640x480 (SURVEY quirk Q15): it cannot run config C4 "
                                  "(1280x720, partitioned over GPUs), which is what the own arm runs at --gpus > 1.  This line is its C2 number on ONE GPU; "
                                  "dividing a C4 value at N GPUs by it compares different workloads")
    try:
        import torch

        from oracle import binding as ob

        if not torch.cuda.is_available():
            raise RuntimeError("no GPU")
        ob.ref_lib()
    except Exception as e:  # noqa: BLE001 -- any failure means: time the CPU port instead
        out = dict(base, value=cpu["value"], ms_per_step=1e3 / cpu["value"], cpu_baseline=cpu,
                   e2e={"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                   note=f"reference CUDA harness unavailable ({e}); value is the CPU transliteration")
        emit(out)
        return
    from voxelhashing_demo_b200.scenes import pingpong

    order = [pingpong(i, n_unique) for i in range(W + K)]
    n = 640 * 480

    def run_variant(variant: str, steps: int, warm: int):
        ref = ob.ref_lib(variant)
        with silence_stdout():
            torch.cuda.set_device(0)
            rc = ref.ref_init(cfg.numBuckets, cfg.bucketSize, cfg.numVoxelBlocks, cfg.voxelSize, cfg.truncation)
            assert rc == 0, f"ref_init -> {rc}"
            Kc, Kinv = cfg.K(), cfg.Kinv()
            ref.ref_set_intrinsic(Kc.ctypes.data, Kinv.ctypes.data)
            d_frames = torch.from_numpy(frames).cuda()
            maps = [(torch.zeros((n, 4), device="cuda"), torch.zeros((n, 4), device="cuda")) for _ in range(2)]
            solve = ob.ref_solve_callback()
            pose = poses[order[0]].astype(np.float64)
            est = np.zeros(6, np.float32)
            delta = np.eye(4, dtype=np.float32).reshape(16).copy()
            t0 = None
            for i, fi in enumerate(order[:warm + steps]):
                if i == warm:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                cur, prev = maps[i & 1], maps[1 - (i & 1)]
                ref.ref_preprocess(cur[0].data_ptr(), cur[1].data_ptr(), d_frames[fi].data_ptr())          # Application.cpp:73
                if i > 0:
                    ref.ref_align(cur[0].data_ptr(), prev[0].data_ptr(), prev[1].data_ptr(), 20, solve, est.ctypes.data, delta.ctypes.data)
                    pose = pose @ delta.reshape(4, 4).astype(np.float64)
                p32 = np.ascontiguousarray(pose.astype(np.float32).reshape(16))
                ref.ref_integrate(p32.ctypes.data, cur[0].data_ptr(), cur[1].data_ptr())                    # Application.cpp:84
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        return steps / dt, dt

    val, dt = run_variant("", K, W)
    variants = {}
    for v, label in (("noprintf", "device printf calls of insertVoxelEntry (VoxelUtils.cu:433-435, :452) deleted"),
                     ("G", "the shipped flags: -G device debug build (CMakeLists.txt:21)")):
        try:
            steps_v = K if v == "noprintf" else max(3, min(K, 6))
            vv, _ = run_variant(v, steps_v, min(W, 3))
            variants[v] = {"value": vv, "unit": UNIT, "steps": steps_v, "what": label}
        except Exception as e:  # noqa: BLE001
            variants[v] = {"unavailable": str(e)}
    out = dict(base, value=val, ms_per_step=1e3 * dt / K,
               cpu_baseline={"value": val, "unit": UNIT, "cores": 0, "kind": "reference",
                             "sample": f"{K} frames after {W} warm-up; the reference's OWN CUDA kernels rebuilt for sm_100a "
                                       "(-O3 -fmad=false, no -G), its own host syncs and device printf left in; it has no CPU path"},
               variants=variants,
               cpu_port=cpu,
               e2e={"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "C2", "C3", "C4", "C5"],
                    help="C2 (default at N=1), C3 (720p, 5 mm), C4 (2 mm large volume; default at N>1 and under torchrun), "
                         "C5 (C2 tracked frame-to-model: raycast in the loop)")
    ap.add_argument("--overlap", type=int, default=1, help="1: fuse frame k beside the tracking of frame k+1 (VH_PIPE_OVERLAP); 0: strictly serial frames")
    ap.add_argument("--ready", type=int, default=1, help="N=1 device-resident leg: 1 = vh_pipeline_push_device_ready (input complete, no producer on the stream), 0 = vh_pipeline_push_device")
    ap.add_argument("--repeats", type=int, default=5, help="N=1: how many times the W warm-up + K timed steps pass is repeated (median reported)")
    ap.add_argument("--align-ctas", type=int, default=0, help="partitioned runs: CTAs of the persistent Align kernel (0 = automatic)")
    ap.add_argument("--reserve-sms", type=int, default=0, help="partitioned runs: SMs the integrate grid leaves free for the Align grid")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-hbm", action="store_true", help="skip the large-volume integrate roofline leg")
    ap.add_argument("--no-refexact", action="store_true", help="skip the RefExact leg of the repo's own kernels")
    ap.add_argument("--no-single", action="store_true", help="multi-GPU: skip the 1-GPU run of the same workload")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
