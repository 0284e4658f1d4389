"""CPU suite, world_size 2 over gloo: the host-side logic of the multi-GPU path.

The kernels cannot run here (no GPU), so the per-rank computation is done by the CPU oracle with the
SAME partition rule and row split the CUDA path uses; what is tested is the plumbing around it:
row ranges, ownership, the 32-float all-reduce and the frame broadcast giving every rank the same
system / pose, and the union of the per-rank tables equalling the single-table result."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

torch = pytest.importorskip("torch")

from voxelhashing_demo_b200 import POLICY_FIXED  # noqa: E402
from voxelhashing_demo_b200.dist import owner_of, row_range  # noqa: E402


def test_row_range_partitions_the_image():
    for h in (480, 720, 7, 481):
        for world in (1, 2, 3, 4, 8):
            spans = [row_range(r, world, h) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == h
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_owner_matches_oracle_partition(oracle):
    from conftest import render, small_cfg
    from voxelhashing_demo_b200 import scenes

    base = dict(policy=POLICY_FIXED, numBuckets=100003, numVoxelBlocks=4096, truncation=0.06, overflowSlots=1024)
    depth = render(small_cfg(**base), scenes.scene_S1(), np.eye(4))
    for parts in (2, 3, 8):
        for rank in range(parts):
            t = oracle.OracleTable(small_cfg(partCount=parts, partRank=rank, **base))
            v, _, df = t.preprocess(depth)
            t.fuse_frame(np.eye(4), v, df)
            keys = [tuple(int(c) for c in e[:3]) for e in t.entries()]
            assert keys and all(owner_of(*k, parts) == rank for k in keys)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from conftest import render, small_cfg
    from oracle import binding as ob
    from voxelhashing_demo_b200 import scenes

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base = dict(policy=POLICY_FIXED, numBuckets=100003, numVoxelBlocks=4096, truncation=0.06, overflowSlots=1024,
                    width=160, height=120, fx=517.3 / 4, fy=516.5 / 4, cx=318.6 / 4, cy=255.3 / 4, icpNormalThres=0.8)
        cfg = small_cfg(partCount=world, partRank=rank, **base)
        table = ob.OracleTable(cfg)
        pose = scenes.trajectory_C2(0).astype(np.float32)
        prev = None
        est = np.zeros(6, np.float32)
        delta = np.eye(4, dtype=np.float32)
        r0, r1 = row_range(rank, world, cfg.height)
        for k in (0, 6, 12):
            # (1) frame broadcast from the ingest rank
            frame = torch.zeros(cfg.height * cfg.width * 2, dtype=torch.uint8)     # raw bytes: gloo has no int16
            if rank == 0:
                d = render(cfg, scenes.scene_S1T(), scenes.trajectory_C2(k))
                frame = torch.from_numpy(d.reshape(-1).view(np.uint8).copy())
            dist.broadcast(frame, src=0)
            depth = frame.numpy().view(np.uint16).reshape(cfg.height, cfg.width)
            v, n, df = table.preprocess(depth)
            # (2) ICP: rows split, ONE 32-float all-reduce per iteration, identical solve everywhere
            if prev is not None:
                for _ in range(5):
                    part = torch.from_numpy(ob.icp_system(cfg, v, n, prev[0], prev[1], delta, r0, r1))
                    dist.all_reduce(part)
                    ok, est, delta = ob.icp_solve(part.numpy(), est, delta)
                    assert ok
                pose = (pose.astype(np.float64) @ delta.astype(np.float64)).astype(np.float32)
            # (3) fusion of the owned blocks only
            table.fuse_frame(pose, v, df)
            prev = (v, n)
        np.savez(Path(out_dir) / f"rank{rank}.npz", pose=pose, keys=table.entries()[:, :3],
                 **{f"b{i}": table.block(*e[:3]) for i, e in enumerate(table.entries())})
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path, oracle):
    import torch.multiprocessing as mp

    from conftest import render, small_cfg
    from voxelhashing_demo_b200 import scenes

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    # pose bit-identical on both ranks (same all-reduced system, same solve)
    assert np.array_equal(ranks[0]["pose"].view(np.uint32), ranks[1]["pose"].view(np.uint32))

    # single-process run of the same sequence
    base = dict(policy=POLICY_FIXED, numBuckets=100003, numVoxelBlocks=4096, truncation=0.06, overflowSlots=1024,
                width=160, height=120, fx=517.3 / 4, fy=516.5 / 4, cx=318.6 / 4, cy=255.3 / 4, icpNormalThres=0.8)
    cfg = small_cfg(**base)
    table = oracle.OracleTable(cfg)
    pose = scenes.trajectory_C2(0).astype(np.float32)
    prev, est, delta = None, np.zeros(6, np.float32), np.eye(4, dtype=np.float32)
    for k in (0, 6, 12):
        v, n, df = table.preprocess(render(cfg, scenes.scene_S1T(), scenes.trajectory_C2(k)))
        if prev is not None:
            for _ in range(5):
                ok, est, delta = oracle.icp_solve(oracle.icp_system(cfg, v, n, prev[0], prev[1], delta), est, delta)
            pose = (pose.astype(np.float64) @ delta.astype(np.float64)).astype(np.float32)
        table.fuse_frame(pose, v, df)
        prev = (v, n)
    assert np.max(np.abs(ranks[0]["pose"] - pose)) < 1e-5            # row-split sums differ from the whole only by fp32 rounding
    truth = scenes.trajectory_C2(12)
    assert np.max(np.abs(pose[:3, 3] - truth[:3, 3])) < 0.01
    keys = [set(map(tuple, r["keys"].tolist())) for r in ranks]
    assert not keys[0] & keys[1]
    whole = {tuple(int(c) for c in e[:3]) for e in table.entries()}
    assert (keys[0] | keys[1]) == whole and min(len(keys[0]), len(keys[1])) > 0.3 * len(whole)


def _stream_worker(rank, world, port, out_dir):
    """Re-partitioning through the streaming calls: every rank streams its far blocks out, the records are
    all-gathered, and every rank streams in the whole pile -- the ownership test inside stream-in keeps only its own."""
    import torch.distributed as dist

    from conftest import render, small_cfg
    from oracle import binding as ob
    from voxelhashing_demo_b200 import scenes

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base = dict(policy=POLICY_FIXED, numBuckets=1009, numVoxelBlocks=4096, truncation=0.06, overflowSlots=1024)
        cfg = small_cfg(partCount=world, partRank=rank, **base)
        t = ob.OracleTable(cfg)
        pose = np.eye(4, dtype=np.float32)
        v, _, df = t.preprocess(render(cfg, scenes.scene_S1(), pose))
        t.fuse_frame(pose, v, df)
        before = t.block_dict()
        ent, vox = t.stream_out((0.0, 0.0, 0.0), 2.2, 4096)
        # exchange: padded to a common length so a plain all_gather of tensors does it
        n = torch.tensor([len(ent)])
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n)
        cap = int(max(c.item() for c in counts))
        pe = torch.zeros((cap, 5), dtype=torch.int32)
        pv = torch.zeros((cap, 512, 2), dtype=torch.float32)
        pe[: len(ent)] = torch.from_numpy(ent)
        pv[: len(ent)] = torch.from_numpy(vox)
        ge = [torch.zeros_like(pe) for _ in range(world)]
        gv = [torch.zeros_like(pv) for _ in range(world)]
        dist.all_gather(ge, pe)
        dist.all_gather(gv, pv)
        accepted = 0
        for r in range(world):
            k = int(counts[r].item())
            accepted += t.stream_in(ge[r][:k].numpy(), gv[r][:k].numpy())
        after = t.block_dict()
        same = set(after) == set(before) and all(np.array_equal(after[key], before[key]) for key in after)
        np.savez(Path(out_dir) / f"s{rank}.npz", moved=len(ent), accepted=accepted, same=int(same), nblocks=len(after),
                 offered=int(sum(c.item() for c in counts)))
    finally:
        dist.destroy_process_group()


def test_streaming_respects_the_partition(tmp_path, oracle):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_stream_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ranks = [np.load(tmp_path / f"s{r}.npz") for r in range(2)]
    for r in ranks:
        # offered everyone's records, each rank takes back exactly what it streamed out and ends where it started
        assert int(r["moved"]) > 0 and int(r["offered"]) > int(r["moved"])
        assert int(r["accepted"]) == int(r["moved"]) and int(r["same"]) == 1


def test_auto_tuning_splits_the_sms_only_when_the_align_is_latency_bound():
    """PartitionedTracker._auto_tuning: the SM split between the Align grid and the integrate grid (measured on C4 at 2 mm:
    a gain at 4 and 8 ranks, a loss at 2)."""
    from voxelhashing_demo_b200.dist import PartitionedTracker

    class Cfg:
        width, height = 1280, 720

    assert PartitionedTracker._auto_tuning(Cfg, 1) is None and PartitionedTracker._auto_tuning(Cfg, 2) is None
    for world in (4, 8):
        ctas, reserved = PartitionedTracker._auto_tuning(Cfg, world)
        assert ctas == reserved and 16 <= ctas <= 40


def test_bench_arms_describe_the_same_config():
    """The own arm and the reference arm of bench.py must print the identical `config` object (the driver compares them)."""
    import importlib.util
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    spec = importlib.util.spec_from_file_location("bench_for_test", root / "bench.py")
    saved_fd1 = os.dup(1)                                   # bench.py points fd 1 at stderr on import: undo it afterwards
    try:
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
    finally:
        os.dup2(saved_fd1, 1)
        os.close(saved_fd1)
    from voxelhashing_demo_b200 import POLICY_REF_EXACT

    cfg_own, _, _, _ = bench.workload_config("C2")
    cfg_ref, _, _, _ = bench.workload_config("C2", policy=POLICY_REF_EXACT)
    assert bench.config_dict("C2", cfg_own, 25) == bench.config_dict("C2", cfg_ref, 25)
    c4 = bench.workload_config("C4", 8, 3)[0]
    assert c4.voxelSize == 0.002 and c4.numVoxelBlocks == 1048576 and c4.numBuckets == 4000037 and (c4.width, c4.height) == (1280, 720)
    assert "2 mm" in bench.config_dict("C4", c4, 25)["workload"]
    sys.modules.pop("bench_for_test", None)
