"""Pins the CPU oracle (and the new CUDA path) against OUTPUTS OF THE REFERENCE ITSELF.

The reference ships no tests or golden vectors.  Its own CUDA sources do compile
for sm_100a, so oracle/_ref/libvh_ref.so (built from /root/reference by oracle/Makefile, never copied)
is run here next to the oracle and libvh_b200.so on identical synthetic frames.  The pinned quantities
are the ones the reference computes deterministically.
"""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import entries_to_set, rot_err

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from voxelhashing_demo_b200 import Config, Context, scenes  # noqa: E402

HERE = Path(__file__).resolve().parent


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def run_reference(tmp_path, num_buckets, ks, align=False, single=False):
    from oracle import binding as ob

    if not ob.REF_LIB.exists():
        pytest.skip("oracle/_ref/libvh_ref.so not built (needs /root/reference at build time)")
    out = tmp_path / f"ref_{num_buckets}.npz"
    cmd = [sys.executable, str(HERE / "ref_pin_worker.py"), str(out), str(num_buckets), ",".join(map(str, ks))]
    if align:
        cmd.append("align")
    if single:
        cmd.append("single")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    return np.load(out)


def test_reference_fusion_pins_oracle_and_cuda(built_library, oracle, tmp_path):
    """Reference defaults over three frames of a moving camera, allocation run to convergence each
    frame (the reference inserts one block per bucket per pass and the winner is racy, quirk Q4): the
    reference's table, compact list and voxels equal the oracle's and the new kernels' -- sets exact,
    weights exact, sdf bit-exact (bar: 1e-5)."""
    ks = [0, 10, 20]
    ref = run_reference(tmp_path, 5000, ks)
    cfg = Config(numVoxelBlocks=4000)
    ot = oracle.OracleTable(cfg)
    ctx = Context(cfg)
    for i, k in enumerate(ks):
        pose = scenes.trajectory_C2(k).astype(np.float32)
        depth = ref[f"depth{i}"]
        ov, on, _ = ot.preprocess(depth)
        # preProcess of the reference == oracle, bit for bit
        assert np.array_equal(bits(ref[f"verts{i}"]), bits(ov))
        assert np.array_equal(bits(ref[f"normals{i}"]), bits(on))
        for _ in range(4):
            rep = ot.alloc(pose, ov)
        assert rep.requestedNew == 0
        nvis = ot.compact(pose)
        ot.integrate(pose, ov)
        v, n = torch.from_numpy(ov).cuda(), torch.from_numpy(on).cuda()
        for _ in range(4):
            ctx.set_pose(pose)
            ctx.alloc_blocks(v, n)
        ctx.compact()
        ctx.integrate(v)
        occ = ref[f"occupied{i}"]
        assert int(occ[0]) == nvis == ctx.stats().numVisible
        assert int(occ[2]) == ot.heap_counter() == ctx.stats().heapCounter
        assert entries_to_set(ref[f"table{i}"]) == entries_to_set(ot.entries()) == entries_to_set(ctx.export_entries())
        assert entries_to_set(ref[f"compact{i}"]) == entries_to_set(ot.compact_entries()) == entries_to_set(ctx.export_compact())
    table, blocks = ref[f"table{len(ks) - 1}"], ref["blocks"]
    cpu, gpu = ot.block_dict(), ctx.block_dict()
    worst = 0.0
    for e, b in zip(table, blocks):
        key = (int(e[0]), int(e[1]), int(e[2]))
        assert np.array_equal(bits(b[:, 1]), bits(cpu[key][:, 1])) and np.array_equal(bits(b[:, 1]), bits(gpu[key][:, 1]))
        worst = max(worst, float(np.max(np.abs(b[:, 0] - cpu[key][:, 0]))), float(np.max(np.abs(b[:, 0] - gpu[key][:, 0]))))
        assert np.array_equal(bits(b[:, 0]), bits(cpu[key][:, 0])), f"oracle sdf differs from the reference in block {key}"
        assert np.array_equal(bits(b[:, 0]), bits(gpu[key][:, 0])), f"kernel sdf differs from the reference in block {key}"
    assert worst <= 1e-5


def test_reference_c1_default_relaxed(built_library, oracle, tmp_path):
    """C1 exactly as shipped (5000 buckets, identity pose): the reference races on 34 buckets (Q4), so
    the relaxed rule applies: one new block per touched bucket, equality on un-contended buckets."""
    ref = run_reference(tmp_path, 5000, [0], single=True)
    cfg = Config(numVoxelBlocks=4000)
    ot = oracle.OracleTable(cfg)
    ov, _, _ = ot.preprocess(ref["depth0"])
    rep = ot.alloc(np.eye(4, dtype=np.float32), ov)
    requested = ot.last_requested_new()
    got = entries_to_set(ref["table0"])
    assert len(got) == rep.inserted == 199 and got <= requested
    req_by_bucket = {}
    for k in requested:
        req_by_bucket.setdefault(oracle.hash_block(cfg, *k), set()).add(k)
    got_by_bucket = {}
    for k in got:
        got_by_bucket.setdefault(oracle.hash_block(cfg, *k), set()).add(k)
    assert set(got_by_bucket) == set(req_by_bucket)
    for b, req in req_by_bucket.items():
        assert len(got_by_bucket[b]) == 1
        if len(req) == 1:
            assert got_by_bucket[b] == req
    assert int(ref["occupied0"][0]) == 199          # all of them pass the (quirky) frustum test again


def test_reference_icp_pins_oracle_and_cuda(built_library, oracle, tmp_path):
    """FindCorrespondences / Jacobian kernel / cuBLAS normal equations of the reference vs oracle and
    the fused kernel; then the 20-iteration Align (reference device half + closed-form solve)."""
    ref = run_reference(tmp_path, 5000, [0, 12], align=True)
    cfg = Config(numVoxelBlocks=4000)
    ot = oracle.OracleTable(cfg)
    tv, tn, _ = ot.preprocess(ref["depth0"])
    iv, inn, _ = ot.preprocess(ref["depth1"])
    ident = np.eye(4, dtype=np.float32)
    oerr, ocorr, ocorrN, ores = oracle.find_correspondences(cfg, iv, None, tv, tn, ident)
    assert np.array_equal(bits(ref["icp_corr"]), bits(ocorr))
    assert np.array_equal(bits(ref["icp_corrN"]), bits(ocorrN))
    assert np.array_equal(bits(ref["icp_res"]), bits(ores))
    assert abs(float(ref["icp_err"][0]) - float(np.sum(ores.astype(np.float64)))) <= 1e-5 * float(np.sum(np.abs(ores)))
    assert np.array_equal(bits(ref["icp_jac"]), bits(oracle.jacobians(cfg, ocorr, ocorrN)))
    osys = oracle.icp_system(cfg, iv, None, tv, tn, ident)
    JtJ = ref["icp_JtJ"].reshape(6, 6)              # column-major, lower triangle valid
    ref_upper = np.array([JtJ[i, j] for i in range(6) for j in range(i, 6)])   # C(j,i) lower == [i*6+j]
    scale = float(np.max(np.abs(osys[:21])))
    # cublasSsyrk accumulates 307200 fp32 products per entry: its own error is ~1e-4 of the scale (measured 8.3e-5 on B200)
    assert np.max(np.abs(ref_upper - osys[:21])) <= 3e-4 * scale
    assert np.max(np.abs(ref["icp_Jtr"] - osys[21:27])) <= 1e-5 * max(1.0, float(np.max(np.abs(osys[21:27])))) + 1e-8 * scale
    # fused CUDA reduction against the reference's cuBLAS result
    ctx = Context(cfg)
    g = [torch.from_numpy(a).cuda() for a in (iv, inn, tv, tn)]
    dsys = torch.zeros(32, device="cuda")
    ctx.icp_reset(True)
    ctx.icp_reduce(g[0], g[1], g[2], g[3], 0, 480, dsys)
    torch.cuda.synchronize()
    gsys = dsys.cpu().numpy()
    assert np.max(np.abs(gsys[:21] - ref_upper)) <= 3e-4 * scale               # bound by cuBLAS, see above
    assert np.max(np.abs(gsys[:21] - osys[:21])) <= 1e-5 * scale               # the fused reduction itself meets 1e-5
    assert np.max(np.abs(gsys[21:27] - ref["icp_Jtr"])) <= 1e-5 * max(1.0, float(np.max(np.abs(ref["icp_Jtr"])))) + 1e-8 * scale
    # Align: pose within 1e-4 of the reference loop
    assert int(ref["align_iters"][0]) == 20
    ctx.icp_reset(True)
    ctx.icp_align(g[0], g[1], g[2], g[3], 20)
    gdelta = ctx.icp_get()[0]
    its, oest, odelta = oracle.icp_align(cfg, iv, None, tv, tn, 20)
    for d in (gdelta, odelta):
        assert rot_err(d[:3, :3], ref["align_delta"][:3, :3]) <= 1e-4
        assert np.max(np.abs(d[:3, 3] - ref["align_delta"][:3, 3])) <= 1e-4
