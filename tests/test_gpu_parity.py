"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): allocated block sets bit-exact as sets; weights bit-exact;
TSDF within 1e-5 absolute (bit-exact is asserted where the arithmetic is deterministic);
ICP normal equations relative 1e-5; pose within 1e-4 in rotation and translation.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np
import pytest

from conftest import entries_to_set, render, rot_err, small_cfg

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

torch = pytest.importorskip("torch")

from voxelhashing_demo_b200 import POLICY_FIXED, POLICY_REF_EXACT, Config, Context, FramePipeline, scenes  # noqa: E402
from voxelhashing_demo_b200 import lib as L  # noqa: E402


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def gpu_preprocess(ctx, depth):
    d = cu(depth.reshape(-1))
    v, n, df = ctx.new_maps()
    ctx.preprocess(d, v, n, df)
    torch.cuda.synchronize()
    return v, n, df


def fixed_cfg(**kw):
    base = dict(policy=POLICY_FIXED, numBuckets=100003, numVoxelBlocks=8192, truncation=0.06, truncScale=0.01,
                depthMin=0.1, depthMax=4.0, overflowSlots=4096)
    base.update(kw)
    return Config(**base)


def ref_fuse_gpu(ctx, pose, v, n, passes=4):
    """RefExact inserts at most one block per bucket per frame and the winner is racy (quirk Q4; the block
    hash has genuine 32-bit collisions, so contention exists at ANY bucket count).  Repeating the
    allocation pass until every requested block is in makes the table -- and therefore every voxel --
    deterministic; then one compaction and one integration."""
    for _ in range(passes):
        ctx.set_pose(pose)
        ctx.alloc_blocks(v, n)
    ctx.compact()
    ctx.integrate(v)


def ref_fuse_cpu(ot, pose, ov, passes=4):
    inserted = 0
    for _ in range(passes):
        rep = ot.alloc(pose, ov)
        inserted += rep.inserted
    assert rep.requestedNew == 0, "allocation did not converge"
    return inserted, ot.compact(pose), ot.integrate(pose, ov)


def compare_blocks(gpu_blocks: dict, cpu_blocks: dict, sdf_tol=1e-5):
    assert set(gpu_blocks) == set(cpu_blocks)
    worst = 0.0
    exact = True
    for k, g in gpu_blocks.items():
        c = cpu_blocks[k]
        assert np.array_equal(bits(g[:, 1]), bits(c[:, 1])), f"weights differ in block {k}"
        d = float(np.max(np.abs(g[:, 0].astype(np.float64) - c[:, 0].astype(np.float64))))
        worst = max(worst, d)
        exact = exact and np.array_equal(bits(g[:, 0]), bits(c[:, 0]))
    assert worst <= sdf_tol, f"max |sdf_gpu - sdf_oracle| = {worst}"
    return exact, worst


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("policy", [POLICY_REF_EXACT, POLICY_FIXED])
@pytest.mark.parametrize("size", ["vga", "small", "ragged"])
def test_preprocess_bit_exact(built_library, oracle, policy, size):
    if size == "vga":
        cfg = Config(policy=policy)
    elif size == "small":
        cfg = small_cfg(policy=policy)
    else:  # width/height not multiples of the 32x8 tile
        cfg = small_cfg(policy=policy, width=161, height=123)
    depth = render(cfg, scenes.scene_S1(), scenes.trajectory_C2(7))
    depth[5:9, 10:40] = 0                 # holes
    depth[0, 0] = 65535                   # out of sensor range in Fixed
    ctx = Context(cfg)
    v, n, df = gpu_preprocess(ctx, depth)
    ov, on, odf = oracle.OracleTable(cfg).preprocess(depth)
    assert np.array_equal(bits(v.cpu().numpy()), bits(ov))
    assert np.array_equal(bits(n.cpu().numpy()), bits(on))
    assert np.array_equal(bits(df.cpu().numpy()), bits(odf))


def test_refexact_c1_exact(built_library, oracle):
    """C1 as shipped (reference defaults, identity pose), allocation run to convergence."""
    cfg = Config()
    depth = render(cfg, scenes.scene_S1(), np.eye(4))
    pose = np.eye(4, dtype=np.float32)
    ot = oracle.OracleTable(cfg)
    ov, on, _ = ot.preprocess(depth)
    inserted, nvis, nupd = ref_fuse_cpu(ot, pose, ov)
    assert inserted == 234                                       # SURVEY.md Appendix B probe

    ctx = Context(cfg)
    v, n, _ = gpu_preprocess(ctx, depth)
    ref_fuse_gpu(ctx, pose, v, n)
    st = ctx.stats()
    assert entries_to_set(ctx.export_entries()) == entries_to_set(ot.entries())
    assert st.numVisible == nvis and st.numAllocated == inserted
    assert entries_to_set(ctx.export_compact()) == entries_to_set(ot.compact_entries())
    assert int(st.numUpdated) == nupd
    exact, worst = compare_blocks(ctx.block_dict(), ot.block_dict())
    assert exact, f"sdf not bit-exact (worst {worst})"
    # second and third integration of the same frame: weights follow the same fp32 +0.1f chain
    for _ in range(2):
        ref_fuse_cpu(ot, pose, ov, passes=1)
        ref_fuse_gpu(ctx, pose, v, n, passes=1)
    exact, _ = compare_blocks(ctx.block_dict(), ot.block_dict())
    assert exact


def test_refexact_c1_contended_relaxed_rule(built_library, oracle):
    """C1 as shipped (5000 buckets): 34 contended buckets; the reference is racy there (quirk Q4).
    Relaxed rule of SURVEY.md section 8c."""
    cfg = Config()
    depth = render(cfg, scenes.scene_S1(), np.eye(4))
    pose = np.eye(4, dtype=np.float32)
    ot = oracle.OracleTable(cfg)
    ov, _, _ = ot.preprocess(depth)
    rep = ot.alloc(pose, ov)
    assert rep.bucketsContended == 34 and rep.bucketsTouched == 199
    requested_new = ot.last_requested_new()

    ctx = Context(cfg)
    v, n, _ = gpu_preprocess(ctx, depth)
    ctx.set_pose(pose)
    ctx.alloc_blocks(v, n)
    got = entries_to_set(ctx.export_entries())
    assert got <= requested_new                                   # allocated subset of requested
    by_bucket_req, by_bucket_got = {}, {}
    for k in requested_new:
        by_bucket_req.setdefault(oracle.hash_block(cfg, *k), set()).add(k)
    for k in got:
        by_bucket_got.setdefault(oracle.hash_block(cfg, *k), set()).add(k)
    for b, req in by_bucket_req.items():
        assert len(by_bucket_got.get(b, ())) == 1                  # exactly one new block per touched bucket
        if len(req) == 1:
            assert by_bucket_got[b] == req                         # equality on un-contended buckets
    assert ctx.stats().numAllocated == rep.inserted == 199
    # repeating the frame max-multiplicity times converges to the full requested set
    for _ in range(rep.maxNewPerBucket):
        ot.alloc(pose, ov)
        ctx.set_pose(pose)
        ctx.alloc_blocks(v, n)
    assert entries_to_set(ctx.export_entries()) == entries_to_set(ot.entries()) == requested_new


def test_refexact_moving_camera_sequence(built_library, oracle):
    """Non-identity poses under RefExact (geometry is wrong by design, Q1/Q2/Q9, but deterministic)."""
    cfg = Config(numBuckets=100003, numVoxelBlocks=4000)
    ot = oracle.OracleTable(cfg)
    ctx = Context(cfg)
    for k in (0, 10, 20):
        pose = scenes.trajectory_C2(k).astype(np.float32)
        depth = render(cfg, scenes.scene_S1(), pose)
        ov, _, _ = ot.preprocess(depth)
        inserted, nvis, nupd = ref_fuse_cpu(ot, pose, ov)
        v, n, _ = gpu_preprocess(ctx, depth)
        ref_fuse_gpu(ctx, pose, v, n)
        st = ctx.stats()
        assert (st.numVisible, int(st.numUpdated)) == (nvis, nupd)
        assert entries_to_set(ctx.export_entries()) == entries_to_set(ot.entries())
    exact, _ = compare_blocks(ctx.block_dict(), ot.block_dict())
    assert exact


@pytest.mark.parametrize("dense", [True, False])
def test_fixed_sequence_bit_exact(built_library, oracle, dense):
    cfg = fixed_cfg()
    ot = oracle.OracleTable(cfg)
    ctx = Context(cfg)
    for k in (0, 5, 10, 15):
        pose = scenes.trajectory_C2(k).astype(np.float32)
        depth = render(cfg, scenes.scene_S1(), pose)
        ov, _, odf = ot.preprocess(depth)
        rep, nvis, nupd = ot.fuse_frame(pose, ov, odf if dense else None)
        assert rep.dropped == 0
        v, n, df = gpu_preprocess(ctx, depth)
        ctx.fuse_frame(pose, v, n, df if dense else None)
        st = ctx.stats()
        assert entries_to_set(ctx.export_entries()) == entries_to_set(ot.entries())
        assert (st.numVisible, int(st.numUpdated), st.lastInserted, st.dropped) == (nvis, nupd, rep.inserted, 0)
    exact, worst = compare_blocks(ctx.block_dict(), ot.block_dict())
    assert exact, f"Fixed sdf not bit-exact (worst {worst})"


@pytest.mark.parametrize("policy", [POLICY_REF_EXACT, POLICY_FIXED])
@pytest.mark.parametrize("size", ["vga", "ragged"])
def test_alloc_from_raw_depth_equals_alloc_from_vertex_map(built_library, oracle, policy, size):
    """k_alloc fed with the u16 depth image (2 B/px, back-projection in registers) requests exactly the
    blocks it requests from the pre-processed vertex map (16 B/px) -- and both equal the oracle's table."""
    kw = dict(policy=policy, numBuckets=100003, numVoxelBlocks=8192)
    if policy == POLICY_FIXED:
        kw.update(truncation=0.06, overflowSlots=4096)
    cfg = Config(**kw) if size == "vga" else small_cfg(width=161, height=123, **kw)
    pose = scenes.trajectory_C2(9).astype(np.float32)
    depth = render(cfg, scenes.scene_S1(), pose)
    depth[7:19, 30:90] = 0
    depth[0, 0] = 65535
    a, b = Context(cfg), Context(cfg)
    v, n, df = gpu_preprocess(a, depth)
    d16 = cu(depth.reshape(-1))
    for _ in range(4 if policy == POLICY_REF_EXACT else 1):       # RefExact inserts one block per bucket per pass (Q4)
        a.set_pose(pose)
        a.alloc_blocks(v, n)
        b.set_pose(pose)
        b.alloc_blocks_depth(d16)
    torch.cuda.synchronize()
    sa, sb = entries_to_set(a.export_entries()), entries_to_set(b.export_entries())
    assert sa == sb and len(sa) > 100
    ot = oracle.OracleTable(cfg)
    ov, _, _ = ot.preprocess(depth)
    for _ in range(4 if policy == POLICY_REF_EXACT else 1):
        ot.alloc(pose, ov)
    assert sb == entries_to_set(ot.entries())


def test_fixed_overflow_chain(built_library, oracle):
    """64 buckets x 2 slots force most blocks into the overflow arena; set equality must survive."""
    cfg = fixed_cfg(numBuckets=64, bucketSize=2, attachedLinkedListSize=64, overflowSlots=4096, numVoxelBlocks=4096,
                    width=160, height=120, fx=517.3 / 4, fy=516.5 / 4, cx=318.6 / 4, cy=255.3 / 4)
    ot = oracle.OracleTable(cfg)
    ctx = Context(cfg)
    for k in (0, 8):
        pose = scenes.trajectory_C2(k).astype(np.float32)
        depth = render(cfg, scenes.scene_S1(), pose)
        ov, _, odf = ot.preprocess(depth)
        rep, nvis, nupd = ot.fuse_frame(pose, ov, odf)
        assert rep.dropped == 0
        v, n, df = gpu_preprocess(ctx, depth)
        ctx.fuse_frame(pose, v, n, df)
        st = ctx.stats()
        assert st.overflowUsed > 0 and st.dropped == 0
        assert entries_to_set(ctx.export_entries()) == entries_to_set(ot.entries())
        assert (st.numVisible, int(st.numUpdated)) == (nvis, nupd)
    # chain integrity: every entry reachable from its bucket head through the offsets
    ent = ctx.export_entries()
    assert len(entries_to_set(ent)) == len(ent), "duplicate keys in the table"
    exact, _ = compare_blocks(ctx.block_dict(), ot.block_dict())
    assert exact


def test_fixed_capacity_exhaustion_is_graceful(built_library, oracle):
    """Heap smaller than the request: no crash, no duplicates, counters consistent (the reference reads
    out of bounds here, quirk Q6)."""
    cfg = fixed_cfg(numVoxelBlocks=50)
    ot = oracle.OracleTable(cfg)
    depth = render(cfg, scenes.scene_S1(), np.eye(4))
    ov, _, _ = ot.preprocess(depth)
    ctx = Context(cfg)
    v, n, df = gpu_preprocess(ctx, depth)
    ctx.fuse_frame(np.eye(4, dtype=np.float32), v, n, df)
    st = ctx.stats()
    ent = ctx.export_entries()
    ent = ent[ent["ptr"] >= 0]
    assert st.numAllocated == 50 and len(ent) == 50 and st.dropped > 0
    assert len(entries_to_set(ent)) == 50
    assert sorted(int(p) // 512 for p in ent["ptr"]) == list(range(50))


@pytest.mark.parametrize("policy", [POLICY_REF_EXACT, POLICY_FIXED])
def test_icp_system_and_split_kernels(built_library, oracle, policy):
    cfg = Config(policy=policy, depthMax=4.0, icpNormalThres=0.8 if policy == POLICY_FIXED else -1.0)
    ot = oracle.OracleTable(cfg)
    d0 = render(cfg, scenes.scene_S1T(), scenes.trajectory_C2(0))
    d1 = render(cfg, scenes.scene_S1T(), scenes.trajectory_C2(20))
    tv, tn, _ = ot.preprocess(d0)      # target = frame 0
    iv, inn, _ = ot.preprocess(d1)     # input = frame 20
    delta = oracle.se3_exp([0.01, -0.004, 0.002, 0.003, -0.002, 0.001])
    ctx = Context(cfg)
    g_tv, g_tn, g_iv, g_in = cu(tv), cu(tn), cu(iv), cu(inn)

    # split form: correspondences + Jacobians, bit-exact
    n = cfg.width * cfg.height
    corr, corrN = torch.zeros((n, 4), device="cuda"), torch.zeros((n, 4), device="cuda")
    res, err, J = torch.zeros(n, device="cuda"), torch.zeros(1, device="cuda"), torch.zeros((n, 6), device="cuda")
    ctx.find_correspondences(g_iv, g_in, g_tv, g_tn, delta, corr, corrN, res, err)
    ctx.jacobians(corr, corrN, J)
    torch.cuda.synchronize()
    oerr, ocorr, ocorrN, ores = oracle.find_correspondences(cfg, iv, inn, tv, tn, delta)
    assert np.array_equal(bits(corr.cpu().numpy()), bits(ocorr))
    assert np.array_equal(bits(corrN.cpu().numpy()), bits(ocorrN))
    assert np.array_equal(bits(res.cpu().numpy()), bits(ores))
    assert abs(float(err.item()) - float(np.sum(ores.astype(np.float64)))) <= 1e-5 * max(1.0, float(np.sum(np.abs(ores))))
    assert np.array_equal(bits(J.cpu().numpy()), bits(oracle.jacobians(cfg, ocorr, ocorrN)))

    # fused reduction vs fp64 oracle sums, relative 1e-5 of the system's scale
    ctx.icp_set_twist(oracle.se3_log(delta))
    dsys = torch.zeros(32, device="cuda")
    ctx.icp_reduce(g_iv, g_in, g_tv, g_tn, 0, cfg.height, dsys)
    torch.cuda.synchronize()
    gsys = dsys.cpu().numpy()
    gdelta = ctx.icp_get()[0]
    osys = oracle.icp_system(cfg, iv, inn, tv, tn, gdelta)
    assert gsys[28] == osys[28] and gsys[28] > 50000            # same correspondence count
    scale = float(np.max(np.abs(osys[:21])))
    assert np.max(np.abs(gsys[:21] - osys[:21])) <= 1e-5 * scale
    assert np.max(np.abs(gsys[21:27] - osys[21:27])) <= 1e-5 * max(1.0, float(np.max(np.abs(osys[21:27])))) + 1e-5 * scale * 1e-3
    assert abs(gsys[27] - osys[27]) <= 1e-5 * max(1.0, abs(float(osys[27])))
    # rows split in two halves (the multi-GPU decomposition) sums to the whole
    a, b = torch.zeros(32, device="cuda"), torch.zeros(32, device="cuda")
    ctx.icp_reduce(g_iv, g_in, g_tv, g_tn, 0, 200, a)
    ctx.icp_reduce(g_iv, g_in, g_tv, g_tn, 200, cfg.height, b)
    torch.cuda.synchronize()
    assert np.allclose((a + b).cpu().numpy()[:29], gsys[:29], rtol=2e-5, atol=1e-3)
    # reduce from stored correspondences (what Solver::BuildLinearSystem gets) agrees with the fused form
    if policy == POLICY_REF_EXACT:
        ctx.find_correspondences(g_iv, g_in, g_tv, g_tn, gdelta, corr, corrN, res, err)
        c = torch.zeros(32, device="cuda")
        ctx.icp_reduce_corr(corr, corrN, res, c)
        torch.cuda.synchronize()
        assert np.allclose(c.cpu().numpy()[:28], gsys[:28], rtol=2e-5, atol=1e-3)


@pytest.mark.parametrize("policy", [POLICY_REF_EXACT, POLICY_FIXED])
def test_icp_align_pose(built_library, oracle, policy):
    """20 Gauss-Newton iterations on the device vs the oracle loop: pose within 1e-4."""
    cfg = Config(policy=policy)
    ot = oracle.OracleTable(cfg)
    T0, T1 = scenes.trajectory_C2(0), scenes.trajectory_C2(12)
    tv, tn, _ = ot.preprocess(render(cfg, scenes.scene_S1T(), T0))
    iv, inn, _ = ot.preprocess(render(cfg, scenes.scene_S1T(), T1))
    ctx = Context(cfg)
    ctx.icp_reset(True)
    ctx.icp_align(cu(iv), cu(inn), cu(tv), cu(tn), 20)
    gdelta, gtwist, _ = ctx.icp_get()
    its, oest, odelta = oracle.icp_align(cfg, iv, inn, tv, tn, 20)
    assert its == 20
    assert rot_err(gdelta[:3, :3], odelta[:3, :3]) <= 1e-4
    assert np.max(np.abs(gdelta[:3, 3] - odelta[:3, 3])) <= 1e-4
    assert np.max(np.abs(gtwist - oest)) <= 1e-4
    if policy == POLICY_FIXED:       # the corrected pipeline must actually recover the motion
        truth = np.linalg.inv(T0) @ T1
        assert rot_err(gdelta[:3, :3], truth[:3, :3]) <= 2e-3
        assert np.max(np.abs(gdelta[:3, 3] - truth[:3, 3])) <= 5e-3
    # estimate accumulates across Align calls (quirk Q24): a second call starts from the first result
    ctx.icp_align(cu(iv), cu(inn), cu(tv), cu(tn), 5)
    its2, oest2, odelta2 = oracle.icp_align(cfg, iv, inn, tv, tn, 5, oest)
    g2 = ctx.icp_get()[0]
    assert rot_err(g2[:3, :3], odelta2[:3, :3]) <= 1e-4 and np.max(np.abs(g2[:3, 3] - odelta2[:3, 3])) <= 1e-4


def test_track_frame_is_preprocess_plus_align_plus_pose_chain(built_library, oracle):
    """vh_track_frame (one persistent launch: pre-processing prologue, Align, pose chain) == vh_preprocess + vh_icp_align +
    pose * delta, bit for bit, for both policies; the maps it leaves behind are the oracle's."""
    for policy in (POLICY_REF_EXACT, POLICY_FIXED):
        cfg = Config(policy=policy, icpNormalThres=0.8 if policy == POLICY_FIXED else -1.0)
        d0 = render(cfg, scenes.scene_S1T(), scenes.trajectory_C2(0))
        d1 = render(cfg, scenes.scene_S1T(), scenes.trajectory_C2(10))
        d1[100:130, 200:260] = 0
        pose0 = scenes.trajectory_C2(0).astype(np.float32)
        a, b = Context(cfg), Context(cfg)
        tv, tn, _ = gpu_preprocess(a, d0)
        # separate launches
        v1, n1, f1 = gpu_preprocess(a, d1)
        a.icp_reset(True)
        a.icp_align(v1, n1, tv, tn, 12)
        delta_a = a.icp_get()[0]
        # fused
        v2, n2, f2 = b.new_maps()
        p_in, p_out = cu(pose0.reshape(-1)), torch.zeros(16, device="cuda")
        b.icp_reset(True)
        b.track_frame(cu(d1.reshape(-1)), v2, n2, f2, tv, tn, 12, p_in, p_out)
        torch.cuda.synchronize()
        delta_b = b.icp_get()[0]
        for x, y in ((v1, v2), (n1, n2), (f1, f2)):
            assert np.array_equal(bits(x.cpu().numpy()), bits(y.cpu().numpy()))
        ov, on, odf = oracle.OracleTable(cfg).preprocess(d1)
        assert np.array_equal(bits(v2.cpu().numpy()), bits(ov)) and np.array_equal(bits(n2.cpu().numpy()), bits(on))
        assert np.array_equal(bits(delta_a), bits(delta_b))
        want = np.zeros((4, 4), np.float32)
        for r in range(4):                                  # k_set_frame's products, left to right, no contraction
            for c in range(4):
                t = np.float32(pose0[r, 0] * delta_b[0, c])
                for k in (1, 2, 3):
                    t = np.float32(t + np.float32(pose0[r, k] * delta_b[k, c]))
                want[r, c] = t
        assert np.array_equal(bits(p_out.cpu().numpy().reshape(4, 4)), bits(want))
        # in place (pose_out aliases pose_in)
        b.icp_reset(True)
        b.track_frame(cu(d1.reshape(-1)), v2, n2, f2, tv, tn, 12, p_in, p_in)
        torch.cuda.synchronize()
        assert np.array_equal(bits(p_in.cpu().numpy().reshape(4, 4)), bits(want))


def test_host_classes_solver_and_camera_tracking(built_library, oracle, tmp_path):
    """Solver::BuildLinearSystem and CameraTracking::Align are EXECUTED through the C++ classes (a test
    harness written like the reference's own host loop, tests/cpp/host_classes_driver.cpp) and pinned to the oracle."""
    import subprocess

    exe = tmp_path / "host_classes_driver"
    lib_dir = Path(built_library).parent
    r = subprocess.run(["g++", "-std=c++17", "-O1", str(ROOT / "tests/cpp/host_classes_driver.cpp"), "-I", str(ROOT / "include"),
                        "-I", "/usr/local/cuda/include", "-L", str(lib_dir), "-lvh_b200", "-L", "/usr/local/cuda/lib64", "-lcudart",
                        f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cfg = Config(policy=POLICY_REF_EXACT)
    T0, T1 = scenes.trajectory_C2(0), scenes.trajectory_C2(12)
    d0, d1 = render(cfg, scenes.scene_S1T(), T0), render(cfg, scenes.scene_S1T(), T1)
    np.concatenate([d0.reshape(-1), d1.reshape(-1)]).astype(np.uint16).tofile(tmp_path / "in.bin")
    r = subprocess.run([str(exe), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = np.fromfile(tmp_path / "out.bin", dtype=np.float32)
    facade, loop, ldlt = (out[k * 16:(k + 1) * 16].reshape(4, 4) for k in range(3))
    back, JTJ, JTr = out[48:54], out[54:90].reshape(6, 6).astype(np.float64), out[90:96].astype(np.float64)
    ot = oracle.OracleTable(cfg)
    tv, tn, _ = ot.preprocess(d0)
    iv, inn, _ = ot.preprocess(d1)
    its, oest, odelta = oracle.icp_align(cfg, iv, inn, tv, tn, 20)
    for got in (facade, loop):                              # both routes: the oracle's 20-iteration pose within 1e-4
        assert rot_err(got[:3, :3], odelta[:3, :3]) <= 1e-4
        assert np.max(np.abs(got[:3, 3] - odelta[:3, 3])) <= 1e-4
    assert np.max(np.abs(facade - loop)) <= 2e-5           # fused association vs stored correspondences: summation order only
    # host LDLT (Solver::SolveJacobianSystem) against an fp64 solve + matrix exponential of that very system
    from scipy.linalg import expm
    x = -np.linalg.solve(JTJ, JTr)
    Xi = np.zeros((4, 4))
    Xi[:3, :3] = [[0, -x[5], x[4]], [x[5], 0, -x[3]], [-x[4], x[3], 0]]
    Xi[:3, 3] = x[:3]
    assert np.max(np.abs(ldlt - expm(Xi))) <= 1e-6
    assert np.max(np.abs(back - np.array([0.02, -0.01, 0.03, 0.04, -0.02, 0.01], np.float32))) <= 1e-6


def _system32(JtJ, Jtr, count=1e5, err=1.0):
    s = np.zeros(32, np.float32)
    k = 0
    for i in range(6):
        for j in range(i, 6):
            s[k] = JtJ[i, j]
            k += 1
    s[21:27] = Jtr
    s[27], s[28] = err, count
    return s


@pytest.mark.parametrize("policy", [POLICY_REF_EXACT, POLICY_FIXED])
def test_device_solve_against_fp64_on_conditioned_and_degenerate_systems(built_library, oracle, policy):
    """The on-device 6x6 solve + SE(3) update (fp32 LDL^T without pivoting, closed-form exponential; replaces Eigen's
    inverse() / exp() / log(), ref Solver.cpp:91-111, SE3.cpp:4-19 -- the parity-UNPINNED half of a23) against an fp64
    numpy solve + scipy expm: well-conditioned, badly scaled (cond 1e6), and rank-deficient (a planar scene) systems."""
    from scipy.linalg import expm

    ctx = Context(Config(policy=policy, numBuckets=64, numVoxelBlocks=64))
    rng = np.random.default_rng(7)

    def run(JtJ, Jtr, start=None):
        ctx.icp_reset(True)
        if start is not None:
            ctx.icp_set_twist(start)
        before = ctx.icp_get()[0].astype(np.float64)
        ctx.icp_solve(cu(_system32(JtJ, Jtr)))
        torch.cuda.synchronize()
        return before, ctx.icp_get()[0].astype(np.float64)

    def want(JtJ, Jtr, before):
        x = -np.linalg.solve(JtJ.astype(np.float64), Jtr.astype(np.float64))
        Xi = np.zeros((4, 4))
        Xi[:3, :3] = [[0, -x[5], x[4]], [x[5], 0, -x[3]], [-x[4], x[3], 0]]
        Xi[:3, 3] = x[:3]
        return expm(Xi) @ before, x

    # (1) the normal equations of a real frame pair: rows (n, p x n) of a sphere + plane scene
    for trial in range(4):
        P = rng.uniform(-1, 1, (4000, 3)) + np.array([0, 0, 2.5])
        Nn = rng.normal(size=(4000, 3))
        Nn /= np.linalg.norm(Nn, axis=1, keepdims=True)
        J = np.concatenate([Nn, np.cross(P, Nn)], axis=1)
        r = rng.normal(scale=0.01, size=4000)
        JtJ, Jtr = (J.T @ J).astype(np.float32), (J.T @ r).astype(np.float32)
        start = np.array([0.03, -0.02, 0.01, 0.02, 0.01, -0.03], np.float32) if trial % 2 else None
        before, got = run(JtJ, Jtr, start)
        exp_, x = want(JtJ, Jtr, before)
        cond = np.linalg.cond(JtJ.astype(np.float64))
        assert np.max(np.abs(got - exp_)) <= 2e-6 + 4e-7 * cond * np.max(np.abs(x)), (trial, cond)
        assert np.max(np.abs(got[:3, :3].T @ got[:3, :3] - np.eye(3))) <= 1e-6       # re-orthonormalised
    # (2) badly scaled but full rank (cond ~ 1e6): the error grows with the condition number, not beyond it
    Q, _ = np.linalg.qr(rng.normal(size=(6, 6)))
    JtJ = (Q @ np.diag([1e6, 3e5, 1e4, 1e3, 30.0, 1.0]) @ Q.T)
    JtJ = ((JtJ + JtJ.T) / 2).astype(np.float32)
    Jtr = (JtJ.astype(np.float64) @ np.array([1e-3, -2e-3, 1e-3, 2e-3, -1e-3, 1e-3])).astype(np.float32)
    before, got = run(JtJ, Jtr)
    exp_, x = want(JtJ, Jtr, before)
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - exp_)) <= 1e-6 * np.linalg.cond(JtJ.astype(np.float64)) * 1e-3 + 1e-5
    # (3) rank-deficient: a single plane constrains 3 of the 6 degrees of freedom (exactly representable rows, so the
    # elimination meets an exact zero pivot).  The solve must refuse -- state unchanged, iteration stopped -- not return garbage.
    Pp = np.stack([rng.integers(-8, 8, 64), rng.integers(-8, 8, 64), np.full(64, 2)], axis=1).astype(np.float64)
    Np = np.tile([0.0, 0.0, 1.0], (64, 1))
    J = np.concatenate([Np, np.cross(Pp, Np)], axis=1)
    JtJ, Jtr = (J.T @ J).astype(np.float32), (J.T @ np.full(64, 0.25)).astype(np.float32)
    assert np.linalg.matrix_rank(JtJ.astype(np.float64)) == 3
    before, got = run(JtJ, Jtr, np.array([0.01, 0.02, -0.01, 0.01, -0.02, 0.02], np.float32))
    assert np.array_equal(before, got), "a singular system changed the pose"
    # and the whole Align stops cleanly on such a scene (a fronto-parallel wall): finite pose, no iteration counted past the failure
    cfgp = Config(policy=policy)
    wall = np.full((cfgp.height, cfgp.width), 10000, np.uint16)
    c2 = Context(cfgp)
    v, n, _ = gpu_preprocess(c2, wall)
    c2.icp_reset(True)
    c2.icp_align(v, n, v, n, 20)
    d = c2.icp_get()[0]
    assert np.all(np.isfinite(d))


def test_linear_system_300_contract(built_library, oracle):
    """buildLinearSystemOnDevice (the un-built LinearSystem.cu reducer): 300 x 27 partials whose sum is
    A^T A | A^T b with A = (s x n, n), b = n.d - n.s."""
    cfg = Config()
    ot = oracle.OracleTable(cfg)
    tv, tn, _ = ot.preprocess(render(cfg, scenes.scene_S1T(), scenes.trajectory_C2(0)))
    iv, _, _ = ot.preprocess(render(cfg, scenes.scene_S1T(), scenes.trajectory_C2(10)))
    lib = L.load_library()
    n = 640 * 480
    d_out = torch.zeros(300 * 27, device="cuda")
    h_out = np.zeros(300 * 27, np.float32)
    g_iv, g_tv, g_tn = cu(iv), cu(tv), cu(tn)
    lib.buildLinearSystemOnDevice(g_iv.data_ptr(), g_tv.data_ptr(), g_tn.data_ptr(), d_out.data_ptr(), h_out.ctypes.data)
    got = h_out.reshape(300, 27).astype(np.float64).sum(0)
    s, d, nn = iv[:, :3].astype(np.float64), tv[:, :3].astype(np.float64), tn[:, :3].astype(np.float64)
    A = np.concatenate([np.cross(s, nn), nn], axis=1)
    b = np.sum(nn * d, 1) - np.sum(nn * s, 1)
    AtA, Atb = A.T @ A, A.T @ b
    want = np.concatenate([AtA[np.triu_indices(6)], Atb])
    assert np.allclose(got, want, rtol=2e-5, atol=1e-5 * np.max(np.abs(want)))
    assert np.array_equal(h_out, d_out.cpu().numpy())


def test_raycast_vs_oracle_and_scene(built_library, oracle):
    cfg = fixed_cfg(width=320, height=240, fx=517.3 / 2, fy=516.5 / 2, cx=318.6 / 2, cy=255.3 / 2, numVoxelBlocks=4096)
    ot = oracle.OracleTable(cfg)
    ctx = Context(cfg)
    for k in range(0, 12, 2):
        pose = scenes.trajectory_C2(k).astype(np.float32)
        depth = render(cfg, scenes.scene_S1(), pose)
        ov, _, odf = ot.preprocess(depth)
        ot.fuse_frame(pose, ov, odf)
        v, n, df = gpu_preprocess(ctx, depth)
        ctx.fuse_frame(pose, v, n, df)
    pose = scenes.trajectory_C2(6).astype(np.float32)
    ctx.set_pose(pose)
    rv, rn = torch.zeros_like(v), torch.zeros_like(n)
    ctx.raycast(rv, rn)
    torch.cuda.synchronize()
    gv, gn = rv.cpu().numpy(), rn.cpu().numpy()
    cv, cn = ot.raycast(pose)
    hit_g, hit_c = gv[:, 2] > 0, cv[:, 2] > 0
    assert np.mean(hit_g == hit_c) > 0.999
    both = hit_g & hit_c
    assert both.sum() > 0.8 * cfg.width * cfg.height
    assert np.mean(np.abs(gv[both] - cv[both]).max(1) <= 1e-4) > 0.999
    nb = both & (np.abs(gn).sum(1) > 0) & (np.abs(cn).sum(1) > 0)
    assert np.mean(np.abs(gn[nb] - cn[nb]).max(1) <= 1e-3) > 0.999
    # against the analytic scene: depth error below one voxel away from silhouettes
    truth = render(cfg, scenes.scene_S1(), pose).reshape(-1).astype(np.float64) / cfg.depthScale
    err = np.abs(gv[:, 2] - truth)[hit_g & (truth > 0)]
    assert np.median(err) < 0.25 * cfg.voxelSize and np.mean(err < cfg.voxelSize) > 0.97
    # normals point at the camera (negative z), unit length
    nz = gn[nb]
    assert np.all(np.abs(np.linalg.norm(nz[:, :3], axis=1) - 1.0) < 1e-4) and np.mean(nz[:, 2] < 0) > 0.999


def test_legacy_entry_points_match_handle_api(built_library, oracle):
    """The reference's own call sequence (SDF_Hashtable.cpp:60-81 ctor, :11-40 integrate) through the
    legacy C symbols gives the same table as the handle API and the oracle."""
    lib = L.load_library()
    cfg = Config()
    depth = render(cfg, scenes.scene_S1(), np.eye(4))
    ot = oracle.OracleTable(cfg)
    ov, on, _ = ot.preprocess(depth)
    inserted, nvis, nupd = ref_fuse_cpu(ot, np.eye(4, dtype=np.float32), ov)

    p = cfg.to_c().table
    K, Kinv = cfg.K(), cfg.Kinv()
    assert lib.SetCameraIntrinsic(K.ctypes.data, Kinv.ctypes.data)         # CameraTracking.cpp:134
    lib.updateConstantHashTableParams(C.byref(p))                          # SDF_Hashtable.cpp:75
    lib.deviceAllocate(C.byref(p))                                         # :76
    lib.calculateKinectProjectionMatrix()                                  # :79
    d = cu(depth.reshape(-1))
    v = torch.zeros((640 * 480, 4), device="cuda")
    n = torch.zeros((640 * 480, 4), device="cuda")
    lib.preProcess(v.data_ptr(), n.data_ptr(), d.data_ptr())               # Application.cpp:73
    assert np.array_equal(bits(v.cpu().numpy()), bits(ov)) and np.array_equal(bits(n.cpu().numpy()), bits(on))
    lib.mapGLobjectsToCUDApointers(None, None, None)                       # SDF_Hashtable.cpp:13
    lib.updateConstantHashTableParams(C.byref(p))                          # :21
    for _ in range(4):                                                     # to convergence (see ref_fuse_gpu)
        lib.resetHashTableMutexes(C.byref(p))                              # :24
        lib.allocBlocks(v.data_ptr(), n.data_ptr())                        # :27
    count = lib.flattenIntoBuffer(C.byref(p))                              # :30
    assert count == nvis
    p.numOccupiedBlocks = count                                            # :32
    lib.updateConstantHashTableParams(C.byref(p))                          # :33
    lib.integrateDepthMap(C.byref(p), v.data_ptr())                        # :36
    # read back through the stand-ins of the three GL buffers
    cnt = np.zeros(1, np.int32)
    compact = np.zeros(count, dtype=L.VOXEL_ENTRY_DTYPE)
    cudart = torch.cuda.cudart()
    torch.cuda.synchronize()
    h = C.c_void_p(lib.vhLegacyContext())
    n_ent = C.c_int(0)
    L.check(lib.vh_export_compact(h, compact.ctypes.data, count, C.byref(n_ent)))
    assert n_ent.value == count and entries_to_set(compact) == entries_to_set(ot.compact_entries())
    blk = np.zeros(512, dtype=L.VOXEL_DTYPE)
    cpu_blocks = ot.block_dict()
    for e in compact[:: max(1, count // 16)]:
        L.check(lib.vh_export_block(h, int(e["ptr"]), blk.ctypes.data))
        ref = cpu_blocks[(int(e["x"]), int(e["y"]), int(e["z"]))]
        assert np.array_equal(bits(blk.view(np.float32).reshape(512, 2)), bits(ref))
    # computeCorrespondences + CalculateJacobiansAndResiduals (CameraTracking.cpp:53, Solver.cpp:74)
    corr, corrN = torch.zeros((640 * 480, 4), device="cuda"), torch.zeros((640 * 480, 4), device="cuda")
    res, J = torch.zeros(640 * 480, device="cuda"), torch.zeros((640 * 480, 6), device="cuda")
    delta = L.Float4x4()
    dm = oracle.se3_exp([0.005, 0, 0, 0, 0.002, 0]).reshape(16)
    for i in range(16):
        delta.entries[i] = float(dm[i])
    err = lib.computeCorrespondences(v.data_ptr(), v.data_ptr(), n.data_ptr(), corr.data_ptr(), corrN.data_ptr(), res.data_ptr(),
                                     C.byref(delta), 640, 480)
    oerr, ocorr, ocorrN, ores = oracle.find_correspondences(cfg, ov, None, ov, on, dm)
    assert np.array_equal(bits(res.cpu().numpy()), bits(ores))
    assert abs(err - float(np.sum(ores.astype(np.float64)))) <= 1e-5 * float(np.sum(np.abs(ores)))   # fp32 sum, any order
    lib.CalculateJacobiansAndResiduals(v.data_ptr(), corr.data_ptr(), corrN.data_ptr(), J.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(bits(J.cpu().numpy()), bits(oracle.jacobians(cfg, ocorr, ocorrN)))
    lib.deviceFree()
    del cudart, cnt


def test_pipeline_tracks_synthetic_trajectory(built_library, oracle):
    """Native frame loop (graph replay) on 40 frames of config C2: pose follows the ground truth, and the
    graph path is bit-identical to plain launches."""
    cfg = fixed_cfg(numVoxelBlocks=16384, icpNormalThres=0.8)
    poses = [scenes.trajectory_C2(k) for k in range(40)]
    frames = [cu(render(cfg, scenes.scene_S1T(), p).reshape(-1)) for p in poses]
    results = []
    models = []
    # graph replay / plain launches / overlapped schedule (input produced on the stream; input ready: its pre-processing runs
    # beside the previous Align) / pre-processing fused into the Align kernel
    for use_graph, overlap, fused_pre, ready in ((True, False, False, False), (False, False, False, False), (True, True, False, False),
                                                 (True, True, True, False), (True, True, False, True)):
        os.environ["VH_PIPE_FUSED_PRE"] = "1" if fused_pre else "0"      # read at pipeline creation
        ctx = Context(cfg)
        pipe = FramePipeline(ctx, iterations=10, mode=FramePipeline.FRAME_TO_FRAME, use_graph=use_graph, overlap=overlap)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            pipe.reset(poses[0].astype(np.float32))
            for f in frames:
                pipe.push_device_ready(f, None) if ready else pipe.push_device(f)
            pose = pipe.pose()
            st = ctx.stats(s)
        results.append((pose, st.numAllocated, int(st.numUpdated), pipe.launches()))
        models.append(ctx.block_dict())
        truth = poses[-1]
        assert rot_err(pose[:3, :3], truth[:3, :3]) < 5e-3
        assert np.max(np.abs(pose[:3, 3] - truth[:3, 3])) < 0.01
        assert st.numAllocated > 300 and st.dropped == 0
    for r in results[1:]:                                   # plain launches and the overlapped schedule: same bits
        assert np.array_equal(bits(results[0][0]), bits(r[0]))
        assert results[0][1:3] == r[1:3]
    # kernels per frame: preprocess, Align, frame constants, alloc + compact + integrate
    assert results[0][3] == results[1][3] == results[2][3] == results[4][3] == 40 + 39 + 40 + 40 * 3
    assert results[3][3] == 1 + 39 + 40 + 40 * 3              # pre-processing of the tracked frames inside the Align kernel
    os.environ.pop("VH_PIPE_FUSED_PRE", None)
    for m in models[1:]:
        exact, _ = compare_blocks(m, models[0])
        assert exact


def test_frame_to_model_tracking(built_library, oracle):
    """Track-integrate-raycast loop (config C5): ICP target = raycast of the model."""
    cfg = fixed_cfg(width=320, height=240, fx=517.3 / 2, fy=516.5 / 2, cx=318.6 / 2, cy=255.3 / 2, numVoxelBlocks=16384,
                    icpNormalThres=0.8)
    poses = [scenes.trajectory_C2(k) for k in range(30)]
    ctx = Context(cfg)
    pipe = FramePipeline(ctx, iterations=10, mode=FramePipeline.FRAME_TO_MODEL, use_graph=True)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        pipe.reset(poses[0].astype(np.float32))
        for p in poses:
            pipe.push_device(cu(render(cfg, scenes.scene_S1T(), p).reshape(-1)))
        pose = pipe.pose()
    assert rot_err(pose[:3, :3], poses[-1][:3, :3]) < 5e-3
    assert np.max(np.abs(pose[:3, 3] - poses[-1][:3, 3])) < 0.01


def test_partition_union_equals_single(built_library, oracle):
    """owner = mix(block) mod P: the union over ranks equals the single-table result, block by block."""
    base = fixed_cfg(numVoxelBlocks=4096)
    depths = [render(base, scenes.scene_S1(), scenes.trajectory_C2(k)) for k in (0, 6)]
    single = Context(base)
    parts = [Context(fixed_cfg(numVoxelBlocks=4096, partCount=3, partRank=r)) for r in range(3)]
    for k, depth in zip((0, 6), depths):
        pose = scenes.trajectory_C2(k).astype(np.float32)
        for c in [single] + parts:
            v, n, df = gpu_preprocess(c, depth)
            c.fuse_frame(pose, v, n, df)
    whole = single.block_dict()
    union = {}
    for c in parts:
        d = c.block_dict()
        assert not (set(d) & set(union)), "a block is owned by two ranks"
        union.update(d)
    exact, _ = compare_blocks(union, whole)
    assert exact
    assert all(len(c.export_entries()) > 0.2 * len(whole) for c in parts)
    # the oracle applies the same ownership rule
    ot = oracle.OracleTable(fixed_cfg(numVoxelBlocks=4096, partCount=3, partRank=1))
    for k, depth in zip((0, 6), depths):
        ov, _, odf = ot.preprocess(depth)
        ot.fuse_frame(scenes.trajectory_C2(k).astype(np.float32), ov, odf)
    assert entries_to_set(parts[1].export_entries()) == entries_to_set(ot.entries())


def test_checkpoint_roundtrip_and_dump(built_library, oracle, tmp_path):
    cfg = fixed_cfg(numVoxelBlocks=2048)
    depth = render(cfg, scenes.scene_S1(), np.eye(4))
    a = Context(cfg)
    v, n, df = gpu_preprocess(a, depth)
    a.fuse_frame(np.eye(4, dtype=np.float32), v, n, df)
    a.save(tmp_path / "model.vhb")
    b = Context(cfg)
    b.load(tmp_path / "model.vhb")
    ex, _ = compare_blocks(b.block_dict(), a.block_dict())
    assert ex and b.stats().heapCounter == a.stats().heapCounter
    # the loaded model keeps fusing identically
    for c in (a, b):
        c.fuse_frame(np.eye(4, dtype=np.float32), v, n, df)
    ex, _ = compare_blocks(b.block_dict(), a.block_dict())
    assert ex
    a.dump_text(tmp_path / "SDF_dump.txt")                 # format of SDFRenderer.cpp:90-108
    lines = (tmp_path / "SDF_dump.txt").read_text().splitlines()
    assert lines[0] == f"numOccupiedBlocks from GL :{a.stats().numVisible}"
    assert lines[4].startswith("0) : pos : (") and len(lines[5].rstrip("\t").split("\t")) == 512


# ---- full-size, size-independent properties (configs C2 / C3) --------------------------------------
@pytest.mark.parametrize("name", ["C2", "C3", "C4"])
def test_full_size_properties(built_library, oracle, name):
    if name == "C2":
        cfg = fixed_cfg(numVoxelBlocks=65536)
        scene, traj = scenes.scene_S1T(), scenes.trajectory_C2
    elif name == "C4":   # BASELINE.json configs[3] as stated: building-scale scene at 2 mm voxels (1.2 M blocks = 4.9 GB in one frame)
        cfg = fixed_cfg(width=1280, height=720, fx=1034.6, fy=1033.0, cx=637.2, cy=382.95, voxelSize=0.002, truncation=0.008,
                        truncScale=0.001, numBuckets=4000037, numVoxelBlocks=1310720, overflowSlots=524288, depthMax=12.5,
                        maxIntegrationDistance=12.5)
        scene, traj = scenes.scene_S3(), lambda k: scenes.trans(0, 0, 0.5) @ scenes.trajectory_C3(k)
    else:
        cfg = fixed_cfg(width=1280, height=720, fx=1034.6, fy=1033.0, cx=637.2, cy=382.95, voxelSize=0.005, truncation=0.02,
                        numBuckets=1000003, numVoxelBlocks=262144, overflowSlots=65536, depthMax=8.0, maxIntegrationDistance=8.0)
        scene, traj = scenes.scene_S2(), lambda k: scenes.trans(0, 0, 0.3) @ scenes.trajectory_C3(k)
    ctx = Context(cfg)
    pose = traj(0).astype(np.float32)
    depth = render(cfg, scene, pose)
    v, n, df = gpu_preprocess(ctx, depth)
    ctx.fuse_frame(pose, v, n, df)
    s1 = ctx.stats()
    assert s1.numAllocated > 0 and s1.dropped == 0 and s1.numVisible == s1.numAllocated
    ent = ctx.export_entries()
    assert len(entries_to_set(ent)) == len(ent) == s1.numAllocated          # no duplicate keys
    assert sorted(ent["ptr"] // 512) == list(range(cfg.numVoxelBlocks - s1.numAllocated, cfg.numVoxelBlocks))   # heap order (ref :207/:331)
    # idempotence: allocating the same frame again inserts nothing
    ctx.fuse_frame(pose, v, n, df)
    s2 = ctx.stats()
    assert s2.numAllocated == s1.numAllocated and s2.lastInserted == 0 and s2.numUpdated == s1.numUpdated
    # weights: every updated voxel saw the same sample twice => sdf unchanged, weight = min(max, 2 w)
    e = ent[len(ent) // 2]
    blk = ctx.export_block(int(e["ptr"]))
    w = blk["weight"][blk["weight"] > 0]
    assert w.size > 0 and np.all(w >= 2.0) and np.all(w <= cfg.integrationWeightMax)
    # every pixel's surface voxel is allocated: the block containing each valid vertex is in the table
    vv = v.cpu().numpy()
    valid = (vv[:, 2] > cfg.depthMin) & (vv[:, 2] < cfg.depthMax)
    pw = (vv[valid][::97, :3].astype(np.float64) @ pose[:3, :3].T.astype(np.float64)) + pose[:3, 3]
    blocks = np.floor((pw / cfg.voxelSize + 0.5) / 8.0).astype(np.int64)
    have = entries_to_set(ent)
    missing = sum(tuple(int(c) for c in b) not in have for b in blocks)
    assert missing <= 0.001 * len(blocks)
    # oracle cross-check of the counts at full size
    ot = oracle.OracleTable(cfg)
    ov, _, odf = ot.preprocess(depth)
    rep, nvis, nupd = ot.fuse_frame(pose, ov, odf)
    assert (rep.inserted, nvis, nupd) == (s1.numAllocated, s1.numVisible, int(s1.numUpdated))
    assert entries_to_set(ent) == entries_to_set(ot.entries())
    # ... and of the voxels themselves on a sample of blocks, bit for bit (the oracle has fused the frame once, the GPU twice:
    # compare against a fresh context that fused it once)
    once = Context(cfg)
    v1, n1, df1 = gpu_preprocess(once, depth)
    once.fuse_frame(pose, v1, n1, df1)
    ent1 = once.export_entries()
    for e in ent1[:: max(1, len(ent1) // 64)]:
        g = once.export_block(int(e["ptr"]))
        o = ot.block(int(e["x"]), int(e["y"]), int(e["z"]))
        assert o is not None
        assert np.array_equal(bits(g["sdf"]), bits(o[:, 0])) and np.array_equal(bits(g["weight"]), bits(o[:, 1]))


# ---- edge cases: empty, ragged, degenerate inputs -----------------------------------------------------
@pytest.mark.parametrize("policy", [POLICY_REF_EXACT, POLICY_FIXED])
def test_empty_frame_is_a_no_op(built_library, oracle, policy):
    """An all-zero depth image: nothing requested, nothing visible, nothing integrated, ICP declines to solve."""
    cfg = fixed_cfg(numVoxelBlocks=512) if policy == POLICY_FIXED else Config(numVoxelBlocks=512)
    depth = np.zeros((cfg.height, cfg.width), np.uint16)
    ctx = Context(cfg)
    v, n, df = gpu_preprocess(ctx, depth)
    assert not v[:, :3].any().item() and not n.any().item()
    ctx.fuse_frame(np.eye(4, dtype=np.float32), v, n, df if policy == POLICY_FIXED else None)
    st = ctx.stats()
    assert (st.numAllocated, st.numVisible, int(st.numUpdated), st.dropped) == (0, 0, 0, 0)
    assert len(ctx.export_entries()) == 0
    ctx.icp_reset(True)
    ctx.icp_align(v, n, v, n, 3)
    delta, twist, sysv = ctx.icp_get()
    assert np.array_equal(delta, np.eye(4, dtype=np.float32)) and not twist.any() and sysv[28] == 0
    if policy == POLICY_FIXED:
        rv, rn = torch.ones_like(v), torch.ones_like(n)
        ctx.raycast(rv, rn)
        torch.cuda.synchronize()
        assert not rv[:, :3].any().item() and not rn.any().item()


@pytest.mark.parametrize("policy", [POLICY_REF_EXACT, POLICY_FIXED])
def test_ragged_image_size_all_stages(built_library, oracle, policy):
    """161 x 123: not a multiple of any tile (32x8 alloc / preprocess tiles, 16x8 raycast tiles, 5-pixel ICP batches)."""
    kw = dict(width=161, height=123, fx=517.3 / 4, fy=516.5 / 4, cx=318.6 / 4, cy=255.3 / 4, numVoxelBlocks=4096)
    cfg = fixed_cfg(icpNormalThres=0.8, **kw) if policy == POLICY_FIXED else Config(**kw)
    ot = oracle.OracleTable(cfg)
    ctx = Context(cfg)
    maps = []
    for k in (0, 9):
        pose = scenes.trajectory_C2(k).astype(np.float32)
        depth = render(cfg, scenes.scene_S1T(), pose)
        ov, on, odf = ot.preprocess(depth)
        v, n, df = gpu_preprocess(ctx, depth)
        assert np.array_equal(bits(v.cpu().numpy()), bits(ov)) and np.array_equal(bits(n.cpu().numpy()), bits(on))
        if policy == POLICY_FIXED:
            ot.fuse_frame(pose, ov, odf)
            ctx.fuse_frame(pose, v, n, df)
        else:
            ref_fuse_cpu(ot, pose, ov)
            ref_fuse_gpu(ctx, pose, v, n)
        maps.append((ov, on, v, n))
    exact, _ = compare_blocks(ctx.block_dict(), ot.block_dict())
    assert exact
    (tv, tn, gtv, gtn), (iv, inn, giv, gin) = maps
    ctx.icp_reset(True)
    ctx.icp_align(giv, gin, gtv, gtn, 10)
    g = ctx.icp_get()[0]
    _, _, o = oracle.icp_align(cfg, iv, inn, tv, tn, 10)
    assert rot_err(g[:3, :3], o[:3, :3]) <= 1e-4 and np.max(np.abs(g[:3, 3] - o[:3, 3])) <= 1e-4


def test_table_and_chain_full_drops_are_counted(built_library, oracle):
    """4 buckets x 1 slot + chains of 2: at most 12 blocks fit; the rest are dropped, counted, never duplicated."""
    cfg = fixed_cfg(numBuckets=4, bucketSize=1, attachedLinkedListSize=2, overflowSlots=64, numVoxelBlocks=1024,
                    width=160, height=120, fx=517.3 / 4, fy=516.5 / 4, cx=318.6 / 4, cy=255.3 / 4)
    depth = render(cfg, scenes.scene_S1(), np.eye(4))
    ctx = Context(cfg)
    v, n, df = gpu_preprocess(ctx, depth)
    for _ in range(2):
        ctx.fuse_frame(np.eye(4, dtype=np.float32), v, n, df)
    st = ctx.stats()
    ent = ctx.export_entries()
    live = ent[ent["ptr"] >= 0]
    assert st.dropped > 0 and 0 < len(live) <= 4 * (1 + 2) and st.numAllocated == len(live)
    assert len(entries_to_set(live)) == len(live)
    ot = oracle.OracleTable(cfg)
    ov, _, odf = ot.preprocess(depth)
    ot.fuse_frame(np.eye(4), ov, odf)
    assert len(ot.entries()) == len(live)          # capacity is order-independent even if the survivors are not


def test_large_rotation_and_far_translation_pose(built_library, oracle):
    """Negative block coordinates, a 90-degree yaw and a far offset: hash, floor division and frustum math."""
    cfg = fixed_cfg(numVoxelBlocks=8192)
    pose = (scenes.trans(-7.3, 2.1, -4.9) @ scenes.rot_y(90.0)).astype(np.float32)
    depth = render(cfg, scenes.scene_S1(), np.eye(4))       # the camera sees S1 from its own frame
    ot = oracle.OracleTable(cfg)
    ctx = Context(cfg)
    ov, _, odf = ot.preprocess(depth)
    rep, nvis, nupd = ot.fuse_frame(pose, ov, odf)
    v, n, df = gpu_preprocess(ctx, depth)
    ctx.fuse_frame(pose, v, n, df)
    st = ctx.stats()
    keys = entries_to_set(ctx.export_entries())
    assert keys == entries_to_set(ot.entries()) and min(k[0] for k in keys) < -30 and min(k[2] for k in keys) < -20
    assert (st.numVisible, int(st.numUpdated)) == (nvis, nupd)
    exact, _ = compare_blocks(ctx.block_dict(), ot.block_dict())
    assert exact


# ---- starvation + garbage collection ----------------------------------------------------
def _same_model(ctx, ot):
    st = ctx.stats()
    ent = ctx.export_entries()
    assert entries_to_set(ent) == entries_to_set(ot.entries()) and len(entries_to_set(ent)) == len(ent)
    assert st.heapCounter == ot.heap_counter()
    exact, _ = compare_blocks(ctx.block_dict(), ot.block_dict())
    assert exact


@pytest.mark.parametrize("chained", [False, True])
def test_garbage_collect_matches_oracle(built_library, oracle, chained):
    """Release, tombstone reuse and starvation against the oracle; chained=True forces most blocks into overflow chains."""
    kw = dict(numBuckets=64, bucketSize=2, attachedLinkedListSize=64) if chained else {}
    cfg = fixed_cfg(numVoxelBlocks=4096, width=160, height=120, fx=517.3 / 4, fy=516.5 / 4, cx=318.6 / 4, cy=255.3 / 4, **kw)
    ot = oracle.OracleTable(cfg)
    ctx = Context(cfg)
    frames = []
    for k in (0, 12):
        pose = scenes.trajectory_C2(k).astype(np.float32)
        depth = render(cfg, scenes.scene_S1(), pose)
        ov, _, odf = ot.preprocess(depth)
        v, n, df = gpu_preprocess(ctx, depth)
        frames.append((pose, ov, odf, v, n, df))
        ot.fuse_frame(pose, ov, odf)
        ctx.fuse_frame(pose, v, n, df)
    _same_model(ctx, ot)
    # 1. default criterion over every block: only never-observed blocks go
    freed = ot.garbage_collect(scope=1)
    ctx.garbage_collect(L.VH_GC_ALL)
    st = ctx.stats()
    assert freed > 0 and st.lastFreed == freed and st.numVisible == 0
    _same_model(ctx, ot)
    # 2. fuse again: the released keys are re-requested and land in reclaimed tombstones
    pose, ov, odf, v, n, df = frames[0]
    ot.fuse_frame(pose, ov, odf)
    ctx.fuse_frame(pose, v, n, df)
    _same_model(ctx, ot)
    # 3. visible scope with ageing and a tight |sdf| threshold: only blocks of the current view are touched
    pose, ov, odf, v, n, df = frames[1]
    nvis = ot.compact(pose)
    ctx.set_pose(pose)
    ctx.compact()
    assert ctx.stats().numVisible == nvis
    freed = ot.garbage_collect(scope=0, sdf_threshold=0.03, weight_decay=2.5)
    ctx.garbage_collect(L.VH_GC_VISIBLE, 0.03, 2.5)
    assert freed > 0 and ctx.stats().lastFreed == freed
    _same_model(ctx, ot)
    # 4. starve everything, then rebuild: identical to a model that never saw the first frames
    n_all = len(ot.entries())
    assert ot.garbage_collect(scope=1, weight_decay=1e9) == n_all
    ctx.garbage_collect(L.VH_GC_ALL, 0.0, 1e9)
    st = ctx.stats()
    assert st.lastFreed == n_all and st.numAllocated == 0 and len(ctx.export_entries()) == 0
    pose, ov, odf, v, n, df = frames[0]
    ot.fuse_frame(pose, ov, odf)
    ctx.fuse_frame(pose, v, n, df)
    _same_model(ctx, ot)
    fresh = Context(cfg)
    fresh.fuse_frame(pose, v, n, df)
    exact, _ = compare_blocks(ctx.block_dict(), fresh.block_dict())
    assert exact, "released blocks were not zeroed"
    # the raycaster walks through tombstones: same image as from the fresh model
    if not chained:
        ra, rb = ctx.new_maps()[:2], fresh.new_maps()[:2]
        for c, r in ((ctx, ra), (fresh, rb)):
            c.set_pose(pose)
            c.raycast(*r)
        torch.cuda.synchronize()
        assert torch.equal(ra[0], rb[0]) and torch.equal(ra[1], rb[1])


def test_garbage_collect_checkpoint_and_policy(built_library, oracle, tmp_path):
    cfg = fixed_cfg(numVoxelBlocks=2048)
    depth = render(cfg, scenes.scene_S1(), np.eye(4))
    a = Context(cfg)
    v, n, df = gpu_preprocess(a, depth)
    pose = np.eye(4, dtype=np.float32)
    a.fuse_frame(pose, v, n, df)
    a.garbage_collect(L.VH_GC_ALL)                          # leaves holes in the id range
    assert a.stats().lastFreed > 0
    a.save(tmp_path / "gc.vhb")
    b = Context(cfg)
    b.load(tmp_path / "gc.vhb")
    for c in (a, b):
        c.fuse_frame(pose, v, n, df)
    ex, _ = compare_blocks(b.block_dict(), a.block_dict())
    assert ex and b.stats().heapCounter == a.stats().heapCounter
    r = Context(Config(numVoxelBlocks=512))                  # RefExact: the reference has no working removal
    assert r.lib.vh_garbage_collect(r._h, 1, 0.0, 0.0, None) == L.VH_ERR_UNSUPPORTED


# ---- streaming in and out of the device -------------------------------------------------
def _streamed_dict(ent, vox):
    ent, vox = ent.cpu().numpy(), vox.cpu().numpy()
    return {tuple(int(x) for x in e[:3]): vox[e[3] // 512] for e in ent}


@pytest.mark.parametrize("pinned", [True, False])
def test_stream_out_in_matches_oracle(built_library, oracle, pinned):
    cfg = fixed_cfg(numVoxelBlocks=4096, width=160, height=120, fx=517.3 / 4, fy=516.5 / 4, cx=318.6 / 4, cy=255.3 / 4)
    ot = oracle.OracleTable(cfg)
    ctx = Context(cfg)
    frames = []
    for k in (0, 12):
        pose = scenes.trajectory_C2(k).astype(np.float32)
        depth = render(cfg, scenes.scene_S1(), pose)
        ov, _, odf = ot.preprocess(depth)
        v, n, df = gpu_preprocess(ctx, depth)
        frames.append((pose, ov, odf, v, n, df))
        ot.fuse_frame(pose, ov, odf)
        ctx.fuse_frame(pose, v, n, df)
    before = ctx.block_dict()
    # 1. everything farther than 2.2 m from the origin leaves; same blocks, same bytes as the oracle
    oent, ovox = ot.stream_out((0.0, 0.0, 0.0), 2.2, 4096)
    ent, vox = ctx.stream_out((0.0, 0.0, 0.0), 2.2, 4096, pinned_host=pinned)
    assert ent.is_pinned() == pinned and 0 < len(ent) == len(oent) < len(before)
    got, want = _streamed_dict(ent, vox), {tuple(int(x) for x in e[:3]): ovox[e[3] // 512] for e in oent}
    assert set(got) == set(want) and all(np.array_equal(bits(got[k]), bits(want[k])) for k in got)
    assert all(np.array_equal(bits(got[k]), bits(before[k])) for k in got)
    assert ctx.stats().numVisible == 0
    _same_model(ctx, ot)
    # 2. the camera looks at the region again: some streamed keys are re-allocated with new observations
    pose, ov, odf, v, n, df = frames[0]
    ot.fuse_frame(pose, ov, odf)
    ctx.fuse_frame(pose, v, n, df)
    _same_model(ctx, ot)
    reobserved = set(got) & set(ctx.block_dict())
    assert reobserved and reobserved != set(got)
    # 3. stream in: absent keys are copied, re-observed keys merged; order of the records must not matter
    perm = torch.randperm(len(ent), generator=torch.Generator().manual_seed(3))
    assert ot.stream_in(oent, ovox) == len(oent)
    assert ctx.stream_in(ent[perm].contiguous() if not pinned else ent[perm].contiguous().pin_memory(), vox) == len(ent)
    _same_model(ctx, ot)
    # 4. a buffer too small takes exactly its capacity; nothing is lost, nothing duplicated
    all_before = ctx.block_dict()
    ent2, vox2 = ctx.stream_out((0.0, 0.0, 0.0), 0.0, 10, pinned_host=pinned)
    rest = ctx.block_dict()
    moved = _streamed_dict(ent2, vox2)
    assert len(ent2) == 10 and not (set(moved) & set(rest)) and set(moved) | set(rest) == set(all_before)
    assert all(np.array_equal(bits(moved[k]), bits(all_before[k])) for k in moved)
    assert ctx.stream_in(ent2, vox2) == 10
    ex, _ = compare_blocks(ctx.block_dict(), all_before)
    assert ex


def test_headless_app_tracks_a_png_sequence(built_library, tmp_path):
    """The C++ host loop on the reference's input format: 16-bit depth PNGs -> vh_depth_read -> native pipeline."""
    import subprocess

    from voxelhashing_demo_b200 import read_depth, write_depth_png
    from voxelhashing_demo_b200._build import HOST_DEMO

    cfg = fixed_cfg()
    poses = [scenes.trajectory_C2(4 * k) for k in range(6)]
    files = []
    for k, p in enumerate(poses):
        d = render(cfg, scenes.scene_S1T(), p)
        f = tmp_path / f"T{k}.png"
        write_depth_png(f, d, filter_type=k % 5)
        assert np.array_equal(read_depth(f), d)
        files.append(str(f))
    r = subprocess.run([str(HOST_DEMO), "--frames", *files, "--mesh", str(tmp_path / "model.ply")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "mesh: " in r.stdout and (tmp_path / "model.ply").read_bytes().startswith(b"ply\nformat binary_little_endian 1.0")
    lines = [l for l in r.stdout.splitlines() if l.startswith("frame ")]
    assert len(lines) == 6
    t_last = np.array([float(x) for x in lines[-1].split("(")[1].rstrip(")").split()])
    assert np.max(np.abs(t_last - poses[-1][:3, 3])) < 5e-3          # 20 mm of travel tracked to a few mm
    assert "dropped 0" in r.stdout
    # and the reference's own two-frame loop (Application.cpp:24-103) still ends in OK
    r = subprocess.run([str(HOST_DEMO), str(tmp_path / "SDF_dump.txt")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr


# ---- optional bilateral front end --------------------------------------------------------
@pytest.mark.parametrize("size", ["vga", "ragged"])
def test_bilateral_preprocess_bit_exact(built_library, oracle, size):
    kw = dict(bilateralSigmaSpace=1.5, bilateralSigmaRange=0.03)
    cfg = fixed_cfg(**kw) if size == "vga" else fixed_cfg(width=161, height=123, fx=517.3 / 4, fy=516.5 / 4, cx=318.6 / 4, cy=255.3 / 4, **kw)
    depth = render(cfg, scenes.scene_S1T(), scenes.trajectory_C2(7))
    rng = np.random.default_rng(3)
    depth = np.where(depth > 0, (depth.astype(np.int32) + rng.integers(-20, 21, depth.shape)).clip(1, 65535), 0).astype(np.uint16)
    depth[5:9, 10:40] = 0
    ov, on, odf = oracle.OracleTable(cfg).preprocess(depth)
    ctx = Context(cfg)
    v, n, df = gpu_preprocess(ctx, depth)
    assert np.array_equal(bits(v.cpu().numpy()), bits(ov))
    assert np.array_equal(bits(n.cpu().numpy()), bits(on))
    assert np.array_equal(bits(df.cpu().numpy()), bits(odf))
    # and it does something: the filtered map differs from the raw one, the integration depth does not
    raw = Context(fixed_cfg(width=cfg.width, height=cfg.height, fx=cfg.fx, fy=cfg.fy, cx=cfg.cx, cy=cfg.cy))
    rv, rn, rdf = gpu_preprocess(raw, depth)
    assert not torch.equal(rv, v) and torch.equal(rdf, df)


def test_bilateral_front_end_helps_tracking_under_noise(built_library, oracle):
    """+-6 mm uniform depth noise on the C2 sequence: frame-to-frame ICP drifts less with the filtered maps."""
    errs = {}
    rng = np.random.default_rng(17)
    poses = [scenes.trajectory_C2(2 * k) for k in range(25)]
    base = fixed_cfg(numVoxelBlocks=16384, icpNormalThres=0.8)
    clean = [render(base, scenes.scene_S1T(), p) for p in poses]
    noisy = [np.where(d > 0, (d.astype(np.int32) + rng.integers(-30, 31, d.shape)).clip(1, 65535), 0).astype(np.uint16) for d in clean]
    for name, kw in (("raw", {}), ("bilateral", dict(bilateralSigmaSpace=1.5, bilateralSigmaRange=0.03))):
        ctx = Context(fixed_cfg(numVoxelBlocks=16384, icpNormalThres=0.8, **kw))
        pipe = FramePipeline(ctx, iterations=10, mode=FramePipeline.FRAME_TO_FRAME, use_graph=True, overlap=True)
        pipe.reset(poses[0].astype(np.float32))
        for d in noisy:
            pipe.push_device(cu(d.reshape(-1)))
        pose = pipe.pose()
        errs[name] = float(np.max(np.abs(pose[:3, 3] - poses[-1][:3, 3])))
        assert ctx.stats().dropped == 0
    assert errs["bilateral"] < errs["raw"] and errs["bilateral"] < 0.01, errs


# ---- mesh extraction ---------------------------------------------------------------------
def _sorted_tris(t):
    a = np.ascontiguousarray(t, np.float32).reshape(-1, 9).view(np.uint32)
    return a[np.lexsort(a.T[::-1])]


@pytest.mark.parametrize("chained", [False, True])
def test_mesh_extraction_matches_oracle(built_library, oracle, chained, tmp_path):
    """Same triangles, bit for bit, as the oracle's marching tetrahedra (order aside); neighbour blocks found through chains too."""
    kw = dict(numBuckets=64, bucketSize=2, attachedLinkedListSize=64) if chained else {}
    cfg = fixed_cfg(numVoxelBlocks=4096, width=160, height=120, fx=517.3 / 4, fy=516.5 / 4, cx=318.6 / 4, cy=255.3 / 4, **kw)
    ot = oracle.OracleTable(cfg)
    ctx = Context(cfg)
    for k in (0, 12):
        pose = scenes.trajectory_C2(k).astype(np.float32)
        depth = render(cfg, scenes.scene_S1(), pose)
        ov, _, odf = ot.preprocess(depth)
        v, n, df = gpu_preprocess(ctx, depth)
        ot.fuse_frame(pose, ov, odf)
        ctx.fuse_frame(pose, v, n, df)
    want = ot.extract_mesh()
    got = ctx.extract_mesh()
    assert len(got) == len(want) > 5000
    assert np.array_equal(_sorted_tris(got.cpu().numpy()), _sorted_tris(want))
    # a buffer that is too small is filled to its capacity and the true count is still reported
    small = torch.zeros((100, 3, 3), device="cuda")
    n = C.c_int(0)
    assert ctx.lib.vh_extract_mesh(ctx.handle, small.data_ptr(), 100, C.byref(n), None) == L.VH_OK and n.value == len(want)
    assert small.abs().sum().item() > 0
    # PLY round trip of the header and sizes
    cnt = ctx.save_mesh_ply(tmp_path / "m.ply", got)
    blob = (tmp_path / "m.ply").read_bytes()
    head, body = blob.split(b"end_header\n", 1)
    assert f"element vertex {3 * cnt}".encode() in head and f"element face {cnt}".encode() in head
    assert len(body) == cnt * 36 + cnt * 13
    assert np.array_equal(np.frombuffer(body[: cnt * 36], np.float32), got.cpu().numpy().reshape(-1))
    # an empty model has an empty mesh
    assert len(Context(cfg).extract_mesh()) == 0


def test_differential_fuzz_of_table_maintenance(built_library, oracle):
    """The same random interleaving of fusion / garbage collection / stream-out / stream-in on the GPU and on the oracle,
    on a tiny chain-heavy table; after EVERY step the two models must be identical (keys, heap counter, voxels bit for bit)."""
    cfg = fixed_cfg(numBuckets=32, bucketSize=2, attachedLinkedListSize=64, overflowSlots=4096, numVoxelBlocks=2048,
                    width=160, height=120, fx=517.3 / 4, fy=516.5 / 4, cx=318.6 / 4, cy=255.3 / 4)
    rng = np.random.default_rng(99)
    ot = oracle.OracleTable(cfg)
    ctx = Context(cfg)
    frames, parked = {}, []
    ops = []
    for step in range(45):
        op = int(rng.integers(0, 5))
        if op <= 1 or step == 0:
            k = int(rng.integers(0, 40))
            if k not in frames:
                pose = scenes.trajectory_C2(k).astype(np.float32)
                depth = render(cfg, scenes.scene_S1(), pose)
                ov, _, odf = ot.preprocess(depth)
                v, n, df = gpu_preprocess(ctx, depth)
                frames[k] = (pose, ov, odf, v, n, df)
            pose, ov, odf, v, n, df = frames[k]
            ot.fuse_frame(pose, ov, odf)
            ctx.fuse_frame(pose, v, n, df)
            ops.append(f"fuse {k}")
        elif op == 2:
            scope = int(rng.integers(0, 2))
            thr, dec = float(rng.choice([0.0, 0.02, 0.05])), float(rng.choice([0.0, 1.0, 3.0, 1e9]))
            freed = ot.garbage_collect(scope=scope, sdf_threshold=thr, weight_decay=dec)
            ctx.garbage_collect(scope, thr, dec)
            assert ctx.stats().lastFreed == freed, ops
            ops.append(f"gc {scope} {thr} {dec} -> {freed}")
        elif op == 3:
            c = rng.uniform(-1.0, 1.0, 3) + np.array([0.0, 0.0, 2.0])
            r = float(rng.uniform(0.3, 1.5))
            oent, ovox = ot.stream_out(c, r, 4096)
            ent, vox = ctx.stream_out(c, r, 4096, pinned_host=bool(step & 1))
            assert len(ent) == len(oent), ops
            if len(oent):
                parked.append((oent, ovox, ent, vox))
            ops.append(f"out {len(oent)}")
        elif parked:
            oent, ovox, ent, vox = parked.pop(int(rng.integers(0, len(parked))))
            assert ot.stream_in(oent, ovox) == len(oent) == ctx.stream_in(ent, vox), ops
            ops.append(f"in {len(oent)}")
        else:
            continue
        try:
            _same_model(ctx, ot)
        except AssertionError as e:
            raise AssertionError(f"models diverged after {ops}") from e
    assert sum(o.startswith("gc") for o in ops) >= 3 and sum(o.startswith("out") for o in ops) >= 3


def test_gpu_reproduces_the_fixed_policy_fixture(built_library, oracle):
    """tests/golden/fixed_small.npz (outputs of the oracle on two stored frames): the kernels give the same hashes for the
    table, the voxels, the mesh, the garbage-collected key set and the bilateral maps, and the same ICP delta to 1e-4."""
    import hashlib
    import importlib.util

    from pathlib import Path

    gold_dir = Path(__file__).resolve().parent / "golden"
    spec = importlib.util.spec_from_file_location("make_fixed_golden", gold_dir / "make_fixed_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(gold_dir / "fixed_small.npz")
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    cfg = mod.config()
    ctx = Context(cfg)
    maps = []
    for i in range(2):
        v, n, df = gpu_preprocess(ctx, g[f"depth{i}"])
        maps.append((v, n))
        ctx.fuse_frame(g[f"pose{i}"], v, n, df)
    ent = ctx.export_entries()
    keys = np.stack([ent["x"], ent["y"], ent["z"]], axis=1).astype(np.int32)
    order = np.lexsort(keys.T[::-1])
    blocks = ctx.block_dict()
    assert len(keys) == int(g["num_blocks"]) and sha(keys[order]) == str(g["keys_sha"])
    assert sha(np.stack([blocks[tuple(int(c) for c in k)] for k in keys[order]])) == str(g["voxels_sha"])
    tris = ctx.extract_mesh().cpu().numpy().reshape(-1, 9).view(np.uint32)
    assert len(tris) == int(g["num_triangles"]) and sha(tris[np.lexsort(tris.T[::-1])]) == str(g["mesh_sha"])
    ctx.icp_reset(True)
    ctx.icp_align(maps[1][0], maps[1][1], maps[0][0], maps[0][1], 5)
    assert np.max(np.abs(ctx.icp_get()[0] - g["icp_delta"])) < 1e-4
    ctx.garbage_collect(L.VH_GC_ALL, 0.03, 1.0)
    assert ctx.stats().lastFreed == int(g["gc_freed"])
    assert sha(np.array(sorted(entries_to_set(ctx.export_entries())), np.int32)) == str(g["gc_keys_sha"])
    b = Context(mod.config(bilateralSigmaSpace=1.5, bilateralSigmaRange=0.03))
    bv, bn, _ = gpu_preprocess(b, g["depth0"])
    assert sha(np.concatenate([bv.cpu().numpy(), bn.cpu().numpy()], axis=1)) == str(g["bilateral_sha"])
