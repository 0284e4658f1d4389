"""Pins the Fixed policy's DEFINITION against drift: outputs of the CPU oracle on two small stored
frames, as hashes.  Not reference-derived (the reference has no Fixed policy) -- the reference pin is reference_c1.npz;
this file only makes sure that a change to the Fixed arithmetic is a decision and not an accident: the GPU kernels are
checked bit-for-bit against the oracle on the GPU box, this fixture checks the oracle against its own history on CPU.

  python tests/golden/make_fixed_golden.py        (rewrites tests/golden/fixed_small.npz)
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def config(**kw):
    from conftest import small_cfg
    from voxelhashing_demo_b200 import POLICY_FIXED

    base = dict(policy=POLICY_FIXED, numBuckets=1009, numVoxelBlocks=4096, truncation=0.06, overflowSlots=1024, icpNormalThres=0.8)
    base.update(kw)
    return small_cfg(**base)


def run(depths, poses):
    """Everything the fixture pins, from the stored frames only."""
    from oracle import binding as ob

    out = {}
    cfg = config()
    t = ob.OracleTable(cfg)
    maps = []
    for d, p in zip(depths, poses):
        v, n, df = t.preprocess(d)
        maps.append((v, n))
        t.fuse_frame(p, v, df)
    ent = t.entries()
    order = np.lexsort(ent[:, :3].T[::-1])
    out["keys_sha"] = sha(ent[order, :3])
    out["voxels_sha"] = sha(np.stack([t.block(*e[:3]) for e in ent[order]]))
    out["num_blocks"] = len(ent)
    tris = t.extract_mesh().reshape(-1, 9).view(np.uint32)
    out["mesh_sha"] = sha(tris[np.lexsort(tris.T[::-1])])
    out["num_triangles"] = len(tris)
    _, _, delta = ob.icp_align(cfg, maps[1][0], maps[1][1], maps[0][0], maps[0][1], 5)
    out["icp_delta"] = np.asarray(delta, np.float32)
    rv, rn = t.raycast(poses[1])
    out["raycast_sha"] = sha(np.concatenate([rv, rn], axis=1))
    freed = t.garbage_collect(scope=1, sdf_threshold=0.03, weight_decay=1.0)
    out["gc_freed"] = freed
    out["gc_keys_sha"] = sha(np.array(sorted(map(tuple, t.entries()[:, :3].tolist())), np.int32))
    tb = ob.OracleTable(config(bilateralSigmaSpace=1.5, bilateralSigmaRange=0.03))
    bv, bn, _ = tb.preprocess(depths[0])
    out["bilateral_sha"] = sha(np.concatenate([bv, bn], axis=1))
    return out


def main():
    from conftest import render
    from voxelhashing_demo_b200 import scenes

    cfg = config()
    poses = [scenes.trajectory_C2(k).astype(np.float32) for k in (3, 11)]
    rng = np.random.default_rng(7)
    depths = []
    for p in poses:
        d = render(cfg, scenes.scene_S1T(), p)
        d = np.where(d > 0, (d.astype(np.int32) + rng.integers(-4, 5, d.shape)).clip(1, 65535), 0).astype(np.uint16)
        d[40:44, 30:60] = 0
        depths.append(d)
    g = {f"depth{i}": d for i, d in enumerate(depths)}
    g.update({f"pose{i}": p for i, p in enumerate(poses)})
    g.update(run(depths, poses))
    np.savez_compressed(ROOT / "tests" / "golden" / "fixed_small.npz", **g)
    print({k: (v if not isinstance(v, np.ndarray) or v.size < 20 else v.shape) for k, v in g.items() if not k.startswith("depth")})


if __name__ == "__main__":
    main()
