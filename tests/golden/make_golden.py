"""Captures golden vectors from the UNMODIFIED reference CUDA kernels (GPU box only).

  gpurun -- 'python tests/golden/make_golden.py gpurun_out/reference_c1.npz'
then copy gpurun_out/reference_c1.npz to tests/golden/.  The reference ships no fixtures of its own
(SURVEY.md section 4); these are outputs of its own code (oracle/_ref/libvh_ref.so, built from
/root/reference by oracle/Makefile) on the synthetic inputs of voxelhashing_demo_b200/scenes.py.
Kept small: hashes for the image-sized arrays, full data for tables and a handful of voxel blocks.
"""
import hashlib
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def delta_code(depth):
    """Row-wise first differences as int16: analytic depth compresses ~20x; inverse = cumsum along axis 1."""
    return np.diff(depth.reshape(480, 640).astype(np.int32), axis=1, prepend=0).astype(np.int16)


def main():
    out = sys.argv[1]
    ks = [0, 10, 20]
    worker = ROOT / "tests" / "ref_pin_worker.py"
    with tempfile.TemporaryDirectory() as td:
        f1, f2 = Path(td) / "fuse.npz", Path(td) / "icp.npz"
        subprocess.run([sys.executable, str(worker), str(f1), "5000", ",".join(map(str, ks))], check=True)
        subprocess.run([sys.executable, str(worker), str(f2), "5000", "0,12", "align"], check=True)
        a, b = np.load(f1), np.load(f2)
        g = {"frames": np.array(ks)}
        for i in range(len(ks)):
            g[f"depth_delta{i}"] = delta_code(a[f"depth{i}"])          # the input frame itself (~30 KB delta-coded)
            g[f"verts_sha{i}"] = sha(a[f"verts{i}"])
            g[f"normals_sha{i}"] = sha(a[f"normals{i}"])
            g[f"visible{i}"] = a[f"occupied{i}"][0]
            g[f"heap{i}"] = a[f"occupied{i}"][2]
            g[f"table{i}"] = a[f"table{i}"][:, :3]
            g[f"compact{i}"] = a[f"compact{i}"][:, :3]
        table, blocks = a[f"table{len(ks) - 1}"], a["blocks"]
        pick = np.linspace(0, len(table) - 1, 12).astype(int)
        g["block_keys"] = table[pick, :3]
        g["blocks"] = blocks[pick]
        g["icp_depth_delta0"] = delta_code(b["depth0"])
        g["icp_depth_delta1"] = delta_code(b["depth1"])
        JtJ = b["icp_JtJ"].reshape(6, 6)
        g["icp_JtJ_upper"] = np.array([JtJ[i, j] for i in range(6) for j in range(i, 6)], np.float32)
        g["icp_Jtr"] = b["icp_Jtr"]
        g["icp_err"] = b["icp_err"]
        g["icp_res_sha"] = sha(b["icp_res"])
        g["icp_corr_sha"] = sha(b["icp_corr"])
        g["icp_jac_sha"] = sha(b["icp_jac"])
        g["align_iters"] = b["align_iters"][0]
        g["align_delta"] = b["align_delta"]
        g["align_est"] = b["align_est"]
        np.savez_compressed(out, **g)
    print("golden written:", out, Path(out).stat().st_size, "bytes")


if __name__ == "__main__":
    main()
