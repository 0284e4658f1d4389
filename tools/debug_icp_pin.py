"""GPU-box debug helper: where do the reference's ICP residuals differ from the oracle's?"""
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import binding as ob  # noqa: E402
from voxelhashing_demo_b200 import Config  # noqa: E402

with tempfile.TemporaryDirectory() as td:
    out = Path(td) / "icp.npz"
    subprocess.run([sys.executable, str(ROOT / "tests" / "ref_pin_worker.py"), str(out), "5000", "0,12", "align"], check=True)
    ref = np.load(out)
    cfg = Config(numVoxelBlocks=4000)
    ot = ob.OracleTable(cfg)
    tv, tn, _ = ot.preprocess(ref["depth0"])
    iv, inn, _ = ot.preprocess(ref["depth1"])
    print("verts equal", np.array_equal(ref["verts0"].view(np.uint32), tv.view(np.uint32)), np.array_equal(ref["verts1"].view(np.uint32), iv.view(np.uint32)))
    print("normals equal", np.array_equal(ref["normals0"].view(np.uint32), tn.view(np.uint32)))
    err, corr, corrN, res = ob.find_correspondences(cfg, iv, None, tv, tn, np.eye(4, dtype=np.float32))
    r = ref["icp_res"]
    bad = np.nonzero(r.view(np.uint32) != res.view(np.uint32))[0]
    print("res mismatches", len(bad), "of", len(r))
    for i in bad[:10]:
        print(i, i % 640, i // 640, r[i], res[i], ref["icp_corr"][i], corr[i], ref["icp_corrN"][i], corrN[i], iv[i])
    cb = np.nonzero((ref["icp_corr"].view(np.uint32) != corr.view(np.uint32)).any(1))[0]
    print("corr mismatches", len(cb))
    nb = np.nonzero((ref["icp_corrN"].view(np.uint32) != corrN.view(np.uint32)).any(1))[0]
    print("corrN mismatches", len(nb))
    print("ref err", ref["icp_err"], "oracle", err, "sum", res.astype(np.float64).sum())
