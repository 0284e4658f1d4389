"""Runs the UNMODIFIED reference CUDA kernels (oracle/_ref/libvh_ref.so) on synthetic frames and saves
what they produce.  One process per table geometry: the reference keeps its table in process-global
__constant__/host state (VoxelUtils.cu:23-26), so it can be initialised only once.

usage: ref_pin_worker.py OUT.npz NUM_BUCKETS FRAME_K[,FRAME_K...] [align] [single]
"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    out, nb, ks = sys.argv[1], int(sys.argv[2]), [int(k) for k in sys.argv[3].split(",")]
    do_align = "align" in sys.argv[4:]
    passes = 1 if "single" in sys.argv[4:] else 4
    import torch

    from oracle import binding as ob
    from voxelhashing_demo_b200 import Config, scenes

    # the reference prints from host code and from kernels (VoxelUtils.cu:433-435,452): silence fd 1
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        cfg = Config(numBuckets=nb, numVoxelBlocks=4000)
        ref = ob.ref_lib()
        rc = ref.ref_init(cfg.numBuckets, cfg.bucketSize, cfg.numVoxelBlocks, 0.0, 0.0)
        assert rc == 0, rc
        K, Kinv = cfg.K(), cfg.Kinv()
        ref.ref_set_intrinsic(K.ctypes.data, Kinv.ctypes.data)
        n = 640 * 480
        res = {}
        maps = []
        for i, k in enumerate(ks):
            pose = scenes.trajectory_C2(k).astype(np.float32)
            depth = scenes.render_depth((scenes.scene_S1T() if do_align else scenes.scene_S1()), pose, 640, 480, cfg.fx, cfg.fy, cfg.cx, cfg.cy)
            d = torch.from_numpy(depth.reshape(-1).copy()).cuda()
            v = torch.zeros((n, 4), device="cuda")
            nm = torch.zeros((n, 4), device="cuda")
            ref.ref_preprocess(v.data_ptr(), nm.data_ptr(), d.data_ptr())
            p = np.ascontiguousarray(pose.reshape(16))
            if passes == 1:
                occ = ref.ref_integrate(p.ctypes.data, v.data_ptr(), nm.data_ptr())     # SDF_Hashtable::integrate as is
            else:
                # allocation to convergence (one insert per bucket per pass, quirk Q4), then compact + integrate
                for _ in range(passes):
                    ref.ref_stage_begin(p.ctypes.data)
                    ref.ref_stage_alloc(v.data_ptr(), nm.data_ptr())
                occ = ref.ref_stage_compact()
                ref.ref_stage_integrate(v.data_ptr())
            torch.cuda.synchronize()
            maps.append((v, nm))
            slots = ref.ref_num_slots()
            table = np.zeros((slots, 5), np.int32)
            assert ref.ref_export_table(table.ctypes.data, slots) == slots
            comp = np.zeros((max(occ, 1), 5), np.int32)
            nc = ref.ref_export_compact(comp.ctypes.data, occ)
            res[f"depth{i}"] = depth
            res[f"verts{i}"] = v.cpu().numpy()
            res[f"normals{i}"] = nm.cpu().numpy()
            res[f"occupied{i}"] = np.array([occ, nc, ref.ref_heap_counter()])
            res[f"table{i}"] = table[table[:, 3] != -1]
            res[f"compact{i}"] = comp[:nc]
        alloc = res[f"table{len(ks) - 1}"]
        blocks = np.zeros((len(alloc), 512, 2), np.float32)
        for j, e in enumerate(alloc):
            ref.ref_export_block(int(e[3]), blocks[j].ctypes.data)
        res["blocks"] = blocks
        if do_align and len(maps) >= 2:
            (tv, tn), (iv, inn) = maps[0], maps[-1]
            delta = np.eye(4, dtype=np.float32).reshape(16).copy()
            corr = torch.zeros((n, 4), device="cuda")
            corrN = torch.zeros((n, 4), device="cuda")
            rs = torch.zeros(n, device="cuda")
            err = ref.ref_correspond(iv.data_ptr(), tv.data_ptr(), tn.data_ptr(), delta.ctypes.data, corr.data_ptr(), corrN.data_ptr(), rs.data_ptr())
            JtJ, Jtr = np.zeros(36, np.float32), np.zeros(6, np.float32)
            jac = torch.zeros((n, 6), device="cuda")
            ref.ref_build_system(iv.data_ptr(), JtJ.ctypes.data, Jtr.ctypes.data, jac.data_ptr())
            torch.cuda.synchronize()
            res.update(icp_err=np.array([err]), icp_corr=corr.cpu().numpy(), icp_corrN=corrN.cpu().numpy(), icp_res=rs.cpu().numpy(),
                       icp_JtJ=JtJ, icp_Jtr=Jtr, icp_jac=jac.cpu().numpy())
            est = np.zeros(6, np.float32)
            delta = np.eye(4, dtype=np.float32).reshape(16).copy()
            its = ref.ref_align(iv.data_ptr(), tv.data_ptr(), tn.data_ptr(), 20, ob.ref_solve_callback(), est.ctypes.data, delta.ctypes.data)
            res.update(align_iters=np.array([its]), align_est=est, align_delta=delta.reshape(4, 4))
        np.savez_compressed(out, **res)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
    print("ok")


if __name__ == "__main__":
    main()
