"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The hot path shards in exactly two places (BASELINE.json north_star):
  * fusion  -- the block-coordinate hash space is partitioned, owner(block) = mix(block) mod P; every
               rank scans the whole frame, inserts / compacts / integrates only the blocks it owns
               (the ownership test runs inside k_alloc after the warp-level de-duplication), so no
               voxel ever crosses NVLink.  Exchange: one broadcast of the u16 depth frame per frame.
  * ICP     -- frame-to-frame ICP needs no model data: image rows are split over the ranks, each rank
               reduces its 27 (+2) partial sums on the device, ONE all-reduce of 32 floats per
               iteration, every rank solves the identical 6x6 system (bit-identical pose on all ranks,
               no second broadcast).
Everything here is stream-ordered; the only host synchronisation is what NCCL itself needs.
"""
from __future__ import annotations

import numpy as np


def row_range(rank: int, world: int, height: int) -> tuple[int, int]:
    """Contiguous block of image rows owned by `rank` for the ICP reduction (balanced to +-1 row)."""
    base, rem = divmod(height, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def owner_of(x: int, y: int, z: int, parts: int) -> int:
    """Host restatement of ownerMix (csrc/vh_device.cuh): which rank owns block (x, y, z)."""
    m = 0xFFFFFFFF
    u = ((x * 0x9E3779B1) ^ (y * 0x85EBCA77) ^ (z * 0xC2B2AE3D)) & m
    u ^= u >> 16
    u = (u * 0x7FEB352D) & m
    u ^= u >> 15
    u = (u * 0x846CA68B) & m
    u ^= u >> 16
    return u % parts


class PartitionedTracker:
    """Per-rank driver of one partition: broadcast -> preprocess -> split ICP + all-reduce -> fuse."""

    def __init__(self, ctx, rank: int, world: int, iterations: int | None = None, group=None, overlap: bool = True,
                 tuning: tuple[int, int] | None = None):
        """overlap: fuse frame k on a second stream beside the broadcast / pre-processing / tracking of frame k+1
        (frame-to-frame ICP does not read the model); results are identical, the model lags the pose until flush().
        tuning = (Align CTAs, SMs the persistent integrate grid leaves free), None = automatic: see _auto_tuning."""
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.iterations = iterations or ctx.cfg.icpIterations
        cfg = ctx.cfg
        n = cfg.width * cfg.height
        dt = torch.uint16 if hasattr(torch, "uint16") else torch.int16
        # two landing buffers: the copy + broadcast of frame k+1 run on their own stream while frame k is being tracked
        self._depths = [torch.zeros(n, dtype=dt, device="cuda"), torch.zeros(n, dtype=dt, device="cuda")]
        self.depth = self._depths[0]
        # Everything the tracker enqueues goes to its OWN stream: the overlapped schedule of the native pipeline (CUDA graph
        # capture, the Align / fusion side streams) is not available on the legacy default stream, which is what a caller
        # who never set a stream hands in.  push(input_ready=False) / reset order it behind the caller's current stream,
        # flush() orders the caller's current stream behind it.
        self.stream = torch.cuda.Stream()
        self._bcast_stream = torch.cuda.Stream()
        self._ev_arrived = [torch.cuda.Event(), torch.cuda.Event()]
        self._ev_consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self._pushed = 0
        self.maps = [ctx.new_maps(), ctx.new_maps()]
        self.sys = torch.zeros(32, dtype=torch.float32, device="cuda")
        self.d_pose = torch.zeros(16, dtype=torch.float32, device="cuda")
        self.rows = row_range(rank, world, cfg.height)
        self.frame = 0
        self.launches = 0
        self.fused = False
        self.overlap = bool(overlap)
        self._fuse_stream = torch.cuda.Stream() if self.overlap else None
        self._ev_pose = torch.cuda.Event() if self.overlap else None
        self._ev_fused = torch.cuda.Event() if self.overlap else None
        self._fuse_pending = False
        if tuning is None:
            tuning = self._auto_tuning(cfg, world)
        if tuning is not None and overlap:
            ctx.set_tuning(int(tuning[0]), int(tuning[1]))
        self.tuning = tuning
        self._frame_ptrs = (0, 0)                            # device addresses of the (previous, latest) frame as this rank holds them
        self._dist = None                                   # vh_dist handle (C ABI transport: CUDA IPC, no NCCL on the data path)
        if world > 1:
            import os

            # Frame transport of this Python tracker.  Default "nccl": symmetric-memory mailboxes + NCCL frame broadcast on
            # its own stream (the broadcast of frame k+1 is issued as soon as its landing buffer is free, so the frame is
            # pre-processed DURING Align(k)).  "ipc": the library's own transport (vh_dist_*, what a C++ host uses) -- the
            # same results, but its push is ordered behind the pose of the previous frame and arrives when the persistent
            # integrate grid already holds the SMs: 8 GPUs 2 071 vs 3 015 frames/s (r2).
            if os.environ.get("VH_DIST_TRANSPORT", "nccl") == "ipc":
                self._setup_ipc_transport()
            if self._dist is None:
                self._setup_peer_exchange()
        # With the exchange fused into the ICP kernel the whole frame is stream-ordered device work, so it goes through
        # the native frame pipeline (CUDA graphs; the host enqueues ~5 calls per frame instead of ~30 -- the Python
        # loop below needed 0.4 ms of host time per frame and capped the 4- and 8-GPU runs).  The NCCL-per-iteration
        # fallback cannot be captured that way and keeps the Python loop.
        self.pipe = None
        if self.fused or world == 1:
            from .fusion import FramePipeline

            self.pipe = FramePipeline(ctx, iterations=self.iterations, mode=FramePipeline.FRAME_TO_FRAME, use_graph=True,
                                      overlap=self.overlap)

    @staticmethod
    def _auto_tuning(cfg, world: int):
        """The Align grid (one 512-thread CTA per SM, the whole register file of each) and the persistent integrate grid
        cannot share an SM.  With a few image rows per rank the Align is a chain of exchange latencies, not work: a SMALL
        grid on its own SMs costs it little and lets tracking(k+1) run beside fusion(k) instead of after it.  Measured on C4
        at 2 mm (r2, frames/s, default grid -> split): 4 GPUs 1 430 -> 1 740 (40 CTAs + 40 reserved SMs), 8 GPUs 2 319 ->
        3 058 (38 + 38; Align 161 -> 236 us, integrate 205 -> 231 us, but side by side).  With 2 ranks the fusion is five times
        longer than the Align and giving up a quarter of the SMs costs more than the overlap returns."""
        if world < 4:
            return None
        rows = (cfg.height + world - 1) // world
        ctas = max(16, min(40, (rows * cfg.width + 3071) // 3072))
        return ctas, ctas

    def _setup_ipc_transport(self):
        """The library's own transport (vh_dist_*, csrc/vh_dist.cu): one CUDA IPC region per rank holding the ICP mailboxes
        and the frame landing buffers; torch.distributed only carries the 64-byte handles once.  Rank 0's kernel stores
        every frame straight into the peers' landing buffers -- no NCCL broadcast, no stream hand-off."""
        import ctypes as C
        import sys

        torch, dist = self.torch, self.dist
        lib = self.ctx.lib
        ok, why = 0, ""
        handle = C.c_void_p()
        try:
            if lib.vh_dist_create(self.ctx._h, self.rank, self.world, C.byref(handle)) != 0:
                raise RuntimeError("vh_dist_create failed")
            hb = int(lib.vh_dist_handle_bytes())
            mine = (C.c_ubyte * hb)()
            if lib.vh_dist_export(handle, mine) != 0:
                raise RuntimeError("vh_dist_export failed")
            ok = 1
        except Exception as e:  # noqa: BLE001
            why = str(e)
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)           # collective decision, as below
        if int(flag.item()) == 1:
            t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device="cuda")
            allh = [torch.zeros_like(t) for _ in range(self.world)]
            dist.all_gather(allh, t, group=self.group)
            blob = b"".join(bytes(x.cpu().numpy().tobytes()) for x in allh)
            buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
            rc = lib.vh_dist_connect(handle, buf)
            flag = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            self._dist = handle
            self.fused = True
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
        else:
            if handle:
                lib.vh_dist_destroy(handle)
            print(f"[rank {self.rank}] vh_dist transport unavailable" + (f" ({why})" if why else "") + "; trying symmetric memory + NCCL broadcast",
                  file=sys.stderr)

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 -- interpreter shutdown
            pass

    def close(self):
        if getattr(self, "_dist", None) is not None:
            if self.pipe is not None:
                self.pipe.close()
                self.pipe = None
            self.torch.cuda.synchronize()
            self.ctx.lib.vh_dist_destroy(self._dist)
            self._dist = None

    def _setup_peer_exchange(self):
        """Symmetric (peer-mapped) exchange regions so the 32-float all-reduce runs INSIDE the ICP kernel's
        epilogue over NVLink instead of as a separate NCCL launch per iteration.  Falls back to NCCL if the
        symmetric-memory rendezvous is unavailable."""
        torch, dist = self.torch, self.dist
        import sys

        ok, ptrs, why = 0, None, ""
        try:
            import torch.distributed._symmetric_memory as symm

            n = (self.ctx.peer_bytes() + 3) // 4
            self._xbuf = symm.empty(n, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
            self._xbuf.zero_()
            hdl = symm.rendezvous(self._xbuf, self.group if self.group is not None else dist.group.WORLD)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            self._xhdl = hdl
            ok = 1
        except Exception as e:  # noqa: BLE001
            why = str(e)
        # The decision is COLLECTIVE: a rank that fell back to NCCL while its peers spin inside the fused kernel waiting
        # for its mailbox stores would hang the job, so every rank uses the fused exchange only if every rank can.
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        self.fused = bool(int(flag.item()))
        if self.fused:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)                   # every region is zeroed before anyone can store into it
            self.ctx.set_peers(self.rank, self.world, ptrs)
        else:
            print(f"[rank {self.rank}] fused peer exchange unavailable on at least one rank"
                  + (f" (here: {why})" if why else "") + "; ICP all-reduce through NCCL", file=sys.stderr)

    def flush(self):
        """Order the current stream behind the fusion of the latest pushed frame (no-op without overlap)."""
        cur = self.torch.cuda.current_stream()
        with self.torch.cuda.stream(self.stream):
            self._flush_own()
        cur.wait_stream(self.stream)

    def _flush_own(self):
        if self.pipe is not None:
            self.pipe.flush()
            return
        if self._fuse_pending:
            self.torch.cuda.current_stream().wait_event(self._ev_fused)
            self._fuse_pending = False

    def reset(self, pose):
        self.stream.wait_stream(self.torch.cuda.current_stream())
        with self.torch.cuda.stream(self.stream):
            self._flush_own()
            if self.pipe is not None:
                self.pipe.reset(np.ascontiguousarray(pose, dtype=np.float32))
                self.frame = 0
                return
            self.d_pose.copy_(self.torch.from_numpy(np.ascontiguousarray(pose, dtype=np.float32).reshape(16)))
            self.ctx.icp_reset(True)
            self.frame = 0

    def push(self, d_depth=None, input_ready: bool = False):
        """One frame.  Rank 0 passes the depth image (device tensor, or pinned host tensor for the end-to-end
        path); the others pass None.  input_ready=True promises that the image is already complete in memory (a
        resident sequence, a pinned host buffer filled earlier): the copy + broadcast then start right away on their
        own stream and overlap the tracking of the previous frame.  Otherwise they are ordered behind everything
        enqueued on the current stream so far, which is always safe."""
        if not input_ready:
            self.stream.wait_stream(self.torch.cuda.current_stream())
        with self.torch.cuda.stream(self.stream):
            self._push(d_depth, input_ready)

    def _push(self, d_depth, input_ready):
        ctx, dist, torch = self.ctx, self.dist, self.torch
        main = torch.cuda.current_stream()                   # = self.stream
        if self._dist is not None:
            # the library's transport: rank 0's kernel stores the frame (device or pinned host source) into every rank's
            # landing buffer; the pipeline pre-processes it as soon as its arrival event fires
            import ctypes as C

            frame, ready = C.c_void_p(), C.c_void_p()
            src = None
            if self.rank == 0:
                src = d_depth.data_ptr()
            # The push is ordered behind this stream (which is ordered behind the pose of the previous frame).  Letting it run
            # ahead (input_ready) was tried and withdrawn: with the landing-slot flow control spinning inside resident kernels a
            # 4-GPU run with the default scheduling stopped making progress (r2), so the verified ordering stays.
            L_check = self.ctx.lib.vh_dist_broadcast_frame(self._dist, src, C.byref(frame), C.byref(ready), main.cuda_stream)
            if L_check != 0:
                raise RuntimeError("vh_dist_broadcast_frame failed")
            self._frame_ptrs = (self._frame_ptrs[1], int(frame.value))
            self._pushed += 1
            l0 = self.pipe.launches()
            self.pipe.push_device_ready(int(frame.value), int(ready.value))
            if self.ctx.lib.vh_dist_frame_consumed(self._dist, main.cuda_stream) != 0:
                raise RuntimeError("vh_dist_frame_consumed failed")
            self.launches += self.pipe.launches() - l0
            self.frame += 1
            return
        slot = self._pushed & 1
        self.depth = self._depths[slot]
        bs = self._bcast_stream
        if self._pushed >= 2 and input_ready:
            bs.wait_event(self._ev_consumed[slot])           # the pre-processing of frame k-2 has read this buffer
        else:
            bs.wait_stream(main)                             # behind the producer of d_depth (and all earlier work)
        with torch.cuda.stream(bs):
            if self.rank == 0:
                if d_depth.is_cuda:
                    d_depth.record_stream(bs)
                self.depth.copy_(d_depth.view(self.depth.dtype), non_blocking=True)    # device or pinned host source
            if self.world > 1:
                dist.broadcast(self.depth.view(torch.uint8), src=0, group=self.group)   # raw bytes over NVLink / NVSwitch
            self._ev_arrived[slot].record(bs)
        self._pushed += 1
        self._frame_ptrs = (self._frame_ptrs[1], int(self.depth.data_ptr()))
        if self.pipe is not None:
            l0 = self.pipe.launches()
            # the landing buffer is complete when the broadcast has arrived -- an event, not this stream: the pre-processing of
            # frame k+1 then runs beside the Align of frame k instead of behind it
            self.pipe.push_device_ready(self.depth, self._ev_arrived[slot])
            self._ev_consumed[slot].record(main)             # conservative: recorded behind the whole push, not only the pre-processing
            self.launches += self.pipe.launches() - l0
            self.frame += 1
            return
        main.wait_event(self._ev_arrived[slot])
        par = self.frame & 1
        v, n, df = self.maps[par]
        pv, pn, _ = self.maps[1 - par]
        ctx.preprocess(self.depth, v, n, df)
        self._ev_consumed[slot].record(main)
        self.launches += 1
        if self.frame > 0:
            if self.fused or self.world == 1:
                # one kernel per iteration: reduce over my rows + all-reduce over NVLink peers + solve
                ctx.icp_align_rows(v, n, pv, pn, self.rows[0], self.rows[1], self.iterations)
                self.launches += self.iterations
            else:
                for _ in range(self.iterations):
                    ctx.icp_reduce(v, n, pv, pn, self.rows[0], self.rows[1], self.sys)
                    dist.all_reduce(self.sys, group=self.group)                           # 32 floats
                    ctx.icp_solve(self.sys)
                    self.launches += 2
        # the frame constants and per-frame counters change now: the previous frame's fusion must be through
        self._flush_own()
        if self.frame > 0:
            ctx.pose_compose(self.d_pose, self.d_pose)        # T_k = T_{k-1} * delta, also publishes the frame pose
        else:
            ctx.set_pose_device(self.d_pose)
        self.launches += 1
        if self.overlap:
            main = self.torch.cuda.current_stream()
            self._ev_pose.record(main)
            fs = self._fuse_stream
            fs.wait_event(self._ev_pose)
            ctx.alloc_blocks(v, n, fs)
            ctx.compact(fs)
            ctx.integrate_depthf(df, fs)
            self._ev_fused.record(fs)
            self._fuse_pending = True
        else:
            ctx.alloc_blocks(v, n)
            ctx.compact()
            ctx.integrate_depthf(df)
        self.launches += 3
        self.frame += 1

    def pose(self) -> np.ndarray:
        with self.torch.cuda.stream(self.stream):
            if self.pipe is not None:
                p = self.pipe.pose()
                self.torch.cuda.synchronize()
                return p
            self.torch.cuda.synchronize()
            return self.d_pose.cpu().numpy().reshape(4, 4)

    def pose_async(self, h_pose_pinned):
        """Stream-ordered D2H of the current pose into a pinned host tensor of 16 floats (the e2e read-back); valid after
        flush() + a synchronisation of the caller's stream."""
        with self.torch.cuda.stream(self.stream):
            if self.pipe is not None:
                self.pipe.pose_async(h_pose_pinned)
            else:
                h_pose_pinned.copy_(self.d_pose, non_blocking=True)

    def last_frames(self) -> tuple[int, int]:
        """Device addresses of the (previous, latest) raw depth frames in this rank's landing buffers."""
        return self._frame_ptrs

    def last_depthf(self):
        """Dense metric depth of the latest frame: device address (native pipeline) or tensor (Python loop)."""
        if self.pipe is not None:
            return self.pipe.depthf_ptr()
        return self.maps[(self.frame - 1) & 1][2]
