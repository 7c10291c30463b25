"""One very large canvas split into horizontal bands, one band per GPU (second row).

Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/band_canvas.py
Every rank records the SAME draw calls (the edge list is small next to the pixels) on a band surface holding rows
[rank*H/N, (rank+1)*H/N) of an HxH RGBA8 canvas; the only exchange step is the gather of the finished bands to rank 0
(NCCL over NVLink, torch.distributed.gather on zero-copy views of the surfaces' device memory).  Rank 0 then checks the
stacked canvas against its own single-GPU render of the whole canvas (--verify) and prints one JSON line."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from z2d_b200 import abi, workloads  # noqa: E402
from z2d_b200.cuda_backend import CudaBackend  # noqa: E402
from z2d_b200.host import Pixel, Surface  # noqa: E402


class _DevView:  # zero-copy torch view of a surface's device memory
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--paths", type=int, default=100_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--verify", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()  # an explicit stream shared by the library, the copies and the collective (0 would make
    torch.cuda.set_stream(stream)  # the library create a private one)
    cb = CudaBackend(local, stream=stream.cuda_stream)
    H = args.size
    rows = H // world
    assert rows % 16 == 0 and rows * world == H
    scene = workloads.cubic_paths_scene(args.paths, H, seed=0x7A326402, r_log2=(5.0, 9.0))
    band = Surface(abi.Format.rgba, H, H, None, cb, band=(rank * rows, rows))
    cmds = scene.draw_cmds(band.handle)
    cmds_p = cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD))
    view = torch.as_tensor(_DevView(cb.surface_device_ptr(band.handle), band.byte_len()), device="cuda")
    full = torch.empty(H * H * 4, dtype=torch.uint8, device="cuda") if rank == 0 else None
    parts = list(full.split(rows * H * 4)) if rank == 0 else None
    zero = Pixel.rgba(0, 0, 0, 0)

    def step():
        band.paint_pixel(zero)
        cb.submit(cmds_p, scene.n)
        cb.flush()  # everything is enqueued on the (shared) stream
        if world > 1:
            dist.gather(view, parts, dst=0)
        else:
            full.copy_(view)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = (time.perf_counter() - t0) / args.steps * 1e3
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps, wall], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ok = None
    if args.verify and rank == 0:
        ref = Surface(abi.Format.rgba, H, H, None, cb)
        c2 = scene.draw_cmds(ref.handle)
        cb.submit(c2.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), scene.n)
        cb.sync()
        rv = torch.as_tensor(_DevView(cb.surface_device_ptr(ref.handle), ref.byte_len()), device="cuda")
        ok = bool(torch.equal(rv, full))
    if rank == 0:
        print(json.dumps({"workload": f"{H}x{H} RGBA8 canvas, {args.paths} cubic paths, {world} horizontal bands of {rows} rows, "
                                      "gather to rank 0 over NCCL", "n_gpus": world, "ms_per_canvas_device": float(ms[0]),
                          "ms_per_canvas_wall": float(ms[1]), "canvas_bytes": H * H * 4, "gathered_bytes": (world - 1) * rows * H * 4,
                          "bands_equal_full_render": ok}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
