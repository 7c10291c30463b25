"""The other BASELINE.json configurations behind `bench.py --workload ...` (bench.py itself holds config 2, the headline).

  c4   8192^2 compositor sweep: destination format x source x dither x operator x precision, one K5 launch per cell,
       device-timed; GB/s on the algorithmic bytes of SURVEY 8(d) (2 x bpp(dst) per pixel that reads dst, 1 x bpp for
       write-only operators; generated sources cost 0) and the fraction of the measured HBM peak.
       Plus the `Context.fill` form of the same thing: the 8192^2 rectangle through painter.fill / the tile kernel.
  c5   batch of 1024^2 mixed scenes (fills, strokes, gradient fills) sharded scene s -> rank s mod N, strong scaling,
       device-timed and end to end, with the CPU restatement on every host core beside it.
"""
import ctypes as C
import json
import os
import time

import numpy as np

from z2d_b200 import abi, sharding, workloads
from z2d_b200.abi import DitherType, Format, Interp, Operator, Precision
from z2d_b200.host import Dither, Gradient, Operation, Param, Pixel, Surface, SurfaceCompositor

BITS = {Format.rgba: 32, Format.rgb: 32, Format.alpha8: 8, Format.alpha4: 4, Format.alpha2: 2, Format.alpha1: 1}
FLOAT_ONLY = {Operator.color_dodge, Operator.color_burn, Operator.soft_light, Operator.hue, Operator.saturation, Operator.color,
              Operator.luminosity}
WRITE_ONLY = {Operator.clear, Operator.src}  # operators that never read dst (1 x bpp)
OP_CLASS = {
    "write_only": [Operator.clear, Operator.src],
    "porter_duff": [Operator.src_over, Operator.dst_over, Operator.src_in, Operator.dst_in, Operator.src_out, Operator.dst_out, Operator.src_atop,
                    Operator.dst_atop, Operator.xor, Operator.plus, Operator.dst],
    "separable_blend": [Operator.multiply, Operator.screen, Operator.overlay, Operator.darken, Operator.lighten, Operator.hard_light,
                        Operator.difference, Operator.exclusion],
    "float_only_separable": [Operator.color_dodge, Operator.color_burn, Operator.soft_light],
    "non_separable": [Operator.hue, Operator.saturation, Operator.color, Operator.luminosity],
}
CLASS_OF = {op: k for k, ops in OP_CLASS.items() for op in ops}


def _stops(g):
    g.add_stop(0.0, {"rgba": (1, 0, 0, 1)})
    g.add_stop(0.5, {"rgba": (0, 1, 0, 0.5)})
    g.add_stop(1.0, {"rgba": (0, 0, 1, 1)})
    return g


def c4_sources(n, bpc):
    lin = lambda **kw: _stops(Gradient.linear(0, 0, n, n, **kw))  # noqa: E731
    rad = lambda: _stops(Gradient.radial(n / 2, n / 2, 0, n / 2, n / 2, n / 2))  # noqa: E731
    con = lambda: _stops(Gradient.conic(n / 2, n / 2, 0))  # noqa: E731
    out = {("pixel", "none"): Param.pixel(Pixel.rgba(90, 40, 10, 128))}
    for name, mk in (("linear", lin), ("radial", rad), ("conic", con)):
        out[(name, "none")] = Param.gradient(mk())
        out[(name, "bayer")] = Param.dither(Dither(DitherType.bayer, mk(), bpc))
        out[(name, "blue_noise")] = Param.dither(Dither(DitherType.blue_noise, mk(), bpc))
    out[("linear_srgb", "none")] = Param.gradient(lin(method=Interp.srgb))
    out[("linear_hsl", "none")] = Param.gradient(lin(method=Interp.hsl))
    return out


def _prefill(fmt, n):
    """hash32 content in the raw layout of `fmt` (RGBA premultiplied) so that dst-dependent operators are exercised."""
    n_words = n * n * BITS[fmt] // 32
    i = np.arange(n_words, dtype=np.uint64)
    h = (i * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x85EBCA77)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    raw = h.astype(np.uint32)
    if fmt == Format.rgba:
        a = (raw >> 24).astype(np.uint32)
        out = (a << 24)
        for sh in (0, 8, 16):
            out |= ((((raw >> sh) & 255) * a) // 255) << sh
        raw = out
    elif fmt == Format.rgb:
        raw &= np.uint32(0x00FFFFFF)
    return raw.view(np.uint8)


def run_c4(cb, peak_gbs, n=8192, reps=3, full=True):
    """Every cell of the config-4 sweep as one full-surface SurfaceCompositor.run, timed with CUDA events on the library's stream."""
    import torch
    cells = []
    for fmt in (Format.rgba, Format.rgb, Format.alpha8, Format.alpha4, Format.alpha2, Format.alpha1):
        sfc = Surface(fmt, n, n, None, cb)
        sfc.upload(_prefill(fmt, n))
        bpp = BITS[fmt] / 8.0
        bpc = 8 if BITS[fmt] >= 8 else BITS[fmt]
        sources = c4_sources(n, bpc)
        for (sname, dname), prm in sources.items():
            if sname == "pixel" and fmt == Format.rgba:
                ops = list(Operator)  # all 28 on the headline cell
            elif full:
                ops = [Operator.src, Operator.src_over, Operator.multiply, Operator.soft_light, Operator.hue]
            else:
                ops = [Operator.src_over]
            for op in ops:
                for prec in (Precision.integer, Precision.float):
                    if prec == Precision.integer and op in FLOAT_ONLY:
                        continue
                    run = lambda: SurfaceCompositor.run(sfc, 0, 0, [Operation(op, src=prm)], precision=prec)  # noqa: E731
                    run()
                    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    ev0.record()
                    for _ in range(reps):
                        run()
                    ev1.record()
                    torch.cuda.synchronize()
                    ms = ev0.elapsed_time(ev1) / reps
                    algo = n * n * bpp * (1 if op in WRITE_ONLY else 2)
                    gbs = algo / (ms * 1e-3) / 1e9
                    cells.append({"format": fmt.name, "source": sname, "dither": dname, "op": op.name, "op_class": CLASS_OF[op],
                                  "precision": prec.name, "ms": round(ms, 4), "gbs": round(gbs, 1), "frac": round(gbs / peak_gbs, 4),
                                  "gpix_s": round(n * n / (ms * 1e-3) / 1e9, 2)})
        sfc.deinit()
    return cells


def c4_summary(cells):
    """min / median fraction per (format, source kind, precision) group + the worst cells."""
    groups = {}
    for c in cells:
        src = "pixel" if c["source"] == "pixel" else ("gradient" if c["dither"] == "none" else "dither")
        groups.setdefault((c["format"], src, c["precision"]), []).append(c["frac"])
    rows = [{"format": f, "source": s, "precision": p, "cells": len(v), "frac_min": min(v), "frac_median": float(np.median(v))}
            for (f, s, p), v in sorted(groups.items())]
    return rows


def c4_markdown(cells, peak):
    lines = [f"| format | source | dither | op | precision | ms | GB/s | frac of {peak:.1f} | Gpix/s |", "|---|---|---|---|---|---|---|---|---|"]
    for c in cells:
        lines.append(f"| {c['format']} | {c['source']} | {c['dither']} | {c['op']} | {c['precision']} | {c['ms']} | {c['gbs']} | {c['frac']} | {c['gpix_s']} |")
    return "\n".join(lines)


def run_c4_fill(cb, peak_gbs, n=8192):
    """The `Context.fill` form of config 4: painter.fill of the rectangle (0,0)-(n,n), default AA, through the
    tile kernel (K4) instead of the surface compositor.  Device-timed from the call to the end of the batch."""
    import torch
    from z2d_b200.host import FillOptions, Path, Pattern, painter
    out = []
    sfc = Surface(Format.rgba, n, n, None, cb)
    sfc.upload(_prefill(Format.rgba, n))
    path = Path()
    path.move_to(0, 0); path.line_to(n, 0); path.line_to(n, n); path.line_to(0, n); path.close()
    srcs = c4_sources(n, 8)
    pats = {"pixel": Pattern.opaque(Pixel.rgba(90, 40, 10, 128)), "linear": Pattern.gradient(srcs[("linear", "none")].value),
            "radial": Pattern.gradient(srcs[("radial", "none")].value), "linear+bayer": Pattern.dither(srcs[("linear", "bayer")].value)}
    for name, pat in pats.items():
        for op in (Operator.src_over, Operator.multiply):
            run = lambda: (painter.fill(sfc, pat, path.nodes, FillOptions(operator=op)), cb.flush())  # noqa: E731
            run()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(3):
                run()
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / 3
            gbs = n * n * 8.0 / (ms * 1e-3) / 1e9
            out.append({"source": name, "op": op.name, "ms": round(ms, 4), "gbs": round(gbs, 1), "frac": round(gbs / peak_gbs, 4),
                        "gpix_s": round(n * n / (ms * 1e-3) / 1e9, 2)})
    sfc.deinit()
    return out


# ------------------------------------------------------------------------------------------------ config 5
C5_DESC = ("BASELINE config 5: batch of {n} scenes of {s}x{s} RGBA8, 64 ordered draw calls each (32 cubic-path fills r 8-256, 24 strokes with "
           "round/miter joins, round caps, half dashed, 8 gradient fills: 2 linear, 3 radial, 3 conic over rectangles / ellipses), seed = "
           "0x7A326405 + scene")


def _c5_cpu_worker(args):
    """Renders scenes with the CPU restatement in a worker process; returns (busy seconds, covered px, scenes)."""
    indices, size = args
    from tests.oracle_backend import load_oracle, render_scene
    lib = load_oracle(fast=True)
    scenes = [workloads.mixed_scene(s, size) for s in indices]
    lib.z2d_ref_covered_px(1)
    t0 = time.perf_counter()
    for sc in scenes:
        render_scene(lib, sc)
    return time.perf_counter() - t0, int(lib.z2d_ref_covered_px(1)), len(indices)


def cpu_c5(n_scenes_total, size, cores=None, per_core=2):
    """Independent scenes are the one place the single-threaded reference can use every host core: one process per core,
    `per_core` scenes each; throughput = scenes / slowest worker."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    cores = min(cores, n_scenes_total)
    per_core = max(1, min(per_core, n_scenes_total // cores))
    jobs = [(list(range(w * per_core, (w + 1) * per_core)), size) for w in range(cores)]
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_c5_cpu_worker, jobs)
    busy = max(r[0] for r in res)
    px = sum(r[1] for r in res)
    n = sum(r[2] for r in res)
    return {"value": px / busy / 1e6, "unit": "Mpix/s", "scenes_per_s": n / busy, "cores": cores, "kind": "port",
            "sample": f"{n} of {n_scenes_total} scenes, {per_core} per process on {cores} processes (one per host core), slowest worker {busy:.1f} s; "
                      "C++ restatement of z2d's CPU path (oracle/), not z2d itself"}


def reference_c5(args):
    cpu = cpu_c5(args.scenes, 1024)
    line = {"impl": "reference", "metric": "filled+composited Mpix/s", "value": cpu["value"], "unit": "Mpix/s", "n_gpus": args.gpus, "steps": 1,
            "warmup": 0, "ms_per_step": None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "scenes_per_s": cpu["scenes_per_s"], "config": {"workload": C5_DESC.format(n=args.scenes, s=1024), "sample": cpu["sample"]},
            "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_c5(cb, rank, world, dist, n_scenes, size, steps, warmup, with_cpu=False, group=64):
    """Config 5 on `world` ranks: scene s -> rank s mod world, every rank owns its scenes' surfaces and command batches, no
    data-path collective.  Device-timed (batch resident, z2d_replay) and end to end (host node arrays in, host pixels out:
    groups of `group` scenes are submitted and their surfaces read back asynchronously into pinned memory while the next
    group renders).  Times are the max over ranks; strong scaling (the batch is fixed)."""
    import torch
    lib = cb.lib
    my = sharding.scenes_of_rank(n_scenes, world, rank)
    t0 = time.perf_counter()
    scenes = [workloads.mixed_scene(s, size) for s in my]
    gen_s = time.perf_counter() - t0
    sfcs = [Surface(Format.rgba, size, size, None, cb) for _ in my]
    parts = [sc.draw_cmds(sf.handle) for sc, sf in zip(scenes, sfcs)]
    cmds = np.concatenate(parts)
    offs = np.concatenate([[0], np.cumsum([len(p) for p in parts])])
    nbytes = size * size * 4
    host = torch.empty(len(my) * nbytes, dtype=torch.uint8, pin_memory=True)
    hbase = host.data_ptr()
    P = C.POINTER(abi.DrawCmdPOD)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def e2e_step():
        for g0 in range(0, len(my), group):
            g1 = min(g0 + group, len(my))
            sub = cmds[offs[g0]:offs[g1]]
            cb.submit(sub.ctypes.data_as(P), len(sub))
            for i in range(g0, g1):
                cb._check(lib.z2d_surface_download_async(sfcs[i].handle, C.c_void_p(hbase + i * nbytes), nbytes))
        cb.sync()

    cb.set_chunk(32768)
    for _ in range(warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3

    # device-resident: the rank's whole share as ONE batch, replayed
    cb.set_chunk(0)
    cb.submit(cmds.ctypes.data_as(P), len(cmds))
    cb.sync()
    st = cb.stats()
    h2d = st["h2d_bytes"]
    for _ in range(max(warmup, 2)):
        cb.replay()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        cb.replay()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    st = cb.stats()
    cb.set_chunk(32768)
    (dev_ms, e2e_ms), (covered, draws, n_mine, h2d_all) = sharding.reduce_timing(
        [dev_ms, e2e_ms], [st["covered_px"], st["draws"], len(my), h2d], device="cuda", dist=dist)
    stages = {k: st[k] for k in ("ms_flatten", "ms_bin", "ms_lists", "ms_raster", "ms_total")}
    for sf in sfcs:
        sf.deinit()
    del host
    if rank != 0:
        return None
    out = {
        "workload": C5_DESC.format(n=n_scenes, s=size), "parallelism": f"scene s -> rank s mod {world}; no data-path collective", "scaling": "strong",
        "n_gpus": world, "scenes": int(n_mine), "steps": steps, "warmup": max(warmup, 2), "ms_per_step": dev_ms / steps,
        "scenes_per_s": n_mine * steps / (dev_ms * 1e-3), "mpix_s": covered * steps / (dev_ms * 1e-3) / 1e6,
        "canvas_mpix_s": n_mine * size * size * steps / (dev_ms * 1e-3) / 1e6, "draws_per_s": draws * steps / (dev_ms * 1e-3),
        "e2e": {"value": covered * steps / (e2e_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": e2e_ms / steps,
                "scenes_per_s": n_mine * steps / (e2e_ms * 1e-3), "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(n_mine * nbytes),
                "how": f"z2d_submit per group of {group} scenes from host node arrays + z2d_surface_download_async of every surface into pinned "
                       "memory (overlaps the next group), one z2d_sync per step"},
        "rank0_stages_ms": stages, "gpu_launches": int(st["kernel_launches"] * steps), "scene_generation_s_rank0": gen_s,
        "note": "surfaces are not cleared between steps (4096 clears would be 4096 launches): every step composites the same draws over the "
                "previous output, identical work",
    }
    if with_cpu:
        out["cpu_baseline"] = cpu_c5(n_scenes, size)
    return out


# ------------------------------------------------------------------------------------------------ band canvas
def run_band(cb, rank, world, dist, size, n_paths, steps, warmup, verify=False):
    """One very large canvas (second row): rank r renders rows [r*H/N, (r+1)*H/N) of an HxH RGBA8 canvas that lives
    on rank 0.  Every rank records the SAME draw calls on a band VIEW of that canvas -- rank 0 over its own memory, the others
    over a CUDA-IPC peer mapping -- so the raster kernel's tile write-back lands in rank 0's canvas over NVLink and there is no
    separate gather.  Device-timed per canvas (max over ranks), strong scaling."""
    import torch
    H = size
    rows = (H // world) & ~15
    assert rows * world == H, "canvas height must split into bands of whole tile rows"
    scene = workloads.cubic_paths_scene(n_paths, H, seed=0x7A326402, r_log2=(5.0, 9.0))
    canvas = Surface(Format.rgba, H, H, None, cb) if rank == 0 else None
    if world > 1:
        obj = [canvas.ipc_export() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        band = canvas.band_view(0, rows) if rank == 0 else Surface.open_peer_band(obj[0], Format.rgba, H, H, rank * rows, rows, cb)
    else:
        band = canvas.band_view(0, rows)
    cmds = scene.draw_cmds(band.handle)
    P = C.POINTER(abi.DrawCmdPOD)
    zero = Pixel.rgba(0, 0, 0, 0)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        band.paint_pixel(zero)
        cb.submit(cmds.ctypes.data_as(P), scene.n)
        cb.flush()

    for _ in range(warmup):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    st = cb.stats()
    (dev_ms, wall_ms), (covered,) = sharding.reduce_timing([ev0.elapsed_time(ev1), wall_ms], [st["covered_px"]], device="cuda", dist=dist)
    out = None
    if rank == 0:
        out = {"workload": f"band canvas: {H}x{H} RGBA8 ({H * H * 4 / 2**30:.2f} GiB), {n_paths} cubic paths (r 32-512), one band of {rows} rows per GPU; "
                           "all ranks replay the same calls, bands are views of rank 0's canvas (peer mapping over NVLink): the gather is K4's write-back",
               "n_gpus": world, "steps": steps, "ms_per_canvas": dev_ms / steps, "wall_ms_per_canvas": wall_ms / steps, "scaling": "strong",
               "canvas_mpix_s": H * H / (dev_ms / steps * 1e-3) / 1e6}
        if verify:  # the stacked canvas equals a single-GPU render of the whole canvas
            got = canvas.download().copy()
            ref = Surface(Format.rgba, H, H, None, cb)
            c2 = scene.draw_cmds(ref.handle)
            cb.submit(c2.ctypes.data_as(P), scene.n)
            out["equal_to_single_gpu_render"] = bool(np.array_equal(got, ref.download()))
            ref.deinit()
    if dist is not None:
        dist.barrier()  # the peers keep their mappings until rank 0 has read the canvas
    band.deinit()
    if canvas is not None:
        canvas.deinit()
    return out


# ------------------------------------------------------------------------------------------------ config 1
class _Recorder:
    """Backend wrapper that records the (pattern, nodes, opts) of every fill so the scene can be replayed without text layout."""

    def __init__(self, inner):
        self.inner, self.calls = inner, []

    def __getattr__(self, name):
        return getattr(self.inner, name)

    def fill(self, hd, pat, nodes, n, opts):
        self.calls.append((pat, nodes, n, opts))
        return self.inner.fill(hd, pat, nodes, n, opts)


def _logo_loop(backend, reps, warm=3):
    from tests import specs
    from z2d_b200.abi import AntiAliasMode
    rec = _Recorder(backend)
    sfc = specs.PATH_SCENES["080_fill_z2d_logo"](specs.bind(rec), AntiAliasMode.default)
    zero = Pixel.rgba(0, 0, 0, 0)

    def scene():
        sfc.paint_pixel(zero)
        for pat, nodes, n, opts in rec.calls:
            backend.fill(sfc.handle, pat, nodes, n, opts)
        backend.sync()

    for _ in range(warm):
        scene()
    t0 = time.perf_counter()
    for _ in range(reps):
        scene()
    dt = time.perf_counter() - t0
    return dt / reps * 1e6, sum(c[2] for c in rec.calls), len(rec.calls), sfc


def run_c1(cb, reps=1000, with_cpu=True):
    """BASELINE config 1 (spec/080_fill_z2d_logo): 601x172 RGBA8, five non-zero fills (3 polygons, 2 text runs), opaque source,
    default AA.  Latency bound: timed as back-to-back scenes, each = clear + 5 painter.fill calls + wait for completion (glyph
    outlines -> nodes is host-side text layout outside the path, so the node lists are built once)."""
    gpu_us, nodes, calls, sfc = _logo_loop(cb, reps)
    st = cb.stats()
    out = {"workload": "BASELINE config 1: spec/080_fill_z2d_logo, 601x172 RGBA8, 5 fills (3 polygons + 2 text runs), opaque source, src_over, default AA; "
                       "one scene = clear + 5 painter.fill + sync, node lists prebuilt",
           "fills_per_scene": calls, "nodes_per_scene": nodes, "us_per_scene": gpu_us, "scenes_per_s": 1e6 / gpu_us, "paths_per_s": calls * 1e6 / gpu_us,
           "bbox_mpix_s": st["region_px"] / gpu_us, "covered_mpix_s": st["covered_px"] / gpu_us, "kernel_launches_per_scene": st["kernel_launches"]}
    if with_cpu:
        from tests.oracle_backend import OracleBackend
        cpu_us, _, _, _ = _logo_loop(OracleBackend(fast=True), 50)
        out["cpu_baseline"] = {"value": cpu_us, "unit": "us/scene", "cores": 1, "kind": "port",
                               "sample": "50 scenes; C++ restatement of z2d's CPU path (oracle/), not z2d itself"}
    return out


# ------------------------------------------------------------------------------------------------ config 3 (block of the default line)
def run_c3(cb, n_strokes=50_000, steps=5):
    """BASELINE config 3 as a block of the default line: device step of the 2048^2 stroke workload (resident batch, z2d_replay).
    `python bench.py --workload c3` is the full line (e2e, CPU baseline, clocks)."""
    import torch
    scene = workloads.stroke_paths_scene(n_strokes, 2048, seed=0x7A326403)
    sfc = Surface(Format.rgba, 2048, 2048, None, cb)
    cmds = scene.draw_cmds(sfc.handle)
    cb.set_chunk(0)
    cb.submit(cmds.ctypes.data_as(C.POINTER(abi.DrawCmdPOD)), scene.n)
    cb.sync()
    zero = Pixel.rgba(0, 0, 0, 0)
    for _ in range(2):
        sfc.paint_pixel(zero)
        cb.replay()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(steps):
        sfc.paint_pixel(zero)
        cb.replay()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    st = cb.stats()
    cb.set_chunk(32768)
    sfc.deinit()
    return {"workload": f"BASELINE config 3: 2048x2048 RGBA8, {n_strokes} strokes (round / miter joins, round caps, half dashed)", "ms_per_step": ms,
            "strokes_per_s": scene.n / (ms * 1e-3), "mpix_s": st["covered_px"] / (ms * 1e-3) / 1e6,
            "stages_ms": {k: st[k] for k in ("ms_flatten", "ms_bin", "ms_lists", "ms_raster", "ms_total")},
            "counters": {k: int(st[k]) for k in ("edges", "band_edges", "tile_pairs", "crossings", "covered_px")}}
