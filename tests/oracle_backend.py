"""TEST INFRASTRUCTURE: backend for z2d_b200.host that runs on the CPU oracle.

Implements the same backend protocol as z2d_b200.cuda_backend.CudaBackend but
over oracle/libz2d_oracle.so and host numpy buffers, so one scene script can be
rendered by the oracle and by the CUDA library and compared byte for byte.
Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from z2d_b200 import abi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libz2d_oracle.so")


class RefSurface(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("format", C.c_uint32), ("width", C.c_int32), ("height", C.c_int32)]


def build_oracle(fast=False):
    target = "libz2d_oracle_fast.so" if fast else "libz2d_oracle.so"
    subprocess.run(["make", "-s", "-C", os.path.join(_ROOT, "oracle"), target], check=True)
    return os.path.join(_ROOT, "oracle", target)


def load_oracle(fast=False):
    path = build_oracle(fast)
    lib = C.CDLL(path)
    P = C.POINTER
    lib.z2d_ref_surface_byte_len.restype = C.c_size_t
    lib.z2d_ref_surface_byte_len.argtypes = [C.c_uint32, C.c_int32, C.c_int32]
    lib.z2d_ref_surface_paint_pixel.restype = C.c_int32
    lib.z2d_ref_surface_paint_pixel.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, P(abi.PixelPOD)]
    lib.z2d_ref_surface_put_pixel.restype = C.c_int32
    lib.z2d_ref_surface_put_pixel.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, P(abi.PixelPOD)]
    lib.z2d_ref_covered_px.restype = C.c_uint64
    lib.z2d_ref_covered_px.argtypes = [C.c_int32]
    lib.z2d_ref_assert_trips.restype = C.c_uint64
    lib.z2d_ref_assert_trips.argtypes = [C.c_int32]
    lib.z2d_ref_fill.restype = C.c_int32
    lib.z2d_ref_fill.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, P(abi.PatternPOD), P(abi.Node), C.c_size_t, P(abi.FillOptsPOD)]
    lib.z2d_ref_stroke.restype = C.c_int32
    lib.z2d_ref_stroke.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, P(abi.PatternPOD), P(abi.Node), C.c_size_t, P(abi.StrokeOptsPOD)]
    lib.z2d_ref_composite.restype = C.c_int32
    lib.z2d_ref_composite.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, P(abi.CompOpPOD), C.c_size_t, C.c_uint32]
    lib.z2d_ref_run_pixel.restype = None
    lib.z2d_ref_run_pixel.argtypes = [C.c_uint32, P(C.c_uint8), P(C.c_uint8), C.c_uint32, P(C.c_uint8)]
    lib.z2d_ref_flatten_fill.restype = C.c_int64
    lib.z2d_ref_flatten_fill.argtypes = [P(abi.Node), C.c_size_t, C.c_double, C.c_double, P(C.c_double), C.c_size_t, P(C.c_double)]
    lib.z2d_ref_flatten_stroke.restype = C.c_int64
    lib.z2d_ref_flatten_stroke.argtypes = [P(abi.Node), C.c_size_t, P(abi.StrokeOptsPOD), C.c_double, P(C.c_double), C.c_size_t, P(C.c_double)]
    lib.z2d_ref_pattern_pixel.restype = None
    lib.z2d_ref_pattern_pixel.argtypes = [P(abi.PatternPOD), C.c_int32, C.c_int32, P(C.c_uint8)]
    return lib


class _Handle:
    def __init__(self, fmt, w, h):
        self.fmt, self.w, self.h = int(fmt), w, h
        self.buf = np.zeros(abi.surface_byte_len(fmt, w, h), dtype=np.uint8)

    @property
    def ptr(self):
        return self.buf.ctypes.data_as(C.c_void_p)


class OracleBackend:
    name = "oracle"

    def __init__(self, fast=False):
        self.lib = load_oracle(fast)

    def surface_create(self, fmt, w, h, initial_px):
        hd = _Handle(fmt, w, h)
        if initial_px is not None:
            abi.check(self.lib.z2d_ref_surface_paint_pixel(hd.ptr, hd.fmt, w, h, C.byref(initial_px.pod())))
        return hd

    def surface_destroy(self, hd):
        pass

    def surface_download(self, hd, n):
        return hd.buf.copy()

    def surface_upload(self, hd, data):
        hd.buf[:] = data

    def surface_paint_pixel(self, hd, px):
        abi.check(self.lib.z2d_ref_surface_paint_pixel(hd.ptr, hd.fmt, hd.w, hd.h, C.byref(px.pod())))

    def surface_downsample(self, hd):
        w, h = C.c_int32(), C.c_int32()
        self.lib.z2d_ref_surface_downsample.restype = C.c_int32
        self.lib.z2d_ref_surface_downsample.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        abi.check(self.lib.z2d_ref_surface_downsample(hd.ptr, hd.fmt, hd.w, hd.h, C.byref(w), C.byref(h)))
        hd.w, hd.h = w.value, h.value
        hd.buf = hd.buf[:abi.surface_byte_len(hd.fmt, hd.w, hd.h)].copy()
        return hd.w, hd.h

    def surface_put_pixel(self, hd, x, y, px):
        abi.check(self.lib.z2d_ref_surface_put_pixel(hd.ptr, hd.fmt, hd.w, hd.h, x, y, C.byref(px.pod())))

    def surface_param(self, hd, keep):
        rs = RefSurface(hd.buf.ctypes.data, hd.fmt, hd.w, hd.h)
        keep.append(rs)
        return C.cast(C.pointer(rs), C.c_void_p)

    def fill(self, hd, pat, nodes, n, opts):
        return self.lib.z2d_ref_fill(hd.ptr, hd.fmt, hd.w, hd.h, C.byref(pat), nodes, n, C.byref(opts))

    def stroke(self, hd, pat, nodes, n, opts):
        return self.lib.z2d_ref_stroke(hd.ptr, hd.fmt, hd.w, hd.h, C.byref(pat), nodes, n, C.byref(opts))

    # text runs: the oracle has no glyph cache -- the outlines are kept on the host and expanded with Transformation.userToDevice
    # (a.ax * x + a.by * y, then + a.tx: Transformation.zig:194-206) into an ordinary node list, as text.show does
    def glyph_cache_add(self, nodes, n):
        if not hasattr(self, "_glyphs"):
            self._glyphs = []
        self._glyphs.append([(nodes[i].tag, tuple(nodes[i].p)) for i in range(n)])
        return len(self._glyphs) - 1

    def fill_glyphs(self, hd, pat, instances, n, opts):
        out = []
        for i in range(n):
            m = instances[i].m
            for tag, p in self._glyphs[instances[i].glyph]:
                q = [0.0] * 6
                for k in range({0: 1, 1: 1, 2: 3, 3: 0}[tag]):
                    x, y = p[2 * k], p[2 * k + 1]
                    dx, dy = m[0] * x + m[1] * y, m[2] * x + m[3] * y
                    q[2 * k], q[2 * k + 1] = dx + m[4], dy + m[5]
                out.append((tag, q))
        if not out:
            return 0
        arr = (abi.Node * len(out))()
        for i, (tag, q) in enumerate(out):
            arr[i].tag = tag
            for k in range(6):
                arr[i].p[k] = q[k]
        return self.fill(hd, pat, arr, len(out), opts)

    def composite(self, hd, dst_x, dst_y, ops, n, precision):
        return self.lib.z2d_ref_composite(hd.ptr, hd.fmt, hd.w, hd.h, dst_x, dst_y, ops, n, precision)

    def sync(self):
        pass


def render_scene(lib, scene, lo=0, hi=None, fmt=int(abi.Format.rgba), keep=None, asserting=None):
    """Replay draws [lo, hi) of a workloads.Scene / FillScene through the CPU oracle; returns the raw surface bytes.
    keep: boolean mask over [lo, hi) of the draws to replay (default all).  asserting: list that receives the indices
    (relative to lo) of the draws on which one of the reference's own debug.asserts would fire."""
    hi = scene.n if hi is None else hi
    buf = np.zeros(scene.width * scene.height * 4, dtype=np.uint8)
    cmds = scene.draw_cmds(0, lo, hi)
    P = C.POINTER
    ptr = buf.ctypes.data_as(C.c_void_p)
    lib.z2d_ref_assert_trips(1)
    for i in range(hi - lo):
        if keep is not None and not keep[i]:
            continue
        pat = C.cast(C.c_void_p(int(cmds["pattern"][i])), P(abi.PatternPOD))
        nodes = C.cast(C.c_void_p(int(cmds["nodes"][i])), P(abi.Node))
        if int(cmds["kind"][i]) == 0:
            rc = lib.z2d_ref_fill(ptr, fmt, scene.width, scene.height, pat, nodes, int(cmds["n_nodes"][i]),
                                  C.cast(C.c_void_p(int(cmds["fill"][i])), P(abi.FillOptsPOD)))
        else:
            rc = lib.z2d_ref_stroke(ptr, fmt, scene.width, scene.height, pat, nodes, int(cmds["n_nodes"][i]),
                                    C.cast(C.c_void_p(int(cmds["stroke"][i])), P(abi.StrokeOptsPOD)))
        assert rc == 0, f"oracle draw {lo + i} failed with {rc}"
        if asserting is not None and lib.z2d_ref_assert_trips(1):
            asserting.append(i)
    return buf
