"""Product backend: ctypes binding of libz2d_cuda.so (include/z2d_cuda.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is
usable, constructing the backend raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libz2d_cuda.so")

_lib = None


def load_library():
    """Load (building first if nvcc is around and sources are newer) libz2d_cuda.so."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        from . import build as _build
        _build.build()
    if not os.path.exists(SO_PATH):
        raise RuntimeError(f"{SO_PATH} is missing: build it with `python -m z2d_b200.build` (no CPU fallback exists)")
    lib = C.CDLL(os.environ.get("Z2D_CUDA_LIB", SO_PATH))  # Z2D_CUDA_LIB: tuning experiments with variant builds
    P = C.POINTER
    vp = C.c_void_p
    sigs = {
        "z2d_version": (C.c_int32, []),
        "z2d_last_error": (C.c_char_p, [vp]),
        "z2d_ctx_create": (C.c_int32, [C.c_int32, vp, P(vp)]),
        "z2d_ctx_destroy": (None, [vp]),
        "z2d_flush": (C.c_int32, [vp]),
        "z2d_ctx_set_chunk": (C.c_int32, [vp, C.c_uint32]),
        "z2d_sync": (C.c_int32, [vp]),
        "z2d_get_stats": (C.c_int32, [vp, P(abi.StatsPOD)]),
        "z2d_surface_create": (C.c_int32, [vp, C.c_uint32, C.c_int32, C.c_int32, P(abi.PixelPOD), P(vp)]),
        "z2d_surface_create_band": (C.c_int32, [vp, C.c_uint32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, P(abi.PixelPOD), P(vp)]),
        "z2d_surface_band": (C.c_int32, [vp, P(C.c_int32), P(C.c_int32)]),
        "z2d_surface_band_view": (C.c_int32, [vp, C.c_int32, C.c_int32, P(vp)]),
        "z2d_surface_ipc_export": (C.c_int32, [vp, vp]),
        "z2d_surface_open_peer_band": (C.c_int32, [vp, vp, C.c_uint32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, P(vp)]),
        "z2d_surface_destroy": (None, [vp]),
        "z2d_surface_byte_len": (C.c_size_t, [vp]),
        "z2d_surface_width": (C.c_int32, [vp]),
        "z2d_surface_height": (C.c_int32, [vp]),
        "z2d_surface_format": (C.c_uint32, [vp]),
        "z2d_surface_upload": (C.c_int32, [vp, vp, C.c_size_t]),
        "z2d_surface_download": (C.c_int32, [vp, vp, C.c_size_t]),
        "z2d_surface_download_async": (C.c_int32, [vp, vp, C.c_size_t]),
        "z2d_surface_device_ptr": (vp, [vp]),
        "z2d_surface_export_size": (C.c_size_t, [vp, C.c_uint32]),
        "z2d_surface_export": (C.c_int32, [vp, C.c_uint32, vp, C.c_size_t]),
        "z2d_surface_paint_pixel": (C.c_int32, [vp, P(abi.PixelPOD)]),
        "z2d_surface_downsample": (C.c_int32, [vp]),
        "z2d_surface_put_pixel": (C.c_int32, [vp, C.c_int32, C.c_int32, P(abi.PixelPOD)]),
        "z2d_surface_get_pixel": (C.c_int32, [vp, C.c_int32, C.c_int32, P(abi.PixelPOD)]),
        "z2d_fill": (C.c_int32, [vp, vp, P(abi.PatternPOD), P(abi.Node), C.c_size_t, P(abi.FillOptsPOD)]),
        "z2d_stroke": (C.c_int32, [vp, vp, P(abi.PatternPOD), P(abi.Node), C.c_size_t, P(abi.StrokeOptsPOD)]),
        "z2d_composite": (C.c_int32, [vp, vp, C.c_int32, C.c_int32, P(abi.CompOpPOD), C.c_size_t, C.c_uint32]),
        "z2d_submit": (C.c_int32, [vp, P(abi.DrawCmdPOD), C.c_size_t, P(C.c_int32)]),
        "z2d_replay": (C.c_int32, [vp]),
        "z2d_glyph_cache_add": (C.c_int32, [vp, P(abi.Node), C.c_size_t, P(C.c_uint32)]),
        "z2d_fill_glyphs": (C.c_int32, [vp, vp, P(abi.PatternPOD), P(abi.GlyphInstancePOD), C.c_size_t, P(abi.FillOptsPOD)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = ["z2d_version", "z2d_last_error", "z2d_ctx_create", "z2d_ctx_destroy", "z2d_ctx_set_chunk", "z2d_flush", "z2d_sync",
                    "z2d_get_stats", "z2d_surface_create", "z2d_surface_create_band", "z2d_surface_band", "z2d_surface_band_view", "z2d_surface_ipc_export", "z2d_surface_open_peer_band", "z2d_surface_destroy", "z2d_surface_byte_len",
                    "z2d_surface_width", "z2d_surface_height", "z2d_surface_format", "z2d_surface_upload",
                    "z2d_surface_download", "z2d_surface_download_async", "z2d_surface_device_ptr", "z2d_surface_export_size", "z2d_surface_export", "z2d_surface_paint_pixel", "z2d_surface_downsample",
                    "z2d_surface_put_pixel", "z2d_surface_get_pixel", "z2d_fill", "z2d_stroke", "z2d_composite", "z2d_submit", "z2d_replay", "z2d_glyph_cache_add", "z2d_fill_glyphs"]


class CudaBackend:
    """One z2d_ctx on one device.  Surfaces are device resident; draw calls are
    recorded and executed in submission order at the next flush point."""
    name = "cuda"

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        ctx = C.c_void_p()
        rc = self.lib.z2d_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(ctx))
        if rc != abi.OK:
            raise abi.DeviceError(f"z2d_ctx_create(device={device}) failed with status {rc}: no usable CUDA device?")
        self.ctx = ctx
        self.device = device

    def close(self):
        if self.ctx:
            self.lib.z2d_ctx_destroy(self.ctx)
            self.ctx = None

    def _check(self, rc):
        if rc == abi.E_DEVICE:
            raise abi.DeviceError(self.lib.z2d_last_error(self.ctx).decode())
        abi.check(rc)

    # -- surfaces
    def surface_create(self, fmt, w, h, initial_px):
        out = C.c_void_p()
        px = C.byref(initial_px.pod()) if initial_px is not None else None
        self._check(self.lib.z2d_surface_create(self.ctx, int(fmt), w, h, px, C.byref(out)))
        return out

    def surface_create_band(self, fmt, w, canvas_h, y0, rows, initial_px):
        out = C.c_void_p()
        px = C.byref(initial_px.pod()) if initial_px is not None else None
        self._check(self.lib.z2d_surface_create_band(self.ctx, int(fmt), w, canvas_h, y0, rows, px, C.byref(out)))
        return out

    def surface_band_view(self, canvas_hd, y0, rows):
        out = C.c_void_p()
        self._check(self.lib.z2d_surface_band_view(canvas_hd, y0, rows, C.byref(out)))
        return out

    def surface_ipc_export(self, canvas_hd):
        buf = C.create_string_buffer(64)
        self._check(self.lib.z2d_surface_ipc_export(canvas_hd, buf))
        return buf.raw

    def surface_open_peer_band(self, handle, fmt, w, canvas_h, y0, rows):
        out = C.c_void_p()
        self._check(self.lib.z2d_surface_open_peer_band(self.ctx, C.create_string_buffer(handle, 64), int(fmt), w, canvas_h, y0, rows, C.byref(out)))
        return out

    def surface_destroy(self, hd):
        self.lib.z2d_surface_destroy(hd)

    def surface_download(self, hd, n):
        buf = np.empty(n, dtype=np.uint8)
        self._check(self.lib.z2d_surface_download(hd, buf.ctypes.data_as(C.c_void_p), n))
        return buf

    def surface_export(self, hd, srgb=False, filter_byte=False):
        """PNG scanline bytes (export_png.zig:150-373) produced on the device; (rows, row_bytes) uint8."""
        flags = (abi.EXPORT_SRGB if srgb else 0) | (abi.EXPORT_FILTER_BYTE if filter_byte else 0)
        n = self.lib.z2d_surface_export_size(hd, flags)
        rows = self.lib.z2d_surface_height(hd)
        buf = np.empty(n, dtype=np.uint8)
        self._check(self.lib.z2d_surface_export(hd, flags, buf.ctypes.data_as(C.c_void_p), n))
        return buf.reshape(rows, n // max(1, rows))

    def surface_upload(self, hd, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        self._check(self.lib.z2d_surface_upload(hd, data.ctypes.data_as(C.c_void_p), data.size))

    def surface_paint_pixel(self, hd, px):
        self._check(self.lib.z2d_surface_paint_pixel(hd, C.byref(px.pod())))

    def surface_put_pixel(self, hd, x, y, px):
        self._check(self.lib.z2d_surface_put_pixel(hd, x, y, C.byref(px.pod())))

    def surface_downsample(self, hd):
        self._check(self.lib.z2d_surface_downsample(hd))
        return self.lib.z2d_surface_width(hd), self.lib.z2d_surface_height(hd)

    def surface_get_pixel(self, hd, x, y):
        """(format, r, g, b, a) as stored, or None where Surface.getPixel returns null."""
        pod = abi.PixelPOD()
        rc = self.lib.z2d_surface_get_pixel(hd, x, y, C.byref(pod))
        if rc == 1:
            return None
        self._check(rc)
        return (pod.format, pod.r, pod.g, pod.b, pod.a)

    def surface_param(self, hd, keep):
        return hd

    def surface_device_ptr(self, hd):
        return self.lib.z2d_surface_device_ptr(hd)

    # -- drawing
    def fill(self, hd, pat, nodes, n, opts):
        rc = self.lib.z2d_fill(self.ctx, hd, C.byref(pat), nodes, n, C.byref(opts))
        if rc == abi.E_DEVICE:
            raise abi.DeviceError(self.lib.z2d_last_error(self.ctx).decode())
        return rc

    def stroke(self, hd, pat, nodes, n, opts):
        rc = self.lib.z2d_stroke(self.ctx, hd, C.byref(pat), nodes, n, C.byref(opts))
        if rc == abi.E_DEVICE:
            raise abi.DeviceError(self.lib.z2d_last_error(self.ctx).decode())
        return rc

    def composite(self, hd, dst_x, dst_y, ops, n, precision):
        rc = self.lib.z2d_composite(self.ctx, hd, dst_x, dst_y, ops, n, precision)
        if rc == abi.E_DEVICE:
            raise abi.DeviceError(self.lib.z2d_last_error(self.ctx).decode())
        return rc

    def glyph_cache_add(self, nodes, n):
        out = C.c_uint32()
        self._check(self.lib.z2d_glyph_cache_add(self.ctx, nodes, n, C.byref(out)))
        return out.value

    def fill_glyphs(self, hd, pat, instances, n, opts):
        rc = self.lib.z2d_fill_glyphs(self.ctx, hd, C.byref(pat), instances, n, C.byref(opts))
        if rc == abi.E_DEVICE:
            raise abi.DeviceError(self.lib.z2d_last_error(self.ctx).decode())
        return rc

    def submit(self, cmds, n):
        """z2d_submit: `cmds` is a ctypes array of abi.DrawCmdPOD."""
        self._check(self.lib.z2d_submit(self.ctx, cmds, n, None))

    def replay(self):
        self._check(self.lib.z2d_replay(self.ctx))

    def set_chunk(self, max_draws):
        self._check(self.lib.z2d_ctx_set_chunk(self.ctx, int(max_draws)))

    def flush(self):
        self._check(self.lib.z2d_flush(self.ctx))

    def sync(self):
        self._check(self.lib.z2d_sync(self.ctx))

    def stats(self):
        s = abi.StatsPOD()
        self._check(self.lib.z2d_get_stats(self.ctx, C.byref(s)))
        return {k: getattr(s, k) for k, _ in abi.StatsPOD._fields_ if not k.startswith("_")}
