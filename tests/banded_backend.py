"""TEST INFRASTRUCTURE: a backend that splits every surface into horizontal band surfaces (z2d_surface_create_band),
replays every call on all of them and stacks the bands on download -- so any spec scene can be rendered "sharded into
bands" unchanged and compared with the oracle's full-canvas render."""
import numpy as np

from z2d_b200 import abi


class _Banded:
    def __init__(self, fmt, w, h, parts):
        self.fmt, self.w, self.h, self.parts = fmt, w, h, parts  # parts: [(y0, rows, handle)]


class BandedBackend:
    def __init__(self, inner, band_rows=32):
        self.inner, self.band_rows = inner, band_rows

    def surface_create(self, fmt, w, h, initial_px):
        parts, y = [], 0
        while y < h:
            rows = min(self.band_rows, h - y)
            parts.append((y, rows, self.inner.surface_create_band(fmt, w, h, y, rows, initial_px)))
            y += rows
        return _Banded(int(fmt), w, h, parts)

    def surface_destroy(self, hd):
        for _, _, p in hd.parts:
            self.inner.surface_destroy(p)

    def surface_download(self, hd, n):
        bits = abi.format_bits(hd.fmt) if hasattr(abi, "format_bits") else {4: 8, 5: 4, 6: 2, 7: 1}.get(hd.fmt, 32)
        chunks = []
        for _, rows, p in hd.parts:
            nb = (hd.w * rows * bits + 7) // 8
            raw = np.asarray(self.inner.surface_download(p, nb), dtype=np.uint8)
            chunks.append(np.unpackbits(raw, bitorder="little")[: hd.w * rows * bits])
        allbits = np.concatenate(chunks)
        return np.packbits(allbits, bitorder="little")[:n]

    def surface_upload(self, hd, data):
        raise NotImplementedError("banded upload")

    def surface_paint_pixel(self, hd, px):
        for _, _, p in hd.parts:
            self.inner.surface_paint_pixel(p, px)

    def surface_put_pixel(self, hd, x, y, px):
        for _, _, p in hd.parts:
            self.inner.surface_put_pixel(p, x, y, px)

    def surface_param(self, hd, keep):
        raise NotImplementedError("a banded surface cannot be a compositor parameter")

    def fill(self, hd, pat, nodes, n, opts):
        rc = 0
        for _, _, p in hd.parts:
            rc = rc or self.inner.fill(p, pat, nodes, n, opts)
        return rc

    def stroke(self, hd, pat, nodes, n, opts):
        rc = 0
        for _, _, p in hd.parts:
            rc = rc or self.inner.stroke(p, pat, nodes, n, opts)
        return rc

    def composite(self, hd, dst_x, dst_y, ops, n, precision):
        rc = 0
        for _, _, p in hd.parts:
            rc = rc or self.inner.composite(p, dst_x, dst_y, ops, n, precision)
        return rc

    def sync(self):
        self.inner.sync()
