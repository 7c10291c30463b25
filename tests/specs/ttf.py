"""HOST-SIDE TEST TOOLING: glyph outlines -> path nodes, so that the reference's text scenes (spec/074, 080, 085 -- 080 is
BASELINE config 1, the z2d logo) can be replayed through painter.fill.  Restates src/Font.zig (table directory, Meta),
src/internal/Glyph.zig (cmap lookup, hmtx/loca, kern + GPOS pair kerning, simple and composite glyf outlines with the
reference's own contour state machine and quadratic->cubic conversion) and src/text.zig (show).  Text layout is outside the
hot path ("next" row 1); only the resulting node list crosses the boundary.  Fonts: tests/golden/fonts (copied
from the reference's spec/test-fonts by tests/golden/import_goldens.py)."""
import os
import struct

from z2d_b200.host import FillOptions, Path, Transformation, nodes_to_array, painter

FONT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden", "fonts")


def font_bytes(name):
    return open(os.path.join(FONT_DIR, name), "rb").read()


class Font:
    def __init__(self, data):  # Font.loadBuffer (Font.zig:49-64): Directory.init + Meta.init (checksums not re-verified)
        self.d = data
        self.dir = dict.fromkeys(("cmap", "glyf", "head", "hhea", "hmtx", "loca", "kern", "GPOS"), 0)
        for i in range(self.u16(4)):
            off = 12 + 16 * i
            tag = data[off:off + 4].decode("latin1")
            if tag in self.dir:
                self.dir[tag] = self.u32(off + 8)
        for t in ("cmap", "glyf", "head", "hhea", "hmtx", "loca"):
            if self.dir[t] == 0:
                raise ValueError("MissingRequiredTable")
        bmp = full = 0
        cmap = self.dir["cmap"]
        for i in range(self.u16(cmap + 2)):  # Font.zig:268-296
            pid, eid, sub = self.u16(cmap + 4 + 8 * i), self.u16(cmap + 6 + 8 * i), self.u32(cmap + 8 + 8 * i) + cmap
            if pid == 0:
                if eid == 3:
                    bmp = sub
                elif eid == 4:
                    full = sub
                else:
                    continue
            if pid == 3:
                if eid == 1:
                    bmp = sub
                elif eid == 10:
                    full = sub
        if not (bmp or full):
            raise ValueError("NoSuitableCmapSubtable")
        self.cmap_full, self.cmap_bmp = full, bmp
        head, hhea = self.dir["head"], self.dir["hhea"]
        self.lsb_is_at_x_zero = bool(self.u16(head + 14) & (2 >> 1))  # Font.zig:318: `& 2 >> 1` parses as `& (2 >> 1)`
        self.long_loca = {0: False, 1: True}[self.u16(head + 50)]
        self.units_per_em = self.u16(head + 18)
        self.advance_width_max = self.u16(hhea + 10)
        self.number_of_hmetrics = self.u16(hhea + 34)

    def u8(self, o):
        return self.d[o]

    def i8(self, o):
        return struct.unpack_from(">b", self.d, o)[0]

    def u16(self, o):
        return struct.unpack_from(">H", self.d, o)[0]

    def i16(self, o):
        return struct.unpack_from(">h", self.d, o)[0]

    def u32(self, o):
        return struct.unpack_from(">I", self.d, o)[0]

    # ---- Glyph.init / byIndex (Glyph.zig:36-93)
    def glyph(self, codepoint):
        index = self._index_full(codepoint) if self.cmap_full else self._index_bmp(codepoint)
        return self.glyph_by_index(index)

    def glyph_by_index(self, index):
        hm, n = self.dir["hmtx"], self.number_of_hmetrics
        if index < n:
            advance, lsb = self.u16(hm + index * 4), self.i16(hm + index * 4 + 2)
        else:
            advance = self.u16(hm + (n - 1) * 4)
            lsb = self.i16(hm + n * 4 + (index - n) * 2)
        offs = []
        for i in range(2):
            if self.long_loca:
                offs.append(self.dir["glyf"] + self.u32(self.dir["loca"] + (index + i) * 4))
            else:
                offs.append(self.dir["glyf"] + self.u16(self.dir["loca"] + (index + i) * 2) * 2)
        return {"index": index, "advance": advance, "lsb": lsb, "outline": None if offs[0] == offs[1] else offs[0]}

    def _index_bmp(self, cp):  # Glyph.zig:95-158
        if cp > 0xFFFF:
            return 0
        t = self.cmap_bmp
        seg_count = self.u16(t + 6) >> 1
        search_range = self.u16(t + 8) >> 1
        entry_selector = self.u16(t + 10)
        range_shift = self.u16(t + 12) >> 1
        end_count = t + 14
        search = end_count
        if cp >= self.u16(search + range_shift * 2):
            search += range_shift * 2
        search -= 2
        while entry_selector > 0:
            search_range >>= 1
            if cp > self.u16(search + search_range * 2):
                search += search_range * 2
            entry_selector -= 1
        search += 2
        item = (search - end_count) >> 1
        start = self.u16(end_count + seg_count * 2 + 2 + 2 * item)
        last = self.u16(end_count + 2 * item)
        if cp < start or cp > last:
            return 0
        offset = self.u16(end_count + seg_count * 6 + 2 + 2 * item)
        if offset == 0:
            return (cp + self.i16(end_count + seg_count * 4 + 2 + 2 * item)) & 0xFFFFFFFF
        return self.u16(offset + (cp - start) * 2 + end_count + seg_count * 6 + 2 + 2 * item)

    def _index_full(self, cp):  # Glyph.zig:160-193
        t = self.cmap_full
        low, high = 0, self.u32(t + 12)
        while low < high:
            mid = low + ((high - low) >> 1)
            o = t + 16 + mid * 12
            start, end = self.u32(o), self.u32(o + 4)
            if cp < start:
                high = mid
            elif cp > end:
                low = mid + 1
            else:
                return self.u32(o + 8) + cp - start
        return 0

    # ---- kerning (Glyph.zig:195-481)
    def kern_advance(self, cur, nxt):
        if self.dir["GPOS"]:
            return self._kern_gpos(cur, nxt)
        if self.dir["kern"]:
            return self._kern_kern(cur, nxt)
        return 0

    def _kern_kern(self, cur, nxt):
        k = self.dir["kern"]
        if self.u16(k + 2) < 1 or self.u16(k + 8) != 1:
            return 0
        lo, hi = 0, self.u16(k + 10) - 1
        needle = (cur << 16 | nxt) & 0xFFFFFFFF
        while lo <= hi:
            m = (lo + hi) >> 1
            straw = self.u32(k + 18 + m * 6)
            if needle < straw:
                hi = m - 1
            elif needle > straw:
                lo = m + 1
            else:
                return self.i16(k + 18 + m * 6 + 4)
        return 0

    def _coverage_index(self, t, glyph):
        fmt = self.u16(t)
        if fmt == 1:
            lo, hi = 0, self.u16(t + 2) - 1
            while lo <= hi:
                m = (lo + hi) >> 1
                straw = self.u16(t + 4 + 2 * m)
                if glyph < straw:
                    hi = m - 1
                elif glyph > straw:
                    lo = m + 1
                else:
                    return m
        elif fmt == 2:
            lo, hi = 0, self.u16(t + 2) - 1
            while lo <= hi:
                m = (lo + hi) >> 1
                r = t + 4 + 6 * m
                s, e = self.u16(r), self.u16(r + 2)
                if glyph < s:
                    hi = m - 1
                elif glyph > e:
                    lo = m + 1
                else:
                    return self.u16(r + 4) + glyph - s
        return -1

    def _glyph_class(self, t, glyph):
        fmt = self.u16(t)
        if fmt == 1:
            start, count = self.u16(t + 2), self.u16(t + 4)
            if start <= glyph < start + count:
                return self.u16(t + 6 + 2 * (glyph - start))
        elif fmt == 2:
            lo, hi = 0, self.u16(t + 2) - 1
            while lo <= hi:
                m = (lo + hi) >> 1
                r = t + 4 + 6 * m
                s, e = self.u16(r), self.u16(r + 2)
                if glyph < s:
                    hi = m - 1
                elif glyph > e:
                    lo = m + 1
                else:
                    return self.u16(r + 4)
        return -1

    def _kern_gpos(self, cur, nxt):
        g = self.dir["GPOS"]
        if self.u16(g) != 1 or self.u16(g + 2) != 0:
            return 0
        lookup_list = g + self.u16(g + 8)
        for i in range(self.u16(lookup_list)):
            lt = lookup_list + self.u16(lookup_list + 2 + 2 * i)
            ltype, n_sub = self.u16(lt), self.u16(lt + 4)
            if ltype not in (2, 9):
                continue
            for sti in range(n_sub):
                so = self.u16(lt + 6 + 2 * sti)
                if ltype == 2:
                    table = lt + so
                else:
                    if self.u16(lt + so + 2) != 2:
                        break
                    table = lt + so + self.u32(lt + so + 4)
                pos_format = self.u16(table)
                ci = self._coverage_index(table + self.u16(table + 2), cur)
                if ci == -1:
                    continue
                vf1, vf2 = self.u16(table + 4), self.u16(table + 6)
                if pos_format == 1:
                    if not (vf1 == 4 and vf2 == 0):
                        return 0
                    pair_set_count = self.u16(table + 8)
                    pvt = table + self.u16(table + 10 + 2 * ci)
                    count = self.u16(pvt)
                    if ci >= pair_set_count:
                        return 0
                    lo, hi = 0, count - 1
                    while lo <= hi:
                        m = (lo + hi) >> 1
                        pv = pvt + 2 + 4 * m
                        straw = self.u16(pv)
                        if nxt < straw:
                            hi = m - 1
                        elif nxt > straw:
                            lo = m + 1
                        else:
                            return self.i16(pv + 2)
                elif pos_format == 2:
                    if not (vf1 == 4 and vf2 == 0):
                        return 0
                    c1 = self._glyph_class(table + self.u16(table + 8), cur)
                    c2 = self._glyph_class(table + self.u16(table + 10), nxt)
                    n1, n2 = self.u16(table + 12), self.u16(table + 14)
                    if c1 < 0 or c1 >= n1 or c2 < 0 or c2 >= n2:
                        return 0
                    return self.i16(table + 16 + 2 * (c1 * n2) + 2 * c2)
                else:
                    return 0
        return 0

    # ---- Glyph.Outline (Glyph.zig:483-869)
    def outline(self, glyph):
        off = glyph["outline"]
        path = Path()
        path.transformation = Transformation().scale(1.0, -1.0).translate(0.0, float(self.units_per_em) * -1.0)
        self._run_outline(path, off)
        return {"nodes": path.nodes, "x_min": self.i16(off + 2)}

    @staticmethod
    def _f2dot14(x):
        hi = x >> 14
        return float(hi - 4 if hi >= 2 else hi) + float(x & 0x3FFF) / 16384.0

    def _quad_to(self, path, cx, cy, tx, ty):  # Glyph.zig:845-869
        if path.current_point is None:
            raise ValueError("MalformedGlyph")
        x0, y0 = path.transformation.device_to_user(*path.current_point)
        x3, y3 = float(tx), float(ty)
        k = 2.0 / 3.0
        x1, y1 = x0 + k * (float(cx) - x0), y0 + k * (float(cy) - y0)
        x2, y2 = x3 + k * (float(cx) - x3), y3 + k * (float(cy) - y3)
        path.curve_to(x1, y1, x2, y2, x3, y3)

    def _run_outline(self, path, off):
        n_contours = self.i16(off)
        pos = off + 10
        if n_contours < 0:  # composite glyph (Glyph.zig:561-632)
            while True:
                flags, index = self.u16(pos), self.u16(pos + 2)
                pos += 4
                if flags & 0x0001:
                    xo, yo = self.i16(pos), self.i16(pos + 2)
                    pos += 4
                else:
                    xo, yo = self.i8(pos), self.i8(pos + 1)
                    pos += 2
                saved = path.transformation
                scaled_offset = bool(flags & 0x0800)
                if not scaled_offset:
                    path.transformation = path.transformation.translate(float(xo), float(yo))
                if flags & 0x0008:
                    s = self._f2dot14(self.u16(pos))
                    pos += 2
                    path.transformation = path.transformation.scale(s, s)
                elif flags & 0x0040:
                    sx, sy = self._f2dot14(self.u16(pos)), self._f2dot14(self.u16(pos + 2))
                    pos += 4
                    path.transformation = path.transformation.scale(sx, sy)
                elif flags & 0x0080:
                    m = Transformation()
                    m.ax, m.cx = self._f2dot14(self.u16(pos)), self._f2dot14(self.u16(pos + 2))
                    m.by, m.dy = self._f2dot14(self.u16(pos + 4)), self._f2dot14(self.u16(pos + 6))
                    pos += 8
                    path.transformation = path.transformation.mul(m)
                if scaled_offset:
                    path.transformation = path.transformation.translate(float(xo), float(yo))
                g = self.glyph_by_index(index)
                if g["outline"] is not None:
                    self._run_outline(path, g["outline"])
                path.transformation = saved
                if not (flags & 0x0020):
                    return
        ends = set()
        outline_len = 0
        for _ in range(n_contours):
            e = self.u16(pos)
            pos += 2
            ends.add(e)
            outline_len = e + 1
        pos += 2 + self.u16(pos)  # instructions
        flags = []
        while len(flags) < outline_len:  # (the reference's flag_idx bookkeeping amounts to "until outline_len flags")
            f = self.u8(pos)
            pos += 1
            flags.append(f)
            if f & 0x08:
                rep = self.u8(pos)
                pos += 1
                flags.extend([f] * rep)
        xs, cur = [], 0
        for i in range(outline_len):
            f = flags[i]
            if f & 0x02:
                cur += self.u8(pos) if f & 0x10 else -self.u8(pos)
                pos += 1
            elif not f & 0x10:
                cur += self.i16(pos)
                pos += 2
            xs.append(cur)
        ys, cur = [], 0
        for i in range(outline_len):
            f = flags[i]
            if f & 0x04:
                cur += self.u8(pos) if f & 0x20 else -self.u8(pos)
                pos += 1
            elif not f & 0x20:
                cur += self.i16(pos)
                pos += 2
            ys.append(cur)
        on = [bool(f & 1) for f in flags]
        state, j = "move_to", 0
        for i in range(outline_len):  # Glyph.zig:723-814
            if state == "move_to":
                if not on[i] or i in ends:
                    continue
                path.move_to(float(xs[i]), float(ys[i]))
                j = i
                state = "on_curve"
            elif state == "on_curve":
                if not on[i]:
                    state = "off_curve"
                else:
                    path.line_to(float(xs[i]), float(ys[i]))
            else:
                if on[i]:
                    self._quad_to(path, xs[i - 1], ys[i - 1], xs[i], ys[i])
                    state = "on_curve"
                else:
                    self._quad_to(path, xs[i - 1], ys[i - 1], (xs[i] + xs[i - 1]) >> 1, (ys[i] + ys[i - 1]) >> 1)
            if i in ends:
                if state == "off_curve":
                    self._quad_to(path, xs[i], ys[i], xs[j], ys[j])
                path.close()
                state = "move_to"


def text_nodes(font, text, x, y, size, transformation=None):
    """text.show (text.zig:73-196) up to the painter.fill call: returns the device-space node list."""
    glyphs, outlines = {}, {}
    path = Path()
    path.transformation = transformation or Transformation()
    cps = [ord(ch) for ch in text]
    scale = size / float(font.units_per_em)
    advance = 0.0

    def get(cp):
        if cp not in glyphs:
            glyphs[cp] = font.glyph(cp)
        return glyphs[cp]

    for idx, cp in enumerate(cps):
        g = get(cp)
        nxt = get(cps[idx + 1]) if idx + 1 < len(cps) else None
        if g["outline"] is not None:
            saved = path.transformation
            if cp not in outlines:
                outlines[cp] = font.outline(g)
            o = outlines[cp]
            pp1 = float(o["x_min"] - g["lsb"]) if not font.lsb_is_at_x_zero else 0.0
            path.transformation = path.transformation.translate(x + advance + pp1, y).scale(scale, scale)
            for n in o["nodes"]:  # Outline.appendToPath
                tag = int(n[0])
                if tag == 0:
                    path.move_to(n[1], n[2])
                elif tag == 1:
                    path.line_to(n[1], n[2])
                elif tag == 2:
                    path.curve_to(*n[1:7])
                else:
                    path.close()
            path.transformation = saved
        kern = float(font.kern_advance(g["index"], nxt["index"])) if nxt is not None else 0.0
        advance += (float(g["advance"] if g["advance"] > 0 else font.advance_width_max) + kern) * scale
    return path.nodes


_GLYPH_CACHES = None  # dict (backend id, font id) -> GlyphCache while use_glyph_cache(True): show_text goes through z2d_fill_glyphs


def use_glyph_cache(on):
    """Route show_text through the backend's glyph cache (z2d_glyph_cache_add / z2d_fill_glyphs) instead of host-built node lists."""
    global _GLYPH_CACHES
    _GLYPH_CACHES = {} if on else None


def show_text(surface, pattern, font, text, x, y, size, fill_opts=None, transformation=None):
    if _GLYPH_CACHES is not None:
        key = (id(surface.backend), id(font))
        if key not in _GLYPH_CACHES:
            _GLYPH_CACHES[key] = GlyphCache(surface.backend, font)
        return show_text_cached(surface, pattern, font, text, x, y, size, _GLYPH_CACHES[key], fill_opts, transformation)
    painter.fill(surface, pattern, text_nodes(font, text, x, y, size, transformation), fill_opts or FillOptions())


class GlyphCache:
    """Per (backend, font): glyph outlines uploaded once (z2d_glyph_cache_add) in the form Outline.appendToPath produces under the
    identity transformation, i.e. replayed through a Path so that Path.close's trailing move_to and the coordinate clamp are in."""

    def __init__(self, backend, font):
        self.backend, self.font, self.ids = backend, font, {}

    def glyph_id(self, cp, outline):
        if cp not in self.ids:
            path = Path()
            for n in outline["nodes"]:  # Outline.appendToPath
                tag = int(n[0])
                if tag == 0:
                    path.move_to(n[1], n[2])
                elif tag == 1:
                    path.line_to(n[1], n[2])
                elif tag == 2:
                    path.curve_to(*n[1:7])
                else:
                    path.close()
            self.ids[cp] = self.backend.glyph_cache_add(nodes_to_array(path.nodes), len(path.nodes))
        return self.ids[cp]


def show_text_cached(surface, pattern, font, text, x, y, size, cache, fill_opts=None, transformation=None):
    """text.show with the glyph outlines resident in the backend: the host only computes one transformation per glyph
    (text.zig:165-172) and the advance; the per-point work runs where the pixels are."""
    glyphs, outlines, instances = {}, {}, []
    base = transformation or Transformation()
    cps = [ord(ch) for ch in text]
    scale = size / float(font.units_per_em)
    advance = 0.0

    def get(cp):
        if cp not in glyphs:
            glyphs[cp] = font.glyph(cp)
        return glyphs[cp]

    for idx, cp in enumerate(cps):
        g = get(cp)
        nxt = get(cps[idx + 1]) if idx + 1 < len(cps) else None
        if g["outline"] is not None:
            if cp not in outlines:
                outlines[cp] = font.outline(g)
            o = outlines[cp]
            pp1 = float(o["x_min"] - g["lsb"]) if not font.lsb_is_at_x_zero else 0.0
            instances.append((cache.glyph_id(cp, o), base.translate(x + advance + pp1, y).scale(scale, scale)))
        kern = float(font.kern_advance(g["index"], nxt["index"])) if nxt is not None else 0.0
        advance += (float(g["advance"] if g["advance"] > 0 else font.advance_width_max) + kern) * scale
    painter.fill_glyphs(surface, pattern, instances, fill_opts or FillOptions())
