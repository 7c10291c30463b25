"""Deterministic synthetic workloads of BASELINE.json (host-side scene generators).

Everything is produced directly as arrays of the C-ABI PODs (numpy views over
ctypes memory), so a scene of 10^5 draw calls is one `z2d_submit`.
All coordinates are multiples of 1/16 px (exact in f64).
"""
import ctypes as C

import numpy as np

from . import abi
from .abi import AntiAliasMode, FillRule, Format, NodeTag, Operator, PatternKind, Precision


def _np_dtype(ct):
    """numpy dtype with the exact memory layout of a ctypes type."""
    if issubclass(ct, C.Structure):
        names, formats, offsets = [], [], []
        for name, ft in ct._fields_:
            names.append(name)
            formats.append(_np_dtype(ft))
            offsets.append(getattr(ct, name).offset)
        return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": C.sizeof(ct)})
    if issubclass(ct, C.Array):
        return np.dtype((_np_dtype(ct._type_), (ct._length_,)))
    if issubclass(ct, (C._Pointer, C.c_void_p, C.c_char_p)):
        return np.dtype("<u8")
    return np.dtype(ct)


NODE_DT = _np_dtype(abi.Node)
PATTERN_DT = _np_dtype(abi.PatternPOD)
DRAWCMD_DT = _np_dtype(abi.DrawCmdPOD)
STROKEOPTS_DT = _np_dtype(abi.StrokeOptsPOD)
FILLOPTS_DT = _np_dtype(abi.FillOptsPOD)
GRADIENT_DT = _np_dtype(abi.GradientPOD)
STOP_DT = _np_dtype(abi.StopPOD)


class FillScene:
    """An ordered list of painter.fill calls on one surface, as C-ABI arrays."""

    def __init__(self, width, height, nodes, node_off, patterns, fill_opts, opt_index):
        self.width, self.height = width, height
        self.nodes = nodes            # NODE_DT array
        self.node_off = node_off      # int64[n+1]
        self.patterns = patterns      # PATTERN_DT array
        self.fill_opts = fill_opts    # ctypes array of FillOptsPOD (few distinct)
        self.opt_index = opt_index    # which fill_opts entry each draw uses
        self.n = len(patterns)

    def draw_cmds(self, surface_handle, lo=0, hi=None):
        """DrawCmdPOD array (numpy view) for draws [lo, hi) targeting `surface_handle`."""
        hi = self.n if hi is None else hi
        cmds = np.zeros(hi - lo, dtype=DRAWCMD_DT)
        cmds["kind"] = 0
        cmds["surface"] = surface_handle.value if hasattr(surface_handle, "value") else int(surface_handle)
        cmds["pattern"] = self.patterns.ctypes.data + np.arange(lo, hi, dtype=np.uint64) * PATTERN_DT.itemsize
        cmds["nodes"] = self.nodes.ctypes.data + self.node_off[lo:hi].astype(np.uint64) * NODE_DT.itemsize
        cmds["n_nodes"] = (self.node_off[lo + 1:hi + 1] - self.node_off[lo:hi]).astype(np.uint64)
        cmds["fill"] = C.addressof(self.fill_opts) + self.opt_index[lo:hi].astype(np.uint64) * C.sizeof(abi.FillOptsPOD)
        return cmds


def _q16(v):
    return np.round(v * 16.0) / 16.0


def cubic_paths_scene(n_paths=100_000, size=4096, seed=0x7A326402, r_log2=(3.0, 7.0), aa=AntiAliasMode.default,
                      op=Operator.src_over):
    """BASELINE config 2: `size`^2 RGBA8 canvas, n random closed paths of 4 cubic
    Beziers around a centre, radius 2^U[3,7] px, alternating non-zero / even-odd,
    translucent pre-multiplied colours, src_over, integer pipeline, default AA."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(0, size, n_paths)
    cy = rng.uniform(0, size, n_paths)
    r = np.exp2(rng.uniform(r_log2[0], r_log2[1], n_paths))
    ang = (2 * np.pi * np.arange(4) / 4)[None, :] + rng.uniform(-0.3, 0.3, (n_paths, 4))
    rad = r[:, None] * rng.uniform(0.6, 1.0, (n_paths, 4))
    ax = _q16(cx[:, None] + rad * np.cos(ang))
    ay = _q16(cy[:, None] + rad * np.sin(ang))
    # control points: anchor +- r * U[0, 0.5] per axis (c1 off the start anchor, c2 off the end anchor)
    c1x = _q16(ax + r[:, None] * rng.uniform(-0.5, 0.5, (n_paths, 4)))
    c1y = _q16(ay + r[:, None] * rng.uniform(-0.5, 0.5, (n_paths, 4)))
    nx, ny = np.roll(ax, -1, axis=1), np.roll(ay, -1, axis=1)
    c2x = _q16(nx + r[:, None] * rng.uniform(-0.5, 0.5, (n_paths, 4)))
    c2y = _q16(ny + r[:, None] * rng.uniform(-0.5, 0.5, (n_paths, 4)))

    per = 7  # move_to, 4 x curve_to, close_path, trailing move_to (Path.zig:467-476)
    nodes = np.zeros(n_paths * per, dtype=NODE_DT).reshape(n_paths, per)
    nodes["tag"][:, 0] = int(NodeTag.move_to)
    nodes["p"][:, 0, 0] = ax[:, 0]
    nodes["p"][:, 0, 1] = ay[:, 0]
    for k in range(4):
        nodes["tag"][:, 1 + k] = int(NodeTag.curve_to)
        nodes["p"][:, 1 + k, 0] = c1x[:, k]
        nodes["p"][:, 1 + k, 1] = c1y[:, k]
        nodes["p"][:, 1 + k, 2] = c2x[:, k]
        nodes["p"][:, 1 + k, 3] = c2y[:, k]
        nodes["p"][:, 1 + k, 4] = nx[:, k]
        nodes["p"][:, 1 + k, 5] = ny[:, k]
    nodes["tag"][:, 5] = int(NodeTag.close_path)
    nodes["tag"][:, 6] = int(NodeTag.move_to)
    nodes["p"][:, 6, 0] = ax[:, 0]
    nodes["p"][:, 6, 1] = ay[:, 0]
    nodes = np.ascontiguousarray(nodes.reshape(-1))
    node_off = np.arange(n_paths + 1, dtype=np.int64) * per

    a = rng.integers(64, 256, n_paths)
    rgb = (rng.uniform(0, 1, (n_paths, 3)) * a[:, None]).astype(np.int64)  # pre-multiplied: c <= a
    patterns = np.zeros(n_paths, dtype=PATTERN_DT)
    patterns["kind"] = int(PatternKind.opaque)
    patterns["pixel"]["format"] = int(Format.rgba)
    patterns["pixel"]["r"] = rgb[:, 0]
    patterns["pixel"]["g"] = rgb[:, 1]
    patterns["pixel"]["b"] = rgb[:, 2]
    patterns["pixel"]["a"] = a

    opts = (abi.FillOptsPOD * 2)()
    for i, rule in enumerate((FillRule.non_zero, FillRule.even_odd)):
        opts[i] = abi.FillOptsPOD(int(aa), int(rule), int(op), int(Precision.integer), 0.1)
    opt_index = (np.arange(n_paths) & 1).astype(np.int64)
    return FillScene(size, size, nodes, node_off, patterns, opts, opt_index)


class Scene:
    """An ordered list of painter.fill / painter.stroke calls on one surface, as C-ABI arrays (general form of FillScene)."""

    def __init__(self, width, height, nodes, node_off, patterns, kind, fill_opts, stroke_opts, opt_index, keep=()):
        self.width, self.height = width, height
        self.nodes, self.node_off, self.patterns = nodes, node_off, patterns
        self.kind = kind                # uint32[n]: 0 fill, 1 stroke
        self.fill_opts = fill_opts      # FILLOPTS_DT array
        self.stroke_opts = stroke_opts  # STROKEOPTS_DT array
        self.opt_index = opt_index      # per draw: index into fill_opts (fills) or stroke_opts (strokes)
        self.keep = keep                # arrays the PODs point into (dash arrays, gradients, stops)
        self.n = len(patterns)

    def draw_cmds(self, surface_handle, lo=0, hi=None):
        hi = self.n if hi is None else hi
        cmds = np.zeros(hi - lo, dtype=DRAWCMD_DT)
        k = self.kind[lo:hi]
        cmds["kind"] = k
        cmds["surface"] = surface_handle.value if hasattr(surface_handle, "value") else int(surface_handle)
        cmds["pattern"] = self.patterns.ctypes.data + np.arange(lo, hi, dtype=np.uint64) * PATTERN_DT.itemsize
        cmds["nodes"] = self.nodes.ctypes.data + self.node_off[lo:hi].astype(np.uint64) * NODE_DT.itemsize
        cmds["n_nodes"] = (self.node_off[lo + 1:hi + 1] - self.node_off[lo:hi]).astype(np.uint64)
        oi = self.opt_index[lo:hi].astype(np.uint64)
        f_base = self.fill_opts.ctypes.data if len(self.fill_opts) else 0
        s_base = self.stroke_opts.ctypes.data if len(self.stroke_opts) else 0
        cmds["fill"] = np.where(k == 0, f_base + oi * FILLOPTS_DT.itemsize, 0).astype(np.uint64)
        cmds["stroke"] = np.where(k == 1, s_base + oi * STROKEOPTS_DT.itemsize, 0).astype(np.uint64)
        return cmds


def _premultiplied_colours(rng, n):
    a = rng.integers(64, 256, n)
    rgb = (rng.uniform(0, 1, (n, 3)) * a[:, None]).astype(np.int64)  # pre-multiplied: c <= a
    patterns = np.zeros(n, dtype=PATTERN_DT)
    patterns["kind"] = int(PatternKind.opaque)
    patterns["pixel"]["format"] = int(Format.rgba)
    patterns["pixel"]["r"], patterns["pixel"]["g"], patterns["pixel"]["b"] = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    patterns["pixel"]["a"] = a
    return patterns


STROKE_WIDTHS = (1.5, 2.0, 3.0, 4.0, 6.0, 8.0, 12.0)


def stroke_paths_scene(n_paths=50_000, size=2048, seed=0x7A326403, aa=AntiAliasMode.default, op=Operator.src_over):
    """BASELINE config 3: `size`^2 RGBA8, n open sub-paths -- even i: polyline of 5-12 vertices with steps <= 96 px,
    odd i: two-segment cubic Bezier; width from STROKE_WIDTHS; join round (i % 4 < 2) / miter limit 10; round caps; every
    2nd path dashed [3w, 2w] with offset 0; identity CTM; translucent pre-multiplied colours, src_over."""
    rng = np.random.default_rng(seed)
    tags, pts, node_off = [], [], [0]
    for i in range(n_paths):
        x, y = rng.uniform(0, size, 2)
        if i % 2 == 0:
            nv = int(rng.integers(5, 13))
            steps = rng.uniform(-96, 96, (nv - 1, 2))
            xy = np.vstack([[x, y], [x, y] + np.cumsum(steps, axis=0)])
            xy = np.round(xy * 16) / 16
            tags.append(int(NodeTag.move_to)); pts.append((xy[0, 0], xy[0, 1], 0, 0, 0, 0))
            for k in range(1, nv):
                tags.append(int(NodeTag.line_to)); pts.append((xy[k, 0], xy[k, 1], 0, 0, 0, 0))
        else:
            c = np.round((np.array([x, y]) + np.cumsum(rng.uniform(-96, 96, (7, 2)), axis=0)) * 16) / 16
            tags.append(int(NodeTag.move_to)); pts.append((c[0, 0], c[0, 1], 0, 0, 0, 0))
            tags.append(int(NodeTag.curve_to)); pts.append((c[1, 0], c[1, 1], c[2, 0], c[2, 1], c[3, 0], c[3, 1]))
            tags.append(int(NodeTag.curve_to)); pts.append((c[4, 0], c[4, 1], c[5, 0], c[5, 1], c[6, 0], c[6, 1]))
        node_off.append(len(tags))
    nodes = np.zeros(len(tags), dtype=NODE_DT)
    nodes["tag"] = np.array(tags, dtype=np.uint32)
    nodes["p"] = np.array(pts, dtype=np.float64)
    node_off = np.array(node_off, dtype=np.int64)
    patterns = _premultiplied_colours(rng, n_paths)

    widths = np.array(STROKE_WIDTHS)[rng.integers(0, len(STROKE_WIDTHS), n_paths)]
    dashes = np.stack([3 * widths, 2 * widths], axis=1).copy()  # one [3w, 2w] row per path
    so = np.zeros(n_paths, dtype=STROKEOPTS_DT)
    so["anti_aliasing_mode"] = int(aa)
    so["line_cap_mode"] = int(abi.CapMode.round)
    so["line_join_mode"] = np.where(np.arange(n_paths) % 4 < 2, int(abi.JoinMode.round), int(abi.JoinMode.miter))
    so["op"] = int(op)
    so["precision"] = int(Precision.integer)
    so["line_width"] = widths
    so["miter_limit"] = 10.0
    so["tolerance"] = 0.1
    dashed = (np.arange(n_paths) // 2) % 2 == 1
    so["dashes"] = np.where(dashed, dashes.ctypes.data + np.arange(n_paths, dtype=np.uint64) * 16, 0).astype(np.uint64)
    so["n_dashes"] = np.where(dashed, 2, 0)
    so["ctm"][:, 0] = 1.0
    so["ctm"][:, 3] = 1.0
    kind = np.ones(n_paths, dtype=np.uint32)
    return Scene(size, size, nodes, node_off, patterns, kind, np.zeros(0, dtype=FILLOPTS_DT), so,
                 np.arange(n_paths, dtype=np.int64), keep=(dashes,))


_ORDER_CACHE = {}


def _interleave_order(counts):
    """Proportional round-robin over the parts (draw order f s f s g f s f ...): [(part, index in part), ...]."""
    if counts not in _ORDER_CACHE:
        order, taken = [], [0] * len(counts)
        for _ in range(sum(counts)):
            k = min((i for i in range(len(counts)) if taken[i] < counts[i]), key=lambda i: (taken[i] + 1) / counts[i])  # furthest behind its share
            order.append((k, taken[k]))
            taken[k] += 1
        _ORDER_CACHE[counts] = order
    return _ORDER_CACHE[counts]


def gradient_fills_scene(n=8, size=1024, seed=0x7A326405, aa=AntiAliasMode.default, op=Operator.src_over):
    """n gradient-filled shapes (2 linear, 3 radial, 3 conic per 8; rectangles and ellipses alternate),
    three random translucent stops each, identity pattern transformation, non-zero rule."""
    rng = np.random.default_rng(seed)
    kinds = [abi.GradientType.linear, abi.GradientType.linear, abi.GradientType.radial, abi.GradientType.radial,
             abi.GradientType.radial, abi.GradientType.conic, abi.GradientType.conic, abi.GradientType.conic]
    K = 0.5522847498307936  # cubic approximation of a quarter circle
    tags, pts, node_off = [], [], [0]
    grads = np.zeros(n, dtype=GRADIENT_DT)
    stops = np.zeros((n, 3), dtype=STOP_DT)
    for i in range(n):
        cx, cy = rng.uniform(0, size, 2)
        rx, ry = rng.uniform(32, 256, 2)
        cx, cy, rx, ry = (float(np.round(v * 16) / 16) for v in (cx, cy, rx, ry))
        q = lambda v: float(np.round(v * 16) / 16)  # noqa: E731
        if i % 2 == 0:  # rectangle
            corners = [(cx - rx, cy - ry), (cx + rx, cy - ry), (cx + rx, cy + ry), (cx - rx, cy + ry)]
            tags.append(int(NodeTag.move_to)); pts.append((corners[0][0], corners[0][1], 0, 0, 0, 0))
            for c in corners[1:]:
                tags.append(int(NodeTag.line_to)); pts.append((c[0], c[1], 0, 0, 0, 0))
        else:  # ellipse: four cubic arcs
            tags.append(int(NodeTag.move_to)); pts.append((cx + rx, cy, 0, 0, 0, 0))
            arcs = [((cx + rx, cy + K * ry), (cx + K * rx, cy + ry), (cx, cy + ry)), ((cx - K * rx, cy + ry), (cx - rx, cy + K * ry), (cx - rx, cy)),
                    ((cx - rx, cy - K * ry), (cx - K * rx, cy - ry), (cx, cy - ry)), ((cx + K * rx, cy - ry), (cx + rx, cy - K * ry), (cx + rx, cy))]
            for c1, c2, e in arcs:
                tags.append(int(NodeTag.curve_to)); pts.append((q(c1[0]), q(c1[1]), q(c2[0]), q(c2[1]), q(e[0]), q(e[1])))
        first = pts[node_off[-1]]
        tags.append(int(NodeTag.close_path)); pts.append((0, 0, 0, 0, 0, 0))
        tags.append(int(NodeTag.move_to)); pts.append((first[0], first[1], 0, 0, 0, 0))  # Path.close leaves a move_to behind
        node_off.append(len(tags))
        kind = kinds[i % 8]
        grads["type"][i] = int(kind)
        grads["method"][i] = int(abi.Interp.linear_rgb)
        grads["polar"][i] = int(abi.Polar.shorter)
        grads["n_stops"][i] = 3
        if kind == abi.GradientType.linear:
            grads["geom"][i] = (cx - rx, cy - ry, cx + rx, cy + ry, 0, 0)
        elif kind == abi.GradientType.radial:
            grads["geom"][i] = (cx, cy, 0.0, cx, cy, max(rx, ry))
        else:
            grads["geom"][i] = (cx, cy, float(rng.uniform(0, 2 * np.pi)), 0, 0, 0)
        grads["inv_ctm"][i] = (1, 0, 0, 1, 0, 0)
        stops["offset"][i] = (0.0, 0.5, 1.0)
        stops["color"]["space"][i] = int(abi.ColorSpace.linear_rgb)
        col = rng.uniform(0, 1, (3, 4)).astype(np.float32)
        col[:, 3] = rng.uniform(0.5, 1.0, 3)
        stops["color"]["c"][i] = col
    grads["stops"] = stops.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(3 * STOP_DT.itemsize)
    nodes = np.zeros(len(tags), dtype=NODE_DT)
    nodes["tag"] = np.array(tags, dtype=np.uint32)
    nodes["p"] = np.array(pts, dtype=np.float64)
    patterns = np.zeros(n, dtype=PATTERN_DT)
    patterns["kind"] = int(PatternKind.gradient)
    patterns["gradient"] = grads.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(GRADIENT_DT.itemsize)
    fo = np.zeros(1, dtype=FILLOPTS_DT)
    fo["anti_aliasing_mode"], fo["fill_rule"], fo["op"], fo["precision"], fo["tolerance"] = int(aa), int(FillRule.non_zero), int(op), int(Precision.integer), 0.1
    return Scene(size, size, nodes, np.array(node_off, dtype=np.int64), patterns, np.zeros(n, dtype=np.uint32), fo,
                 np.zeros(0, dtype=STROKEOPTS_DT), np.zeros(n, dtype=np.int64), keep=(grads, stops))


def mixed_scene(scene_index, size=1024, n_fills=32, n_strokes=24, n_gradients=8, seed_base=0x7A326405):
    """BASELINE config 5 shape: one SVG-like scene of 64 ordered draw calls -- 32 fills (config-2 generator,
    r 8-256), 24 strokes (config-3 generator) and 8 gradient fills (2 linear, 3 radial, 3 conic; rectangles and ellipses),
    interleaved in submission order; seed = seed_base + scene."""
    seed = (seed_base + scene_index) & 0xFFFFFFFF
    parts = []  # (what, scene)
    if n_fills:
        parts.append(("fill", cubic_paths_scene(n_fills, size, seed=seed, r_log2=(3.0, 8.0))))
    if n_strokes:
        parts.append(("stroke", stroke_paths_scene(n_strokes, size, seed=seed ^ 0x5A5A5A5A)))
    if n_gradients:
        parts.append(("gradient", gradient_fills_scene(n_gradients, size, seed=seed ^ 0x3C3C3C3C)))
    counts = tuple(p.n for _, p in parts)
    order = _interleave_order(counts)
    total = sum(counts)
    node_parts, node_off = [], [0]
    patterns = np.zeros(total, dtype=PATTERN_DT)
    kind = np.zeros(total, dtype=np.uint32)
    opt_index = np.zeros(total, dtype=np.int64)
    fill_opts_list, fill_base, stroke_opts, keep = [], {}, np.zeros(0, dtype=STROKEOPTS_DT), ()
    for what, p in parts:
        if what == "stroke":
            stroke_opts = p.stroke_opts
        else:
            fo = p.fill_opts if isinstance(p.fill_opts, np.ndarray) else np.frombuffer(p.fill_opts, dtype=FILLOPTS_DT).copy()
            fill_base[what] = sum(len(x) for x in fill_opts_list)
            fill_opts_list.append(fo)
        keep = keep + tuple(getattr(p, "keep", ()))
    for j, (k, i) in enumerate(order):
        what, p = parts[k]
        node_parts.append(p.nodes[p.node_off[i]:p.node_off[i + 1]])
        node_off.append(node_off[-1] + len(node_parts[-1]))
        patterns[j] = p.patterns[i]
        kind[j] = 1 if what == "stroke" else 0
        opt_index[j] = i if what == "stroke" else fill_base[what] + p.opt_index[i]
    fill_opts = np.concatenate(fill_opts_list) if fill_opts_list else np.zeros(0, dtype=FILLOPTS_DT)
    return Scene(size, size, np.concatenate(node_parts), np.array(node_off, dtype=np.int64), patterns, kind, fill_opts, stroke_opts, opt_index,
                 keep=keep)
