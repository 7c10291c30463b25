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
