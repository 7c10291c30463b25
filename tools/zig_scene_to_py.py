#!/usr/bin/env python3
"""Mechanical Zig -> Python transliteration of the reference's acceptance scenes.

The scenes in spec/NNN_*.zig use a small, regular vocabulary (Context setters,
path verbs, integer-literal arithmetic, simple loops), so they can be carried
over line by line.  The output is Python source for tests/specs/ that drives
z2d_b200.host through the same calls; it is reviewed and, where the
transliteration falls short, fixed by hand (see the header of the generated
file).  Zig comptime-integer arithmetic is kept exact by wrapping integer
literals in `ZI`, an int subclass whose `/` truncates like Zig's.

usage: zig_scene_to_py.py OUT.py SCENE.zig [SCENE.zig ...]
"""
import re
import sys

SETTER_ENUMS = {
    "set_line_join_mode": "JoinMode", "set_line_cap_mode": "CapMode", "set_fill_rule": "FillRule",
    "set_operator": "Operator", "set_precision": "Precision", "set_anti_aliasing_mode": "AntiAliasMode",
    "set_dither": "DitherType",
}
SURFACE_FMT = {"image_surface_rgb": "Format.rgb", "image_surface_rgba": "Format.rgba", "image_surface_xrgb": "Format.xrgb",
               "image_surface_argb": "Format.argb", "image_surface_alpha8": "Format.alpha8", "image_surface_alpha4": "Format.alpha4",
               "image_surface_alpha2": "Format.alpha2", "image_surface_alpha1": "Format.alpha1"}


def snake(name):
    return re.sub(r"(?<!^)(?=[A-Z])", "_", name).lower()


def conv_expr(e):
    e = e.strip()
    # pixel literals
    e = re.sub(r"\.\{\s*\.rgb\s*=\s*\.\{\s*\.r\s*=\s*([^,]+),\s*\.g\s*=\s*([^,]+),\s*\.b\s*=\s*([^,}]+?)\s*\}\s*\}", r"z.Pixel.rgb(\1, \2, \3)", e)
    e = re.sub(r"\.\{\s*\.rgba\s*=\s*\.\{\s*\.r\s*=\s*([^,]+),\s*\.g\s*=\s*([^,]+),\s*\.b\s*=\s*([^,]+),\s*\.a\s*=\s*([^,}]+?)\s*\}\s*\}", r"z.Pixel.rgba(\1, \2, \3, \4)", e)
    e = re.sub(r"\.\{\s*\.alpha(\d)\s*=\s*\.\{\s*\.a\s*=\s*([^,}]+?)\s*\}\s*\}", r"z.Pixel.alpha\1(\2)", e)
    e = re.sub(r"&\.\{\s*\}", "()", e)
    e = re.sub(r"&\.\{\s*([^{}]*?)\s*\}", r"(\1,)", e)       # &.{ 25, 5 } -> tuple
    e = re.sub(r"@as\(\s*\w+\s*,\s*", "(", e)
    e = e.replace("@floatFromInt(", "float(").replace("@intCast(", "(").replace("@intFromFloat(", "int(")
    e = e.replace("@cos(", "math.cos(").replace("@sin(", "math.sin(").replace("@floor(", "math.floor(").replace("@sqrt(", "math.sqrt(")
    e = re.sub(r"@mod\(([^,]+),\s*([^)]+)\)", r"((\1) % (\2))", e)
    e = e.replace("&context", "context").replace("&sfc", "sfc")
    e = re.sub(r"\bcontext\.(\w+)\(", lambda m: "context." + snake(m.group(1)) + "(", e)
    # integer literals -> ZI (skip floats, hex handled too)
    e = re.sub(r"(?<![\w.])(0x[0-9A-Fa-f]+|\d+)(?![\w.]|\.\d)", r"ZI(\1)", e)
    e = re.sub(r"\btrue\b", "True", e)
    e = re.sub(r"\bfalse\b", "False", e)
    return e


def conv_call_enums(line):
    m = re.search(r"context\.(\w+)\(\s*\.(\w+)\s*\)", line)
    if m and m.group(1) in SETTER_ENUMS:
        return line[:m.start()] + f"context.{m.group(1)}({SETTER_ENUMS[m.group(1)]}.{m.group(2)})" + line[m.end():]
    return line


LW_IDIOM = re.compile(r"context\.setLineWidth\(lw: \{\s*var ux = (\w+);\s*var uy = \1;\s*try context\.deviceToUserDistance\(&ux, &uy\);\s*"
                      r"if \(ux < uy\) \{\s*break :lw uy;\s*\}\s*break :lw ux;\s*\}\);", re.S)


def translate(path):
    src = open(path).read()
    src = re.sub(r"//[^\n]*", "", src)
    src = LW_IDIOM.sub(lambda m: f"context.setLineWidth(_lw_max(context, {m.group(1)}));", src)
    # direct pattern assignment == setSourceToPixel
    src = re.sub(r"context\.pattern = \.\{\s*\.opaque_pattern = \.\{\s*\.pixel = (.*?\}\s*\}),\s*\},\s*\};", r"context.setSourceToPixel(\1);", src, flags=re.S)
    # Pixel.fromColor(.{ .rgb = .{ a, b, c } })
    src = re.sub(r"z2d\.(?:pixel\.)?Pixel\.fromColor\(\.\{\s*\.(\w+) = \.\{([^}]*)\}\s*\}\)", r'z.Pixel.from_color({"\1": (\2)})', src)
    src = re.sub(r"\(z2d\.pixel\.RGBA\{ \.r = (\d+), \.g = (\d+), \.b = (\d+), \.a = (\d+) \}\)\.multiply\(\)\.asPixel\(\)",
                 lambda m: "z.Pixel.rgba(%d, %d, %d, %d)" % tuple([int(m.group(k)) * int(m.group(4)) // 255 for k in (1, 2, 3)] + [int(m.group(4))]), src)
    src = src.replace("saved_source.opaque_pattern.pixel", "saved_source.value")
    # gradients
    src = re.sub(r"z2d\.Gradient\.init\(\.\{\s*\.type = \.\{\s*\.linear = \.\{\s*\.x0 = ([^,]+),\s*\.y0 = ([^,]+),\s*\.x1 = ([^,]+),\s*\.y1 = ([^,]+),\s*\}\s*\}\s*,?\s*\}\)",
                 r"z.Gradient.linear(\1, \2, \3, \4)", src)
    src = re.sub(r"(\w+)\.addStop\(alloc, ([^,]+), \.\{\s*\.(\w+) = \.\{([^}]*)\}\s*\}\)", r'\1.add_stop(\2, {"\3": (\4)})', src)
    src = re.sub(r"(\w+)\.asPattern\(\)", r"z.Pattern.gradient(\1)", src)
    src = src.replace("context.setLineCapMode(if (round) .round else .square);", "context.setLineCapMode(CAPSEL);")
    src = re.sub(r"const y_offset = yoff: \{\s*var y: f64 = 50;\s*if \(reverse\) y \+= 100;\s*if \(round\) y \+= 200;\s*break :yoff y;\s*\};",
                 "var y_offset: f64 = 50;\nif (reverse) y_offset += 100;\nif (round) y_offset += 200;", src)
    stem = re.search(r'pub const filename = "([^"]+)"', src).group(1)
    out = []
    depth = 0
    fn_defers = []
    block_defers = {}
    in_fn = False
    lines = src.split("\n")
    i = 0

    def emit(s, d=None):
        out.append("    " * (depth if d is None else d) + s)

    while i < len(lines):
        ln = lines[i].strip()
        i += 1
        if not ln or ln.startswith("const ") and "@import" in ln or ln.startswith("pub const filename"):
            continue
        # join multi-line statements until ; or { or }
        while (not (ln.endswith(";") or ln.endswith("{") or ln.endswith("}") or ln.endswith("},")) or ln.count("(") > ln.count(")")) and i < len(lines):
            ln += " " + lines[i].strip()
            i += 1
        ln = re.sub(r"^_ = ", "", ln)
        m = re.match(r"(pub )?fn (\w+)\((.*)\) (!?[\w.]+) \{$", ln)
        if m:
            ptypes = {p.strip().split(":")[0].strip(): p.strip().split(":")[1].strip() for p in m.group(3).split(",") if ":" in p}
            params = [p for p in ptypes if p not in ("io", "alloc")]
            if m.group(2) == "render":
                emit(f'@path_scene("{stem}")' if "aa_mode" in params else f'@compositor_scene("{stem}")', 0)
                emit(f"def s{stem[:3]}(z{', aa_mode' if 'aa_mode' in params else ''}):", 0)
            else:
                emit(f"def {m.group(2)}_{stem[:3]}(z, {', '.join(params)}):", 0)
            depth = 1
            emit("try:")
            depth = 2
            in_fn = True
            fn_defers = []
            for pn in params:
                if ptypes[pn] == "f64":
                    emit(f"{pn} = float({pn})")
            continue
        if ln == "}" and depth == 2 and in_fn:
            depth = 1
            emit("finally:")
            if fn_defers:
                for d in reversed(fn_defers):
                    emit(d, 2)
            else:
                emit("pass", 2)
            depth = 0
            in_fn = False
            emit("", 0)
            continue
        if ln.startswith("}"):
            if depth in block_defers:  # defers of a bare { } scope run when the scope closes
                for d in reversed(block_defers.pop(depth)):
                    emit(d)
            depth -= 1
            rest = ln[1:].strip()
            if rest.startswith("else if"):
                cond = re.match(r"else if \((.*)\) \{$", rest).group(1)
                emit(f"elif {conv_expr(cond)}:")
                depth += 1
            elif rest.startswith("else"):
                emit("else:")
                depth += 1
            continue
        if ln.startswith("defer "):
            body = ln[6:].rstrip(";")
            if "deinit" in body:
                continue
            if depth in block_defers:
                block_defers[depth].append(conv_stmt(body, stem))
            else:
                fn_defers.append(conv_stmt(body, stem))
            continue
        m = re.match(r"for \((\w+|\d+)\.\.(.+?)\) \|(\w+)\| \{$", ln)
        if m:
            emit(f"for {m.group(3)} in map(ZI, range({conv_expr(m.group(1))}, {conv_expr(m.group(2))})):")
            depth += 1
            continue
        m = re.match(r"for \((\w+|\d+)\.\.(.+?)\) \|(\w+)\| (.*);$", ln)
        if m:
            emit(f"for {m.group(3)} in map(ZI, range({conv_expr(m.group(1))}, {conv_expr(m.group(2))})):")
            emit(conv_stmt(m.group(4), stem), depth + 1)
            continue
        if ln == "{":
            emit("if True:")
            depth += 1
            block_defers[depth] = []
            continue
        m = re.match(r"if \((.*)\) \{$", ln)
        if m:
            emit(f"if {conv_expr(m.group(1))}:")
            depth += 1
            continue
        m = re.match(r"if \((.*?)\) ((?:try )?(?:context\.|\w+ [+-]?= ).*);$", ln)
        if m:
            emit(f"if {conv_expr(m.group(1))}:")
            emit(conv_stmt(m.group(2), stem), depth + 1)
            continue
        emit(conv_stmt(ln.rstrip(";"), stem))
    return "\n".join(out)


def conv_stmt(s, stem):
    s = s.strip()
    s = re.sub(r"^try ", "", s)
    m = re.match(r"(?:comptime )?(?:var|const) (\w+)(?::\s*[\w.\[\]]+)? = (.*)$", s)
    if m:
        name, rhs = m.group(1), m.group(2)
        rhs = re.sub(r"^try ", "", rhs)
        ms = re.match(r"z2d\.Surface\.init\(\s*\.(\w+),\s*alloc,\s*(.+),\s*(.+?)\s*,?\s*\)$", rhs)
        if ms:
            return f"{name} = z.Surface({SURFACE_FMT[ms.group(1)]}, {conv_expr(ms.group(2))}, {conv_expr(ms.group(3))})"
        ms = re.match(r"z2d\.Surface\.initPixel\(\s*(.+?),\s*alloc,\s*(.+?),\s*(.+?)\s*,?\s*\)$", rhs)
        if ms:
            return f"{name} = z.SurfacePixel({conv_expr(ms.group(1))}, {conv_expr(ms.group(2))}, {conv_expr(ms.group(3))})"
        if rhs.startswith("z2d.Context.init("):
            return f"{name} = z.Context(sfc)"
        return f"{name} = {conv_expr(rhs)}"
    s2 = conv_expr(s)
    s2 = conv_call_enums(s2)
    s2 = s2.replace("CAPSEL", "(CapMode.round if round else CapMode.square)")
    # calls to local helper functions get the scene suffix + namespace
    m = re.match(r"(\w+)\((.*)\)$", s2)
    if m and m.group(1) not in ("ZI", "float", "int") and not s2.startswith("context."):
        return f"{m.group(1)}_{stem[:3]}(z, {m.group(2)})"
    return s2


HEADER = '''"""Stroke / dash / transform scenes, transliterated from the reference's spec/NNN_*.zig by
tools/zig_scene_to_py.py and fixed up by hand where noted.  `ZI` keeps Zig's comptime
integer arithmetic (`/` truncates)."""
import math

from . import compositor_scene, path_scene
from z2d_b200.abi import (AntiAliasMode, CapMode, DitherType, FillRule, Format, Interp, JoinMode, Operator, Polar, Precision)


class ZI(int):
    """Zig comptime_int: + - * stay ZI, / between two ZI truncates toward zero."""

    def _w(self, v):
        return ZI(v) if isinstance(v, int) else v

    def __add__(self, o):
        return self._w(int(self) + o) if isinstance(o, int) else float(self) + o

    __radd__ = __add__

    def __sub__(self, o):
        return self._w(int(self) - o) if isinstance(o, int) else float(self) - o

    def __rsub__(self, o):
        return self._w(o - int(self)) if isinstance(o, int) else o - float(self)

    def __mul__(self, o):
        return self._w(int(self) * o) if isinstance(o, int) else float(self) * o

    __rmul__ = __mul__

    def __neg__(self):
        return ZI(-int(self))

    def __truediv__(self, o):
        if isinstance(o, int):
            q = abs(int(self)) // abs(int(o))
            return ZI(q if (int(self) >= 0) == (int(o) >= 0) else -q)
        return float(self) / o

    def __rtruediv__(self, o):
        if isinstance(o, int):
            return ZI(o) / self
        return o / float(self)


def _lw_max(context, lw):
    """`lw: { var ux = lw; var uy = lw; deviceToUserDistance(&ux, &uy); break :lw max }` idiom."""
    ux, uy = context.transformation.device_to_user_distance(float(lw), float(lw))
    return uy if ux < uy else ux

'''

if __name__ == "__main__":
    out_path, files = sys.argv[1], sys.argv[2:]
    parts = [HEADER]
    for f in files:
        parts.append(translate(f))
        parts.append("")
    open(out_path, "w").write("\n".join(parts))
    print("wrote", out_path)
