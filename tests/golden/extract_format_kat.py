"""Extracts the reference's per-format `src_over` / `dst_in` known-answer tests (src/compositor.zig:2452-3076) into
tests/golden/format_kat.json: {operator, dst pixel, src pixel, expected pixel}, every pixel as {format, r, g, b, a}.

The Zig tests build their arguments with pixel conversions (`pixel.RGB.fromPixel(fg.asPixel())`, `fg.multiply()`, ...); the
small evaluator below restates exactly those conversions (pixel.zig:399-433 fromPixel, 476-484 multiply, 569-626 Alpha
fromPixel / shlr) so that the fixture holds plain pixels.  Run in the build container; the JSON is the committed fixture.
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/compositor.zig"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "format_kat.json")
BITS = {"alpha8": 8, "alpha4": 4, "alpha2": 2, "alpha1": 1}
TYPES = {"RGB": "rgb", "RGBA": "rgba", "ARGB": "argb", "XRGB": "xrgb", "Alpha8": "alpha8", "Alpha4": "alpha4", "Alpha2": "alpha2", "Alpha1": "alpha1"}


def shlr(val, fb, tb):  # pixel.zig:587-626
    if fb == 1:
        return val * ((1 << tb) - 1)
    if val == 0 or fb == tb:
        return val
    if fb > tb:
        return val >> (fb - tb)
    d = tb - fb
    if d in (2, 4):
        return (val << d) + val
    return (val << d) | (val << (tb - 2 * fb)) | (val << (tb - 3 * fb)) | val


def alpha_of(p, bits):  # Alpha(T).fromPixel
    if p["format"] in ("rgb", "xrgb"):
        return (1 << bits) - 1
    fb = BITS.get(p["format"], 8)
    return shlr(p["a"], fb, bits)


def from_pixel(fmt, p):
    if p["format"] == fmt:
        return dict(p)
    if fmt in BITS:
        return {"format": fmt, "r": 0, "g": 0, "b": 0, "a": alpha_of(p, BITS[fmt])}
    rgba = fmt in ("rgba", "argb")
    if p["format"] in ("rgb", "xrgb", "rgba", "argb"):
        a = p["a"] if p["format"] in ("rgba", "argb") else 255
        return {"format": fmt, "r": p["r"], "g": p["g"], "b": p["b"], "a": a if rgba else 0}
    return {"format": fmt, "r": 0, "g": 0, "b": 0, "a": alpha_of(p, 8) if rgba else 0}


def multiply(p):
    return {"format": p["format"], "r": p["r"] * p["a"] // 255, "g": p["g"] * p["a"] // 255, "b": p["b"] * p["a"] // 255, "a": p["a"]}


def literal(fmt, body):
    v = {k: int(x) for k, x in re.findall(r"\.(\w)\s*=\s*(\d+)", body)}
    return {"format": fmt, "r": v.get("r", 0), "g": v.get("g", 0), "b": v.get("b", 0), "a": v.get("a", 0)}


def evaluate(expr, env):
    expr = expr.strip().rstrip(",").strip()
    if expr.endswith(".asPixel()"):
        return evaluate(expr[:-len(".asPixel()")], env)
    if expr.endswith(".multiply()"):
        return multiply(evaluate(expr[:-len(".multiply()")], env))
    m = re.fullmatch(r"pixel\.(\w+)\.fromPixel\((.*)\)", expr, re.S)
    if m:
        return from_pixel(TYPES[m.group(1)], evaluate(m.group(2), env))
    m = re.fullmatch(r"(?:pixel\.Pixel)?\s*\.?\{\s*\.(\w+)\s*=\s*(.*)\}", expr, re.S)
    if m:
        fmt, inner = m.group(1), m.group(2).strip().rstrip(",").strip()
        if inner.startswith(".{"):
            return literal(fmt, inner)
        return dict(evaluate(inner, env), format=fmt)
    if re.fullmatch(r"\w+", expr):
        return dict(env[expr])
    raise ValueError(f"cannot evaluate {expr!r}")


def split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({":
            depth += 1
        elif ch in ")}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


text = open(REF).read()
cases = []
for name in ("src_over", "dst_in"):
    start = text.index(f'test "{name}" {{')
    end = text.index('\ntest "', start + 10)
    blk = text[start:end]
    env = {}
    for m in re.finditer(r"const (\w+): pixel\.RGBA = \.\{([^}]*)\};", blk):
        env[m.group(1)] = literal("rgba", m.group(2))
    pos = 0
    # walk the block in order so that block-local consts are defined before the expectations that use them
    token = re.compile(r"(?:const|var) (\w+) = ([^;]+);|(\w+)\.a = (\d+);|try testing\.expectEqualDeep\(", re.S)
    while True:
        m = token.search(blk, pos)
        if not m:
            break
        if m.group(1):
            env[m.group(1)] = evaluate(m.group(2), env)
            pos = m.end()
            continue
        if m.group(3):  # `bg_alpha1.a = 0;` style mutation of a var
            env[m.group(3)]["a"] = int(m.group(4))
            pos = m.end()
            continue
        # find the matching close paren of expectEqualDeep(
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(blk[i], 0)
            i += 1
        args = split_args(blk[m.end():i - 1])
        expected = evaluate(args[0], env)
        call = args[1].strip()
        inner = call[call.index("(") + 1:call.rindex(")")]
        prec, dst, src, op = [a.strip() for a in split_args(inner)]
        cases.append({"operator": op.lstrip("."), "precision": prec.lstrip("."), "dst": evaluate(dst, env), "src": evaluate(src, env),
                      "expected": expected})
        pos = i
json.dump({"source": "src/compositor.zig:2452-3076", "cases": cases}, open(OUT, "w"), indent=0)
print(len(cases), "cases ->", OUT)
