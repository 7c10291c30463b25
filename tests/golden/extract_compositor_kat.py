"""Extracts the reference's operator known-answer tables into tests/golden/compositor_kat.json.

Source: /root/reference/src/compositor.zig, tests "composite, all operators (integer)" (3078-3443) and
"composite, all operators (float)" (3445-3860): {name, operator, expected RGBA8, bg colour, fg colour}.
Run in the build container (the reference is not present on the GPU box); the JSON is the committed fixture.
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/compositor.zig"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compositor_kat.json")

text = open(REF).read()
blocks = {}
for prec in ("integer", "float"):
    start = text.index(f'test "composite, all operators ({prec})"')
    end = text.index("const TestFn", start)
    blocks[prec] = text[start:end]

case_re = re.compile(
    r'\.name\s*=\s*"([^"]+)",\s*\.operator\s*=\s*\.(\w+),\s*'
    r'\.expected\s*=\s*\.\{\s*\.rgba\s*=\s*\.\{\s*\.r\s*=\s*(\d+),\s*\.g\s*=\s*(\d+),\s*\.b\s*=\s*(\d+),\s*\.a\s*=\s*(\d+)\s*\}\s*\},\s*'
    r'\.bg\s*=\s*\.\{\s*\.(rgba?)\s*=\s*\.\{([^}]*)\}\s*\},\s*'
    r'\.fg\s*=\s*\.\{\s*\.(rgba?)\s*=\s*\.\{([^}]*)\}\s*\},', re.S)

out = []
for prec, blk in blocks.items():
    for m in case_re.finditer(blk):
        name, op, r, g, b, a, bk, bv, fk, fv = m.groups()
        out.append({"precision": prec, "name": name, "operator": op, "expected": [int(r), int(g), int(b), int(a)],
                    "bg": {bk: [float(v) for v in bv.split(",") if v.strip()]},
                    "fg": {fk: [float(v) for v in fv.split(",") if v.strip()]}})
json.dump({"source": "src/compositor.zig:3078-3860", "cases": out}, open(OUT, "w"), indent=1)
print(len(out), "cases ->", OUT)
