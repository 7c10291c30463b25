"""Extracts the reference's Polygon.inBox known-answer table (src/internal/tess/Polygon.zig:452-784, 24 cases) into
tests/golden/inbox_kat.json.  Note the reference test calls inBox(scale, box_height, box_width), i.e. with the two box
arguments in the opposite order of the signature (scale, box_width, box_height): the fixture stores them as passed."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/internal/tess/Polygon.zig"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "inbox_kat.json")
text = open(REF).read()
blk = text[text.index('test "Polygon.inBox"'):]
blk = blk[:blk.index("const TestFn")]
case_re = re.compile(r'\.name = "([^"]+)",\s*\.polygon = \.\{\s*\.extent_left = ([-\d.]+),\s*\.extent_top = ([-\d.]+),\s*\.extent_right = ([-\d.]+),\s*'
                     r'\.extent_bottom = ([-\d.]+),\s*\},\s*\.scale = ([\d.]+),\s*\.box_height = (\d+),\s*\.box_width = (\d+),\s*\.expected = (true|false)', re.S)
cases = [{"name": n, "left": float(l), "top": float(t), "right": float(r), "bottom": float(b), "scale": float(s),
          "arg_width": int(bh), "arg_height": int(bw), "expected": e == "true"} for n, l, t, r, b, s, bh, bw, e in case_re.findall(blk)]
json.dump({"source": "src/internal/tess/Polygon.zig:452-784", "cases": cases}, open(OUT, "w"), indent=0)
print(len(cases), "cases ->", OUT)
