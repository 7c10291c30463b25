#!/usr/bin/env python3
"""Import the reference's golden images as test fixtures.

Copies the PNG goldens produced by z2d's own acceptance suite
(`zig build spec`, spec/main_spec.zig:772-831; files under spec/files/) into
tests/golden/spec_files/ unchanged, and writes MANIFEST.json (name, size,
mode, sha256).  They are the vectors that pin the CPU oracle: tests decode
them with Pillow and compare per pixel (the reference compares PNG file
hashes, which would need Zig's deflate).  Run only where /root/reference
exists; /root/reference is never read at test time.
"""
import hashlib, json, pathlib, shutil, sys
from PIL import Image

src = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/spec/files")
dst = pathlib.Path(__file__).resolve().parent / "spec_files"
dst.mkdir(exist_ok=True)
manifest = {}
for p in sorted(src.glob("*.png")):
    shutil.copyfile(p, dst / p.name)
    im = Image.open(p)
    manifest[p.name] = {"size": list(im.size), "mode": im.mode, "sha256": hashlib.sha256(p.read_bytes()).hexdigest()}
(dst / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, sort_keys=True) + "\n")
print(len(manifest), "goldens imported")

# The text scenes (spec/074, 080, 085) read glyph outlines from the reference's test fonts (spec/test-fonts; licences in the
# reference's LICENSE file: Inter and Montserrat are OFL-1.1, DejaVu Sans is under the Bitstream Vera / DejaVu licence).
fsrc = src.parent / "test-fonts"
fdst = pathlib.Path(__file__).resolve().parent / "fonts"
fdst.mkdir(exist_ok=True)
fonts = {}
for p in sorted(fsrc.glob("*.ttf")):
    shutil.copyfile(p, fdst / p.name)
    fonts[p.name] = {"bytes": p.stat().st_size, "sha256": hashlib.sha256(p.read_bytes()).hexdigest()}
(fdst / "MANIFEST.json").write_text(json.dumps(fonts, indent=1, sort_keys=True) + "\n")
print(len(fonts), "fonts imported")
