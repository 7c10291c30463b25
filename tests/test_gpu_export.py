"""Device-side export transform (export_png.zig:150-373) through z2d_surface_export.

The checker is the numpy restatement of the reference's export in tests/golden_util.py (the same code every golden comparison
goes through, so it is pinned by the 163 reference PNGs): bit-exact for every format, with and without the sRGB profile, at widths
that leave a ragged last byte in the packed formats.  Then whole files: a PNG written from the device bytes must decode to exactly
the pixels of the reference's own golden image.
"""
import numpy as np
import pytest
from PIL import Image

from tests import golden_util, specs
from z2d_b200 import abi, host
from z2d_b200.abi import AntiAliasMode, Format

pytestmark = pytest.mark.gpu

FORMATS = [Format.argb, Format.xrgb, Format.rgb, Format.rgba, Format.alpha8, Format.alpha4, Format.alpha2, Format.alpha1]


def expected_rows(sfc, profile, filter_byte):
    view, kind = golden_util.export_view(sfc, profile)  # decoded export samples, (h, w[, c])
    h, w = view.shape[:2]
    if kind in ("RGB", "RGBA", "L8"):
        rows = view.reshape(h, -1)
    else:  # PNG packs sub-byte greys most significant first, every row padded to a byte
        bits = {"L4": 4, "L2": 2, "L1": 1}[kind]
        planes = ((view[..., None].astype(np.uint8) >> np.arange(bits - 1, -1, -1, dtype=np.uint8)) & 1).reshape(h, w * bits)
        rows = np.packbits(planes, axis=1, bitorder="big")
    if filter_byte:
        rows = np.concatenate([np.zeros((h, 1), np.uint8), rows], axis=1)
    return np.ascontiguousarray(rows, dtype=np.uint8)


def random_surface(cuda, fmt, w, h, seed):
    rng = np.random.default_rng(seed)
    sfc = host.Surface(fmt, w, h, None, cuda)
    if fmt in (Format.argb, Format.rgba):  # premultiplied: colour <= alpha, with plenty of a == 0 and a == 255
        a = rng.choice([0, 1, 2, 127, 128, 254, 255] + list(range(256)), size=(h, w)).astype(np.int64)
        c = (rng.integers(0, 256, size=(h, w, 3)) * a[..., None] + 127) // 255
        px = np.concatenate([c, a[..., None]], axis=-1).astype(np.uint8)
        raw = px[..., [2, 1, 0, 3]] if fmt == Format.argb else px
        sfc.upload(np.ascontiguousarray(raw).reshape(-1))
    else:
        sfc.upload(rng.integers(0, 256, size=sfc.byte_len(), dtype=np.uint8))
    return sfc


@pytest.mark.parametrize("fmt", FORMATS, ids=[f.name for f in FORMATS])
@pytest.mark.parametrize("profile", [None, "srgb"])
@pytest.mark.parametrize("size", [(1, 1), (37, 11), (256, 64), (1031, 9)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_export_rows_match_reference_transform(cuda, fmt, profile, size):
    w, h = size
    sfc = random_surface(cuda, fmt, w, h, seed=1000 * int(fmt) + w)
    for filter_byte in (False, True):
        got = sfc.export(profile, filter_byte)
        exp = expected_rows(sfc, profile, filter_byte)
        assert got.shape == exp.shape
        assert np.array_equal(got, exp), f"{int((got != exp).sum())} of {got.size} exported bytes differ"


def test_export_size_and_argument_checks(cuda):
    import ctypes as C
    sfc = host.Surface(Format.alpha2, 13, 5, None, cuda)
    lib = cuda.lib
    assert lib.z2d_surface_export_size(sfc.handle, 0) == 4 * 5
    assert lib.z2d_surface_export_size(sfc.handle, 2) == 5 * 5
    buf = np.zeros(64, np.uint8)
    assert lib.z2d_surface_export(sfc.handle, 0, buf.ctypes.data_as(C.c_void_p), 19) < 0   # wrong size
    assert lib.z2d_surface_export(sfc.handle, 8, buf.ctypes.data_as(C.c_void_p), 20) < 0   # unknown flag
    assert lib.z2d_surface_export(sfc.handle, 0, None, 20) < 0


def test_export_of_a_band_is_the_rows_of_the_canvas(cuda):
    px = host.Pixel.rgba(10, 20, 30, 40)
    whole = host.Surface(Format.rgba, 64, 64, None, cuda)
    band = host.Surface(Format.rgba, 64, 64, None, cuda, band=(16, 32))
    for s in (whole, band):
        ctx = host.Context(s)
        ctx.set_source_to_pixel(px)
        ctx.move_to(5, 3)
        ctx.line_to(60, 20)
        ctx.line_to(30, 61)
        ctx.close_path()
        ctx.fill()
    assert np.array_equal(band.export(), whole.export()[16:48])


PNG_PATH_SCENES = ["003_fill_triangle", "007_fill_bezier", "049_fill_triangle_alpha4_gray", "050_fill_triangle_alpha2_gray",
                   "051_fill_triangle_alpha1_gray", "047_fill_triangle_alpha_gray", "046_fill_triangle_alpha"]


def _same_png(tmp_path, sfc, golden, profile=None):
    out = tmp_path / "out.png"
    sfc.write_png(str(out), profile)
    got, exp = Image.open(out), Image.open(golden)
    assert got.mode == exp.mode and got.size == exp.size
    assert np.array_equal(np.asarray(got), np.asarray(exp))


@pytest.mark.parametrize("stem", PNG_PATH_SCENES)
def test_written_png_decodes_to_the_reference_golden(cuda, tmp_path, stem):
    aa = AntiAliasMode.multisample_4x
    sfc = specs.PATH_SCENES[stem](specs.bind(cuda), aa)
    _same_png(tmp_path, sfc, golden_util.golden_path(stem, aa))


def test_written_png_with_srgb_profile(cuda, tmp_path):
    stem = "072_gamma_srgb"
    sfc = specs.COMPOSITOR_SCENES[stem](specs.bind(cuda))
    out = tmp_path / "out.png"
    sfc.write_png(str(out), "srgb")
    got, exp = Image.open(out), Image.open(golden_util.golden_path(stem))
    assert got.info.get("gamma") == pytest.approx(exp.info.get("gamma"))
    d = np.abs(np.asarray(got).astype(np.int32) - np.asarray(exp).astype(np.int32))
    assert d.max() <= 2  # float gradient scene: +-1 LSB before the gamma curve


@pytest.mark.parametrize("fmt", FORMATS, ids=[f.name for f in FORMATS])
def test_get_pixel_reads_what_the_surface_stores(cuda, fmt):
    w, h = 29, 7
    sfc = random_surface(cuda, fmt, w, h, seed=77 + int(fmt))
    px = sfc.pixels()
    raw_alpha = host.decode_pixels(sfc.download(), fmt, w, h)[..., 3]
    for x, y in [(0, 0), (28, 6), (13, 3), (1, 5), (27, 0)]:
        got = sfc.get_pixel(x, y)
        assert got.format == fmt
        if fmt in (Format.argb, Format.rgba):
            assert (got.r, got.g, got.b, got.a) == tuple(int(v) for v in px[y, x])
        elif fmt in (Format.xrgb, Format.rgb):
            assert (got.r, got.g, got.b) == tuple(int(v) for v in px[y, x, :3])
        else:
            assert got.a == int(raw_alpha[y, x])
    for x, y in [(-1, 0), (0, -1), (29, 0), (0, 7)]:
        assert sfc.get_pixel(x, y) is None
    sfc.put_pixel(4, 2, sfc.get_pixel(13, 3))
    again = sfc.get_pixel(4, 2)
    ref = sfc.get_pixel(13, 3)
    assert (again.r, again.g, again.b, again.a) == (ref.r, ref.g, ref.b, ref.a)


@pytest.mark.parametrize("fmt", [Format.rgba, Format.argb, Format.rgb, Format.xrgb, Format.alpha8, Format.alpha4, Format.alpha2, Format.alpha1])
def test_downsample_matches_reference_box_average(cuda, oracle, fmt):
    """Surface.downsample (surface.zig:447-490, 687-709) as a standalone device call: every format, sizes that are not multiples
    of 4 (the remainder rows / columns are dropped) and surfaces too small to downsample."""
    from tests import specs
    rng = np.random.default_rng(5)
    for w, h in ((64, 32), (37, 23), (601, 172), (3, 9), (4, 4)):
        data = rng.integers(0, 256, abi.surface_byte_len(fmt, w, h), dtype=np.uint8)
        res = []
        for z in (specs.bind(cuda), specs.bind(oracle)):
            s = z.Surface(fmt, w, h)
            s.upload(data.copy())
            s.downsample()
            res.append((s.get_width(), s.get_height(), s.download().copy()))
        assert res[0][:2] == res[1][:2] == ((w // 4, h // 4) if w >= 4 and h >= 4 else (w, h))
        a, b = res[0][2], res[1][2]
        nbits = res[0][0] * res[0][1] * abi.FORMAT_BITS[fmt]
        if nbits % 8:  # the last byte of a packed surface is only partly defined
            a, b = a.copy(), b.copy()
            m = (1 << (nbits % 8)) - 1
            a[-1] &= m
            b[-1] &= m
        if fmt in (Format.rgb, Format.xrgb):
            a, b = a.reshape(-1, 4).copy(), b.reshape(-1, 4).copy()
            pad = 3 if fmt == Format.rgb else 3
            a[:, pad] = 0
            b[:, pad] = 0
        assert np.array_equal(a, b), f"{fmt.name} {w}x{h}"
