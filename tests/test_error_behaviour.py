"""Error behaviour at the drop-in boundary, mirroring the reference's own tests:
painter.zig:346-398 (InvalidMatrix before the empty-node early-out, PixelSourceNotPreMultiplied),
painter.zig:81-82 (empty node list is a no-op, PathNotClosed), surface.zig:1469-1547 (dimension validation),
fill_plotter.zig:50,53 / stroke_plotter.zig:113,134 (InvalidState), compositor.zig:3862-3877 (out-of-bounds
composition is a silent no-op).  Each case runs on the CPU oracle and, with -m gpu, through the C ABI."""
import numpy as np
import pytest

from tests import specs
from z2d_b200 import abi, host
from z2d_b200.abi import AntiAliasMode, Format, Operator


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def z(request):
    return specs.bind(request.getfixturevalue(request.param))


WHITE = host.Pattern.opaque(host.Pixel.rgb(0xFF, 0xFF, 0xFF))
NOT_PREMUL = host.Pattern.opaque(host.Pixel.rgba(0xFF, 0xFF, 0xFF, 0xAA))


def _closed_triangle(z):
    p = z.Path()
    p.move_to(1.0, 1.0)
    p.line_to(7.0, 1.0)
    p.line_to(4.0, 7.0)
    p.close()
    return p


def test_stroke_uninvertible_matrix_is_checked_before_empty_nodes(z):  # painter.zig:346-367
    sfc = z.Surface(Format.rgb, 1, 1)
    with pytest.raises(abi.InvalidMatrix):
        z.painter.stroke(sfc, WHITE, [], z.StrokeOptions(transformation=z.Transformation(1, 1, 2, 2, 5, 6)))


def test_fill_non_premultiplied_pixel(z):  # painter.zig:369-383
    sfc = z.Surface(Format.rgb, 1, 1)
    with pytest.raises(abi.PixelSourceNotPreMultiplied):
        z.painter.fill(sfc, NOT_PREMUL, [])


def test_stroke_non_premultiplied_pixel(z):  # painter.zig:385-398
    sfc = z.Surface(Format.rgb, 1, 1)
    with pytest.raises(abi.PixelSourceNotPreMultiplied):
        z.painter.stroke(sfc, NOT_PREMUL, [])


def test_empty_node_list_is_a_no_op(z):  # painter.zig:81, 234
    sfc = z.Surface(Format.rgba, 4, 4)
    z.painter.fill(sfc, WHITE, [])
    z.painter.stroke(sfc, WHITE, [])
    assert not sfc.download().any()


def test_fill_path_not_closed(z):  # painter.zig:82
    sfc = z.Surface(Format.rgba, 8, 8)
    p = z.Path()
    p.move_to(1.0, 1.0)
    p.line_to(7.0, 1.0)
    p.line_to(4.0, 7.0)
    with pytest.raises(abi.PathNotClosed):
        z.painter.fill(sfc, WHITE, p.nodes)
    assert not sfc.download().any(), "a failed call draws nothing"


def test_fill_after_failed_call_still_works(z):
    sfc = z.Surface(Format.rgba, 8, 8)
    p = z.Path()
    p.move_to(1.0, 1.0)
    p.line_to(7.0, 1.0)
    with pytest.raises(abi.PathNotClosed):
        z.painter.fill(sfc, WHITE, p.nodes)
    z.painter.fill(sfc, WHITE, _closed_triangle(z).nodes, z.FillOptions(anti_aliasing_mode=AntiAliasMode.none))
    assert sfc.pixels()[2, 4].tolist() == [255, 255, 255, 255]


def test_line_to_without_current_point_is_invalid_state(z):  # fill_plotter.zig:50; stroke_plotter.zig:113
    sfc = z.Surface(Format.rgba, 8, 8)
    nodes = [(1, 1.0, 1.0, 0, 0, 0, 0), (1, 5.0, 5.0, 0, 0, 0, 0), (3, 0, 0, 0, 0, 0, 0)]  # line_to, line_to, close_path
    with pytest.raises(abi.InvalidState):
        z.painter.fill(sfc, WHITE, nodes)
    with pytest.raises(abi.InvalidState):
        z.painter.stroke(sfc, WHITE, nodes)


@pytest.mark.parametrize("fmt", [Format.rgba, Format.rgb, Format.alpha8, Format.alpha4, Format.alpha2, Format.alpha1])
def test_surface_dimension_validation(z, fmt):  # surface.zig:1469-1547
    with pytest.raises(abi.InvalidWidth):
        z.Surface(fmt, 0, 10)
    with pytest.raises(abi.InvalidHeight):
        z.Surface(fmt, 10, 0)
    with pytest.raises(abi.InvalidWidth):
        z.Surface(fmt, -1, 10)
    with pytest.raises(abi.InvalidHeight):
        z.Surface(fmt, 10, -1)


def test_out_of_bounds_composite_is_a_no_op(z):  # compositor.zig:3862-3877 (and 311-345)
    dst = z.Surface(Format.rgba, 4, 4)
    src = z.SurfacePixel(host.Pixel.rgba(255, 0, 0, 255), 4, 4)
    for x, y in ((4, 0), (0, 4), (100, 100), (-4, 0), (0, -4)):
        z.SurfaceCompositor.run(dst, x, y, [z.Operation(Operator.src_over, src=z.Param.surface(src))])
    assert not dst.download().any()
    z.SurfaceCompositor.run(dst, 0, 0, [])  # zero operations
    z.SurfaceCompositor.run(dst, 0, 0, [z.Operation(Operator.src_over)])  # no source at all
    assert not dst.download().any()
    z.SurfaceCompositor.run(dst, 2, -1, [z.Operation(Operator.src_over, src=z.Param.surface(src))])  # partial overlap
    px = dst.pixels()
    assert px[:3, 2:, 0].min() == 255 and not px[3].any() and not px[:, :2].any()


def test_put_pixel_out_of_bounds_is_ignored(z):  # surface.zig:288
    sfc = z.Surface(Format.alpha4, 3, 3)
    for x, y in ((-1, 0), (0, -1), (3, 0), (0, 3)):
        sfc.put_pixel(x, y, host.Pixel.alpha4(15))
    assert not sfc.download().any()
    sfc.put_pixel(2, 2, host.Pixel.alpha4(15))
    assert int(sfc.pixels()[2, 2, 3]) == 15
