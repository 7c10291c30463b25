// Scenes written against include/z2d.hpp (the C++ mirror of z2d's API) -- the compiled-language host side of the drop-in
// boundary.  Each scene has a line-for-line Python twin in tests/test_gpu_cpp_host.py that is rendered by the CPU oracle; the raw
// surface bytes written here must match it.  usage: scenes <output directory>
#include <cstdio>
#include <fstream>
#include <string>

#include "z2d.hpp"

using namespace z2d;
static const double kPi = 3.14159265358979323846;

static void save(const std::string& dir, const std::string& name, Surface& sfc) {
  const std::vector<uint8_t> px = sfc.download();
  std::ofstream f(dir + "/" + name + ".bin", std::ios::binary);
  f.write(reinterpret_cast<const char*>(px.data()), (std::streamsize)px.size());
}

static void bezierFillRgba(const std::string& dir) {
  Surface sfc(Format::rgba, 300, 300);
  Context ctx(sfc);
  ctx.setSourceToPixel(Pixel::rgba(90, 40, 10, 128));
  ctx.setFillRule(FillRule::even_odd);
  ctx.moveTo(19, 249);
  ctx.curveTo(89, 49, 209, 49, 279, 249);
  ctx.curveTo(209, 149, 89, 149, 19, 20.5);
  ctx.closePath();
  ctx.moveTo(100, 100);
  ctx.lineTo(250.25, 120);
  ctx.lineTo(140, 280.75);
  ctx.closePath();
  ctx.fill();
  ctx.resetPath();
  ctx.setSourceToPixel(Pixel::rgba(0, 100, 200, 200));
  ctx.setOperator(Operator::multiply);
  ctx.setPrecision(Precision::float_);
  ctx.setFillRule(FillRule::non_zero);
  ctx.setAntiAliasingMode(AntiAliasMode::supersample_4x);
  ctx.moveTo(10, 150);
  ctx.lineTo(290, 130);
  ctx.lineTo(290, 190);
  ctx.lineTo(10, 170);
  ctx.closePath();
  ctx.fill();
  save(dir, "bezier_fill_rgba", sfc);
}

static void dashedStrokeArc(const std::string& dir) {
  Surface sfc(Format::rgb, 400, 400);
  Context ctx(sfc);
  ctx.setSourceToPixel(Pixel::rgb(0xFF, 0xFF, 0xFF));
  ctx.setLineWidth(6);
  ctx.setLineJoinMode(JoinMode::round);
  ctx.setLineCapMode(CapMode::round);
  ctx.setDashes({25, 10, 5, 10});
  ctx.setDashOffset(7.5);
  ctx.translate(200, 200);
  ctx.scale(150, 100);
  ctx.arc(0, 0, 1, 0, 2 * kPi);
  ctx.closePath();
  ctx.stroke();
  ctx.resetPath();
  ctx.setIdentity();
  ctx.setDashes({});
  ctx.setLineJoinMode(JoinMode::miter);
  ctx.setLineCapMode(CapMode::square);
  ctx.setSourceToPixel(Pixel::rgb(0x20, 0xC0, 0x40));
  ctx.rotate(0.25);
  ctx.moveTo(120, 20);
  ctx.lineTo(300, 60);
  ctx.relLineTo(-60, 120);
  ctx.relCurveTo(-30, 40, -90, 40, -120, 0);
  ctx.stroke();
  save(dir, "dashed_stroke_arc", sfc);
}

static void conicGradientAlpha4(const std::string& dir) {
  Surface sfc(Format::alpha4, 301, 299);
  Gradient g = Gradient::conic(149, 149, 0.5);
  g.addStop(0, Color::rgb(1, 0, 0, 1));
  g.addStop(0.5f, Color::rgb(0, 1, 0, 0.25f));
  g.addStop(1, Color::rgb(0, 0, 1, 1));
  Context ctx(sfc);
  ctx.setSource(Pattern::gradient(g));
  ctx.arc(149, 149, 120, 0, kPi * 2);
  ctx.closePath();
  ctx.fill();
  save(dir, "conic_gradient_alpha4", sfc);
}

static void compositorOps(const std::string& dir) {
  Surface dst(Pixel::rgba(40, 80, 120, 160), 128, 96);
  Gradient g = Gradient::linear(0, 0, 127, 95, InterpolationMethod::srgb);
  g.addStop(0, Color::rgb(1, 0, 0));
  g.addStop(0.5f, Color::rgb(0, 1, 0, 0.5f));
  g.addStop(1, Color::rgb(0, 0, 1));
  compositor::SurfaceCompositor::run(dst, 0, 0, {{Operator::xor_, compositor::Param::none(), compositor::Param::gradient(g)}},
                                     Precision::float_);
  Surface stamp(Pixel::rgba(100, 0, 50, 100), 64, 64);
  stamp.putPixel(3, 3, Pixel::rgba(255, 255, 255, 255));
  compositor::SurfaceCompositor::run(dst, -10, 5, {{Operator::src_over, compositor::Param::none(), compositor::Param::fromSurface(stamp)}});
  compositor::SurfaceCompositor::run(dst, 100, 70, {{Operator::plus, compositor::Param::none(), compositor::Param::fromSurface(stamp)}});
  save(dir, "compositor_ops", dst);
}

static void hairlineAndUnbounded(const std::string& dir) {
  Surface sfc(Pixel::rgba(10, 20, 30, 255), 200, 150);
  Context ctx(sfc);
  ctx.setSourceToPixel(Pixel::rgba(200, 100, 50, 200));
  ctx.setAntiAliasingMode(AntiAliasMode::none);
  ctx.setOperator(Operator::dst_in);
  ctx.moveTo(30, 20);
  ctx.lineTo(170, 40);
  ctx.lineTo(100, 130);
  ctx.closePath();
  ctx.fill();
  ctx.resetPath();
  ctx.setOperator(Operator::src_over);
  ctx.setAntiAliasingMode(AntiAliasMode::default_);
  ctx.setHairline(true);
  ctx.setSourceToPixel(Pixel::rgba(255, 255, 255, 255));
  ctx.moveTo(5, 5);
  ctx.lineTo(190, 140);
  ctx.lineTo(190, 10);
  ctx.curveTo(150, 60, 60, 60, 10, 140);
  ctx.stroke();
  save(dir, "hairline_unbounded", sfc);
}

static int errorBehaviour() {  // painter.zig:346-398, painter.zig:82, surface.zig:1469-1547
  int ok = 0;
  Surface sfc(Format::rgb, 4, 4);
  const Pattern white = Pattern::opaque(Pixel::rgb(255, 255, 255));
  try {
    StrokeOptions o;
    o.transformation = {1, 1, 2, 2, 5, 6};
    painter::stroke(sfc, white, {}, o);
  } catch (const InvalidMatrix&) { ok++; }
  try {
    painter::fill(sfc, Pattern::opaque(Pixel::rgba(255, 255, 255, 0xAA)), {});
  } catch (const PixelSourceNotPreMultiplied&) { ok++; }
  try {
    Path p;
    p.moveTo(0, 0);
    p.lineTo(3, 0);
    p.lineTo(3, 3);
    painter::fill(sfc, white, p.nodes);
  } catch (const PathNotClosed&) { ok++; }
  try {
    Surface bad(Format::rgba, 0, 10);
  } catch (const InvalidWidth&) { ok++; }
  try {
    Surface bad(Format::alpha2, 10, -1);
  } catch (const InvalidHeight&) { ok++; }
  painter::fill(sfc, white, {});  // empty node list: silent no-op
  return ok;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <output directory>\n", argv[0]);
    return 2;
  }
  const std::string dir = argv[1];
  bezierFillRgba(dir);
  dashedStrokeArc(dir);
  conicGradientAlpha4(dir);
  compositorOps(dir);
  hairlineAndUnbounded(dir);
  const int ok = errorBehaviour();
  std::printf("errors_ok=%d\n", ok);
  return ok == 5 ? 0 : 1;
}
