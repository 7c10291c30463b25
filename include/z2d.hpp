// z2d.hpp -- C++17 host-side mirror of z2d's public interface for the fill / stroke / composite path, over the C ABI of
// libz2d_cuda (z2d_cuda.h).  Header only.  Names, argument meaning and error behaviour follow the reference (file:line cited at
// each item, relative to the z2d tree) so that code written against z2d's Zig API translates line by line:
//
//   z2d::Surface sfc(z2d::Format::rgba, 300, 300);                // Surface.init                (surface.zig:97)
//   z2d::Context ctx(sfc);                                        // Context.init                (Context.zig:88)
//   ctx.setSourceToPixel(z2d::Pixel::rgb(0xFF, 0xFF, 0xFF));      // Context.setSourceToPixel    (Context.zig:124)
//   ctx.moveTo(10, 10); ctx.lineTo(200, 50); ctx.closePath();     // Path.moveTo / lineTo / close (Path.zig:124,175,453)
//   ctx.fill();                                                   // Context.fill -> painter.fill (Context.zig:592, painter.zig:66)
//   std::vector<uint8_t> px = sfc.download();                     // reads back what `sfc.image_surface_rgba.buf` holds
//
// Zig error unions become exceptions derived from z2d::Error.  Surfaces live on the device; there is no CPU fallback.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <optional>
#include <utility>
#include <vector>

#include "z2d_cuda.h"

namespace z2d {

// ------------------------------------------------------------------------------------------------ errors
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
#define Z2D_HPP_ERROR(NAME, CODE) \
  struct NAME : Error {           \
    NAME() : Error(CODE, #NAME) {} \
  };
Z2D_HPP_ERROR(PathNotClosed, Z2D_E_PATH_NOT_CLOSED)                              // painter.zig:57
Z2D_HPP_ERROR(PixelSourceNotPreMultiplied, Z2D_E_PIXEL_SOURCE_NOT_PREMULTIPLIED)  // painter.zig:62
Z2D_HPP_ERROR(InvalidWidth, Z2D_E_INVALID_WIDTH)                                 // surface.zig:87
Z2D_HPP_ERROR(InvalidHeight, Z2D_E_INVALID_HEIGHT)                               // surface.zig:90
Z2D_HPP_ERROR(InvalidState, Z2D_E_INVALID_STATE)                                 // internal/InternalError.zig
Z2D_HPP_ERROR(OutOfMemory, Z2D_E_OUT_OF_MEMORY)
Z2D_HPP_ERROR(InvalidMatrix, Z2D_E_INVALID_MATRIX)                               // Transformation.zig:27
Z2D_HPP_ERROR(InvalidArg, Z2D_E_INVALID_ARG)
Z2D_HPP_ERROR(NoCurrentPoint, -100)                                              // Path.zig:26 (host side only)
#undef Z2D_HPP_ERROR
struct DeviceError : Error {
  explicit DeviceError(const std::string& w) : Error(Z2D_E_DEVICE, w) {}
};

inline void check(int32_t rc, const z2d_ctx* ctx = nullptr) {
  switch (rc) {
    case Z2D_OK: return;
    case Z2D_E_PATH_NOT_CLOSED: throw PathNotClosed();
    case Z2D_E_PIXEL_SOURCE_NOT_PREMULTIPLIED: throw PixelSourceNotPreMultiplied();
    case Z2D_E_INVALID_WIDTH: throw InvalidWidth();
    case Z2D_E_INVALID_HEIGHT: throw InvalidHeight();
    case Z2D_E_INVALID_STATE: throw InvalidState();
    case Z2D_E_OUT_OF_MEMORY: throw OutOfMemory();
    case Z2D_E_INVALID_MATRIX: throw InvalidMatrix();
    case Z2D_E_INVALID_ARG: throw InvalidArg();
    default: throw DeviceError(ctx ? z2d_last_error(ctx) : "device error");
  }
}

// ------------------------------------------------------------------------------------------------ enums (reference order)
enum class Format : uint32_t { argb, xrgb, rgb, rgba, alpha8, alpha4, alpha2, alpha1 };  // pixel.zig:47-56
enum class Operator : uint32_t {                                                         // compositor.zig:46-155
  clear, src, dst, src_over, dst_over, src_in, dst_in, src_out, dst_out, src_atop, dst_atop, xor_, plus, multiply, screen, overlay,
  darken, lighten, color_dodge, color_burn, hard_light, soft_light, difference, exclusion, hue, saturation, color, luminosity
};
enum class Precision : uint32_t { integer, float_ };                                      // compositor.zig:214-217
enum class FillRule : uint32_t { non_zero, even_odd };                                    // options.zig
enum class JoinMode : uint32_t { miter, round, bevel };
enum class CapMode : uint32_t { butt, round, square };
enum class AntiAliasMode : uint32_t { none, default_, multisample_4x, supersample_4x };
enum class DitherType : uint32_t { none, bayer, blue_noise };                             // Dither.zig:28-36
enum class InterpolationMethod : uint32_t { linear_rgb, srgb, hsl };                      // color.zig
enum class Polar : uint32_t { shorter, longer, increasing, decreasing };
constexpr double default_tolerance = 0.1;                                                 // options.zig:12

// ------------------------------------------------------------------------------------------------ Transformation.zig
struct Transformation {
  double ax = 1, by = 0, cx = 0, dy = 1, tx = 0, ty = 0;
  static Transformation identity() { return {}; }
  Transformation mul(const Transformation& b) const {  // Transformation.zig:58-78
    return {ax * b.ax + by * b.cx, ax * b.by + by * b.dy, cx * b.ax + dy * b.cx, cx * b.by + dy * b.dy, ax * b.tx + by * b.ty + tx,
            cx * b.tx + dy * b.ty + ty};
  }
  double determinant() const { return ax * dy - by * cx; }
  Transformation inverse() const {  // Transformation.zig:103-162
    if (by == 0 && cx == 0) {
      if (ax == 0 || dy == 0) throw InvalidMatrix();
      if (ax != 1 || dy != 1) return {1 / ax, 0, 0, 1 / dy, -tx / ax, -ty / dy};
      return {1, 0, 0, 1, -tx, -ty};
    }
    const double det = determinant();
    if (det == 0) throw InvalidMatrix();
    const double k = 1 / det;
    return {dy * k, -by * k, -cx * k, ax * k, (by * ty - dy * tx) * k, (cx * tx - ax * ty) * k};
  }
  Transformation translate(double x, double y) const { return mul({1, 0, 0, 1, x, y}); }
  Transformation scale(double sx, double sy) const { return mul({sx, 0, 0, sy, 0, 0}); }
  Transformation rotate(double angle) const {
    const double s = std::sin(angle), c = std::cos(angle);
    return mul({c, -s, s, c, 0, 0});
  }
  void userToDeviceDistance(double& x, double& y) const {
    const double nx = ax * x + by * y, ny = cx * x + dy * y;
    x = nx;
    y = ny;
  }
  void userToDevice(double& x, double& y) const {
    userToDeviceDistance(x, y);
    x += tx;
    y += ty;
  }
  void deviceToUser(double& x, double& y) const { inverse().userToDevice(x, y); }
  void deviceToUserDistance(double& x, double& y) const { inverse().userToDeviceDistance(x, y); }
};

// ------------------------------------------------------------------------------------------------ internal/arc.zig
namespace detail {
inline double transformedCircleMajorAxis(const Transformation& m, double radius) {  // arc.zig:92-267
  const double eps = 0.00390625;
  const double det = m.ax * m.dy - m.by * m.cx;
  if (std::fabs(det * det - 1.0) < eps) {
    if (std::fabs(m.by) < eps && std::fabs(m.cx) < eps) return radius;
    if (std::fabs(m.ax) < eps && std::fabs(m.dy) < eps) return radius;
  }
  const double i = m.ax * m.ax + m.by * m.by, j = m.cx * m.cx + m.dy * m.dy;
  const double f = 0.5 * (i + j), g = 0.5 * (i - j), h = m.ax * m.cx + m.by * m.dy;
  return radius * std::sqrt(f + std::hypot(g, h));
}
inline double arcMaxAngle(double tolerance) {  // arc.zig:44-83
  static const double pi = 3.14159265358979323846;
  static const double table[11][2] = {{pi / 1.0, 0.0185185185185185036127},   {pi / 2.0, 0.000272567143730179811158},
                                      {pi / 3.0, 2.38647043651461047433e-05}, {pi / 4.0, 4.2455377443222443279e-06},
                                      {pi / 5.0, 1.11281001494389081528e-06}, {pi / 6.0, 3.72662000942734705475e-07},
                                      {pi / 7.0, 1.47783685574284411325e-07}, {pi / 8.0, 6.63240432022601149057e-08},
                                      {pi / 9.0, 3.2715520137536980553e-08},  {pi / 10.0, 1.73863223499021216974e-08},
                                      {pi / 11.0, 9.81410988043554039085e-09}};
  for (const auto& t : table)
    if (t[1] < tolerance) return t[0];
  double angle = 0;
  for (int i = 11; i < 1000; i++) {
    angle = pi / (double)i;
    const double err = 2.0 / 27.0 * std::pow(std::sin(angle / 4), 6) / std::pow(std::cos(angle / 4), 2);  // arc.zig:40-42
    if (err <= tolerance) break;
  }
  return angle;
}
}  // namespace detail

// ------------------------------------------------------------------------------------------------ Path.zig
using PathNode = z2d_node;  // internal/path_nodes.zig: {tag, p[6]} in DEVICE space

class Path {
 public:
  std::vector<PathNode> nodes;
  Transformation transformation;
  double tolerance = default_tolerance;
  bool has_current = false;
  double initial_x = 0, initial_y = 0, current_x = 0, current_y = 0;  // device space

  void reset() {
    nodes.clear();
    has_current = false;
  }
  void moveTo(double x, double y) {  // Path.zig:124-145
    dev(x, y);
    if (!nodes.empty() && nodes.back().tag == Z2D_NODE_MOVE_TO && nodes.back().p[0] == x && nodes.back().p[1] == y) return;
    push(Z2D_NODE_MOVE_TO, x, y);
    initial_x = current_x = x;
    initial_y = current_y = y;
    has_current = true;
  }
  void relMoveTo(double x, double y) {
    double ux, uy;
    user(ux, uy);
    moveTo(ux + x, uy + y);
  }
  void lineTo(double x, double y) {  // Path.zig:175-183
    if (!has_current) return moveTo(x, y);
    dev(x, y);
    push(Z2D_NODE_LINE_TO, x, y);
    current_x = x;
    current_y = y;
  }
  void relLineTo(double x, double y) {
    double ux, uy;
    user(ux, uy);
    lineTo(ux + x, uy + y);
  }
  void curveTo(double x1, double y1, double x2, double y2, double x3, double y3) {  // Path.zig:237-260
    if (!has_current) throw NoCurrentPoint();
    dev(x1, y1);
    dev(x2, y2);
    dev(x3, y3);
    PathNode n{};
    n.tag = Z2D_NODE_CURVE_TO;
    n.p[0] = x1; n.p[1] = y1; n.p[2] = x2; n.p[3] = y2; n.p[4] = x3; n.p[5] = y3;
    nodes.push_back(n);
    current_x = x3;
    current_y = y3;
  }
  void relCurveTo(double x1, double y1, double x2, double y2, double x3, double y3) {
    double ux, uy;
    user(ux, uy);
    curveTo(ux + x1, uy + y1, ux + x2, uy + y2, ux + x3, uy + y3);
  }
  void arc(double xc, double yc, double radius, double angle1, double angle2) {  // Path.zig:304-332
    while (angle2 < angle1) angle2 += kPi * 2;
    arcInDirection(xc, yc, radius, angle1, angle2, true, std::max(tolerance, 0.001));
  }
  void arcNegative(double xc, double yc, double radius, double angle1, double angle2) {  // Path.zig:334-362
    while (angle2 > angle1) angle2 -= kPi * 2;
    arcInDirection(xc, yc, radius, angle2, angle1, false, std::max(tolerance, 0.001));
  }
  void close() {  // Path.zig:453-476: close_path + explicit move_to(initial point)
    if (!has_current) return;
    push(Z2D_NODE_CLOSE_PATH, 0, 0);
    push(Z2D_NODE_MOVE_TO, initial_x, initial_y);
  }
  bool isClosed() const {  // path_nodes.zig:23-37
    if (nodes.empty()) return false;
    bool closed = false;
    for (size_t i = 0; i < nodes.size(); i++) {
      if (nodes[i].tag == Z2D_NODE_MOVE_TO) {
        if (!closed && i != 0) break;
      } else {
        closed = nodes[i].tag == Z2D_NODE_CLOSE_PATH;
      }
    }
    return closed;
  }

 private:
  static constexpr double kPi = 3.14159265358979323846;
  void push(uint32_t tag, double x, double y) {
    PathNode n{};
    n.tag = tag;
    n.p[0] = x;
    n.p[1] = y;
    nodes.push_back(n);
  }
  void dev(double& x, double& y) const {  // clamp to the i24 range, then user -> device (Path.zig:131-134)
    x = std::max(-8388608.0, std::min(x, 8388607.0));
    y = std::max(-8388608.0, std::min(y, 8388607.0));
    transformation.userToDevice(x, y);
  }
  void user(double& ux, double& uy) const {
    if (!has_current) throw NoCurrentPoint();
    ux = current_x;
    uy = current_y;
    transformation.deviceToUser(ux, uy);
  }
  void arcLineTo(double x, double y) {  // Path.zig:365-384
    if (has_current && current_x == x && current_y == y) return;
    lineTo(x, y);
  }
  void arcSegment(double xc, double yc, double radius, double a, double b) {  // arc.zig:294-318
    const double rsa = radius * std::sin(a), rca = radius * std::cos(a), rsb = radius * std::sin(b), rcb = radius * std::cos(b);
    const double h = 4.0 / 3.0 * std::tan((b - a) / 4.0);
    curveTo(xc + rca - h * rsa, yc + rsa + h * rca, xc + rcb + h * rsb, yc + rsb - h * rcb, xc + rcb, yc + rsb);
  }
  void arcInDirection(double xc, double yc, double radius, double amin, double amax, bool forward, double tol) {  // arc.zig:330-392
    if (!(amax * amax >= 0.0) || !(amin * amin >= 0.0)) return;
    const double max_full = 65536;
    if (amax - amin > 2 * kPi * max_full) {
      amax = std::fmod(amax - amin, 2 * kPi);
      amin = std::fmod(amin, 2 * kPi);
      if (amin < 0) amin += 2 * kPi;
      amax += amin + 2 * kPi * max_full;
    }
    if (amax - amin > kPi) {
      const double amid = amin + (amax - amin) / 2.0;
      if (forward) {
        arcInDirection(xc, yc, radius, amin, amid, forward, tol);
        arcInDirection(xc, yc, radius, amid, amax, forward, tol);
      } else {
        arcInDirection(xc, yc, radius, amid, amax, forward, tol);
        arcInDirection(xc, yc, radius, amin, amid, forward, tol);
      }
    } else if (amax != amin) {
      const double major = detail::transformedCircleMajorAxis(transformation, radius);
      int segments = (int)std::ceil(std::fabs(amax - amin) / detail::arcMaxAngle(tol / major));
      double step = (amax - amin) / (double)segments;
      segments -= 1;
      if (!forward) {
        std::swap(amin, amax);
        step = -step;
      }
      arcLineTo(xc + radius * std::cos(amin), yc + radius * std::sin(amin));
      for (int i = 0; i < segments; i++) {
        arcSegment(xc, yc, radius, amin, amin + step);
        amin += step;
      }
      arcSegment(xc, yc, radius, amin, amax);
    } else {
      arcLineTo(xc + radius * std::cos(amin), yc + radius * std::sin(amin));
    }
  }
};

// ------------------------------------------------------------------------------------------------ pixel.zig / color.zig
struct Color {  // color.Color.init (color.zig:58-69): clamped, de-multiplied
  z2d_color pod{};
  static Color rgb(float r, float g, float b, float a = 1) { return make(Z2D_COLOR_LINEAR_RGB, r, g, b, a); }
  static Color srgb(float r, float g, float b, float a = 1) { return make(Z2D_COLOR_SRGB, r, g, b, a); }
  static Color hsl(float h, float s, float l, float a = 1) {
    if (h < 0 || h > 360) {  // color.zig:392-394 (@mod: floored)
      h = std::fmod(h, 360.0f);
      if (h < 0) h += 360.0f;
    }
    Color c = make(Z2D_COLOR_HSL, 0, s, l, a);
    c.pod.c[0] = h;
    return c;
  }

 private:
  static float clamp01(float v) { return std::max(0.0f, std::min(v, 1.0f)); }
  static Color make(uint32_t space, float a, float b, float c, float d) {
    Color r;
    r.pod.space = space;
    r.pod.c[0] = clamp01(a);
    r.pod.c[1] = clamp01(b);
    r.pod.c[2] = clamp01(c);
    r.pod.c[3] = clamp01(d);
    return r;
  }
};

struct Pixel {  // pixel.Pixel (pixel.zig:100-140): channel values as stored by the format
  z2d_pixel pod{};
  static Pixel make(Format f, uint8_t r, uint8_t g, uint8_t b, uint8_t a) {
    Pixel p;
    p.pod.format = (uint32_t)f;
    p.pod.r = r; p.pod.g = g; p.pod.b = b; p.pod.a = a;
    return p;
  }
  static Pixel rgb(uint8_t r, uint8_t g, uint8_t b) { return make(Format::rgb, r, g, b, 255); }
  static Pixel xrgb(uint8_t r, uint8_t g, uint8_t b) { return make(Format::xrgb, r, g, b, 255); }
  static Pixel rgba(uint8_t r, uint8_t g, uint8_t b, uint8_t a) { return make(Format::rgba, r, g, b, a); }
  static Pixel argb(uint8_t r, uint8_t g, uint8_t b, uint8_t a) { return make(Format::argb, r, g, b, a); }
  static Pixel alpha8(uint8_t a) { return make(Format::alpha8, 0, 0, 0, a); }
  static Pixel alpha4(uint8_t a) { return make(Format::alpha4, 0, 0, 0, a); }
  static Pixel alpha2(uint8_t a) { return make(Format::alpha2, 0, 0, 0, a); }
  static Pixel alpha1(uint8_t a) { return make(Format::alpha1, 0, 0, 0, a); }
  // Pixel.fromColor for linear colours (pixel.zig:115-117, color.zig:214-232: round, then integer premultiply)
  static Pixel fromColor(const Color& c) {
    auto enc = [](float v) { return (int)std::round(255.0f * v); };
    const int r = enc(c.pod.c[0]), g = enc(c.pod.c[1]), b = enc(c.pod.c[2]), a = enc(c.pod.c[3]);
    return make(Format::rgba, (uint8_t)(r * a / 255), (uint8_t)(g * a / 255), (uint8_t)(b * a / 255), (uint8_t)a);
  }
};

// ------------------------------------------------------------------------------------------------ gradient.zig / Dither.zig / pattern.zig
class Gradient {
 public:
  static Gradient linear(double x0, double y0, double x1, double y1, InterpolationMethod m = InterpolationMethod::linear_rgb,
                         Polar p = Polar::shorter) {
    return Gradient(Z2D_GRADIENT_LINEAR, {x0, y0, x1, y1, 0, 0}, m, p);
  }
  static Gradient radial(double ix, double iy, double ir, double ox, double oy, double orad,
                         InterpolationMethod m = InterpolationMethod::linear_rgb, Polar p = Polar::shorter) {
    return Gradient(Z2D_GRADIENT_RADIAL, {ix, iy, ir, ox, oy, orad}, m, p);
  }
  static Gradient conic(double x, double y, double angle, InterpolationMethod m = InterpolationMethod::linear_rgb, Polar p = Polar::shorter) {
    return Gradient(Z2D_GRADIENT_CONIC, {x, y, angle, 0, 0, 0}, m, p);
  }
  void addStop(float offset, const Color& color) {  // Stop.List.add (gradient.zig:797-811): sorted by offset, ties by insertion
    z2d_stop s{};
    s.offset = std::max(0.0f, std::min(offset, 1.0f));
    s.color = color.pod;
    auto it = std::upper_bound(stops_.begin(), stops_.end(), s, [](const z2d_stop& a, const z2d_stop& b) { return a.offset < b.offset; });
    stops_.insert(it, s);
  }
  void setTransformation(const Transformation& t) { inv_ = t.inverse(); }  // gradient.zig:201-203: stored inverted
  z2d_gradient pod() const {
    z2d_gradient g{};
    g.type = type_;
    g.method = (uint32_t)method_;
    g.polar = (uint32_t)polar_;
    g.n_stops = (uint32_t)stops_.size();
    std::memcpy(g.geom, geom_, sizeof geom_);
    const double t[6] = {inv_.ax, inv_.by, inv_.cx, inv_.dy, inv_.tx, inv_.ty};
    std::memcpy(g.inv_ctm, t, sizeof t);
    g.stops = stops_.data();
    return g;
  }

 private:
  Gradient(uint32_t type, std::initializer_list<double> geom, InterpolationMethod m, Polar p) : type_(type), method_(m), polar_(p) {
    std::copy(geom.begin(), geom.end(), geom_);
  }
  uint32_t type_;
  InterpolationMethod method_;
  Polar polar_;
  double geom_[6] = {0, 0, 0, 0, 0, 0};
  Transformation inv_;
  std::vector<z2d_stop> stops_;
};

class Pattern {  // pattern.Pattern (pattern.zig:32-44): opaque pixel | gradient | dither.  Gradients are borrowed (as in z2d).
 public:
  static Pattern opaque(const Pixel& px) {
    Pattern p;
    p.pod_.kind = Z2D_PATTERN_OPAQUE;
    p.pod_.pixel = px.pod;
    return p;
  }
  static Pattern gradient(const Gradient& g) {
    Pattern p;
    p.pod_.kind = Z2D_PATTERN_GRADIENT;
    p.grad_ = &g;
    return p;
  }
  static Pattern ditherPixel(DitherType t, const Pixel& px, uint32_t scale) {  // Dither.zig:28-58
    Pattern p = opaque(px);
    p.pod_.kind = Z2D_PATTERN_DITHER;
    p.pod_.dither_type = (uint32_t)t;
    p.pod_.dither_source = Z2D_DITHER_SRC_PIXEL;
    p.pod_.dither_scale = scale;
    return p;
  }
  static Pattern ditherGradient(DitherType t, const Gradient& g, uint32_t scale) {
    Pattern p = gradient(g);
    p.pod_.kind = Z2D_PATTERN_DITHER;
    p.pod_.dither_type = (uint32_t)t;
    p.pod_.dither_source = Z2D_DITHER_SRC_GRADIENT;
    p.pod_.dither_scale = scale;
    return p;
  }
  bool isOpaquePixel() const { return pod_.kind == Z2D_PATTERN_OPAQUE; }
  bool isGradient() const { return pod_.kind == Z2D_PATTERN_GRADIENT; }
  const Pixel pixel() const {
    Pixel px;
    px.pod = pod_.pixel;
    return px;
  }
  const Gradient* gradientPtr() const { return grad_; }
  // POD for one call; `scratch` receives the gradient POD the pattern points to
  z2d_pattern pod(z2d_gradient& scratch) const {
    z2d_pattern p = pod_;
    if (grad_) {
      scratch = grad_->pod();
      p.gradient = &scratch;
    }
    return p;
  }

 private:
  z2d_pattern pod_{};
  const Gradient* grad_ = nullptr;
};

// ------------------------------------------------------------------------------------------------ device + surface.zig
class Device {  // one z2d_ctx (no reference equivalent: z2d has no device)
 public:
  explicit Device(int device = 0, void* stream = nullptr) { check(z2d_ctx_create(device, stream, &ctx_)); }
  ~Device() { z2d_ctx_destroy(ctx_); }
  Device(const Device&) = delete;
  Device& operator=(const Device&) = delete;
  z2d_ctx* handle() const { return ctx_; }
  void flush() { check(z2d_flush(ctx_), ctx_); }
  void sync() { check(z2d_sync(ctx_), ctx_); }
  static Device& instance() {  // process-wide default, created on first use
    static Device d;
    return d;
  }

 private:
  z2d_ctx* ctx_ = nullptr;
};

class Surface {
 public:
  Surface(Format format, int32_t width, int32_t height, Device& dev = Device::instance()) : dev_(&dev) {  // Surface.init (surface.zig:97)
    check(z2d_surface_create(dev.handle(), (uint32_t)format, width, height, nullptr, &sfc_), dev.handle());
  }
  Surface(const Pixel& px, int32_t width, int32_t height, Device& dev = Device::instance()) : dev_(&dev) {  // Surface.initPixel (surface.zig:128)
    check(z2d_surface_create(dev.handle(), px.pod.format, width, height, &px.pod, &sfc_), dev.handle());
  }
  ~Surface() { z2d_surface_destroy(sfc_); }
  Surface(const Surface&) = delete;
  Surface& operator=(const Surface&) = delete;
  int32_t getWidth() const { return z2d_surface_width(sfc_); }
  int32_t getHeight() const { return z2d_surface_height(sfc_); }
  Format getFormat() const { return (Format)z2d_surface_format(sfc_); }
  void paintPixel(const Pixel& px) { check(z2d_surface_paint_pixel(sfc_, &px.pod), dev_->handle()); }               // surface.zig:295
  void putPixel(int32_t x, int32_t y, const Pixel& px) { check(z2d_surface_put_pixel(sfc_, x, y, &px.pod), dev_->handle()); }  // surface.zig:288
  std::optional<Pixel> getPixel(int32_t x, int32_t y) {  // surface.zig:280: nullopt outside the surface
    Pixel px;
    const int32_t rc = z2d_surface_get_pixel(sfc_, x, y, &px.pod);
    if (rc == 1) return std::nullopt;
    check(rc, dev_->handle());
    return px;
  }
  // export_png.zig:150-373: the scanline bytes a PNG holds before zlib (de-multiplied, optional sRGB curve, packed greys
  // most-significant-first), produced on the device.  flags: Z2D_EXPORT_SRGB | Z2D_EXPORT_FILTER_BYTE.
  std::vector<uint8_t> exportRows(uint32_t flags = 0) {
    std::vector<uint8_t> out(z2d_surface_export_size(sfc_, flags));
    check(z2d_surface_export(sfc_, flags, out.data(), out.size()), dev_->handle());
    return out;
  }
  std::vector<uint8_t> download() {  // the bytes of the reference's `buf` slice (flushes and waits)
    std::vector<uint8_t> out(z2d_surface_byte_len(sfc_));
    check(z2d_surface_download(sfc_, out.data(), out.size()), dev_->handle());
    return out;
  }
  void upload(const std::vector<uint8_t>& bytes) { check(z2d_surface_upload(sfc_, bytes.data(), bytes.size()), dev_->handle()); }
  z2d_sfc* handle() const { return sfc_; }
  Device& device() const { return *dev_; }

 private:
  Device* dev_;
  z2d_sfc* sfc_ = nullptr;
};

// ------------------------------------------------------------------------------------------------ painter.zig
struct FillOptions {  // painter.zig:28-48
  AntiAliasMode anti_aliasing_mode = AntiAliasMode::default_;
  FillRule fill_rule = FillRule::non_zero;
  Operator op = Operator::src_over;
  Precision precision = Precision::integer;
  double tolerance = default_tolerance;
};
struct StrokeOptions {  // painter.zig:145-198
  AntiAliasMode anti_aliasing_mode = AntiAliasMode::default_;
  std::vector<double> dashes;
  double dash_offset = 0;
  CapMode line_cap_mode = CapMode::butt;
  JoinMode line_join_mode = JoinMode::miter;
  double line_width = 2.0;
  double miter_limit = 10.0;
  Operator op = Operator::src_over;
  Precision precision = Precision::integer;
  double tolerance = default_tolerance;
  Transformation transformation;
  bool hairline = false;
};

namespace painter {
inline void fill(Surface& sfc, const Pattern& pattern, const std::vector<PathNode>& nodes, const FillOptions& o = {}) {  // painter.zig:66
  z2d_gradient g;
  const z2d_pattern p = pattern.pod(g);
  const z2d_fill_opts fo{(uint32_t)o.anti_aliasing_mode, (uint32_t)o.fill_rule, (uint32_t)o.op, (uint32_t)o.precision, o.tolerance};
  check(z2d_fill(sfc.device().handle(), sfc.handle(), &p, nodes.data(), nodes.size(), &fo), sfc.device().handle());
}
inline void stroke(Surface& sfc, const Pattern& pattern, const std::vector<PathNode>& nodes, const StrokeOptions& o = {}) {  // painter.zig:214
  z2d_gradient g;
  const z2d_pattern p = pattern.pod(g);
  z2d_stroke_opts so{};
  so.anti_aliasing_mode = (uint32_t)o.anti_aliasing_mode;
  so.line_cap_mode = (uint32_t)o.line_cap_mode;
  so.line_join_mode = (uint32_t)o.line_join_mode;
  so.op = (uint32_t)o.op;
  so.precision = (uint32_t)o.precision;
  so.hairline = o.hairline ? 1u : 0u;
  so.line_width = o.line_width;
  so.miter_limit = o.miter_limit;
  so.tolerance = o.tolerance;
  so.dash_offset = o.dash_offset;
  so.dashes = o.dashes.data();
  so.n_dashes = o.dashes.size();
  const Transformation& t = o.transformation;
  const double m[6] = {t.ax, t.by, t.cx, t.dy, t.tx, t.ty};
  std::memcpy(so.ctm, m, sizeof m);
  check(z2d_stroke(sfc.device().handle(), sfc.handle(), &p, nodes.data(), nodes.size(), &so), sfc.device().handle());
}
}  // namespace painter

// ------------------------------------------------------------------------------------------------ compositor.zig (surface level)
namespace compositor {
struct Param {  // SurfaceCompositor.Operation.Param (compositor.zig:232-283)
  uint32_t kind = Z2D_PARAM_NONE;
  Pattern pattern = Pattern::opaque(Pixel::rgba(0, 0, 0, 0));
  const Surface* surface = nullptr;
  static Param none() { return {}; }
  static Param pixel(const Pixel& px) { return {Z2D_PARAM_PIXEL, Pattern::opaque(px), nullptr}; }
  static Param gradient(const Gradient& g) { return {Z2D_PARAM_GRADIENT, Pattern::gradient(g), nullptr}; }
  static Param dither(const Pattern& d) { return {Z2D_PARAM_DITHER, d, nullptr}; }
  static Param fromSurface(const Surface& s) { return {Z2D_PARAM_SURFACE, Pattern::opaque(Pixel::rgba(0, 0, 0, 0)), &s}; }
};
struct Operation {  // compositor.zig:220-230
  Operator op;
  Param dst = Param::none();
  Param src = Param::none();
};
struct SurfaceCompositor {
  static void run(Surface& dst, int32_t dst_x, int32_t dst_y, const std::vector<Operation>& ops,
                  Precision precision = Precision::integer) {  // compositor.zig:302-309
    std::vector<z2d_comp_op> pods(ops.size());
    std::vector<z2d_gradient> grads(ops.size() * 2);
    for (size_t i = 0; i < ops.size(); i++) {
      pods[i].op = (uint32_t)ops[i].op;
      pack(ops[i].dst, pods[i].dst, grads[2 * i]);
      pack(ops[i].src, pods[i].src, grads[2 * i + 1]);
    }
    check(z2d_composite(dst.device().handle(), dst.handle(), dst_x, dst_y, pods.data(), pods.size(), (uint32_t)precision),
          dst.device().handle());
  }

 private:
  static void pack(const Param& p, z2d_comp_param& out, z2d_gradient& scratch) {
    out = z2d_comp_param{};
    out.kind = p.kind;
    if (p.kind == Z2D_PARAM_SURFACE) out.surface = p.surface->handle();
    else if (p.kind != Z2D_PARAM_NONE) out.pattern = p.pattern.pod(scratch);
  }
};
}  // namespace compositor

// ------------------------------------------------------------------------------------------------ Context.zig
class Context {
 public:
  explicit Context(Surface& surface) : surface_(&surface) {}  // Context.init (Context.zig:88-113)
  // --- source and options (Context.zig:115-343)
  void setSource(const Pattern& p) { pattern_ = p; }
  void setSourceToPixel(const Pixel& px) { pattern_ = Pattern::opaque(px); }
  void setAntiAliasingMode(AntiAliasMode m) { anti_aliasing_mode_ = m; }
  void setDashes(std::vector<double> d) { dashes_ = std::move(d); }
  void setDashOffset(double o) { dash_offset_ = o; }
  void setDither(DitherType d) { dither_ = d; }
  void setFillRule(FillRule r) { fill_rule_ = r; }
  void setHairline(bool h) { hairline_ = h; }
  void setLineCapMode(CapMode m) { line_cap_mode_ = m; }
  void setLineJoinMode(JoinMode m) { line_join_mode_ = m; }
  void setLineWidth(double w) { line_width_ = w; }
  void setMiterLimit(double m) { miter_limit_ = m; }
  void setOperator(Operator op) { operator_ = op; }
  void setPrecision(Precision p) { precision_ = p; }
  void setTolerance(double t) {  // Context.zig:303-307
    tolerance_ = t;
    path_.tolerance = t;
  }
  // --- transformation (Context.zig:345-420)
  const Transformation& getTransformation() const { return transformation_; }
  void setTransformation(const Transformation& t) {
    transformation_ = t;
    path_.transformation = t;
  }
  void setIdentity() { setTransformation(Transformation::identity()); }
  void mul(const Transformation& a) { setTransformation(transformation_.mul(a)); }
  void translate(double tx, double ty) { setTransformation(transformation_.translate(tx, ty)); }
  void rotate(double angle) { setTransformation(transformation_.rotate(angle)); }
  void scale(double sx, double sy) { setTransformation(transformation_.scale(sx, sy)); }
  // --- path (Context.zig:422-590)
  void resetPath() { path_.reset(); }
  void moveTo(double x, double y) { path_.moveTo(x, y); }
  void relMoveTo(double x, double y) { path_.relMoveTo(x, y); }
  void lineTo(double x, double y) { path_.lineTo(x, y); }
  void relLineTo(double x, double y) { path_.relLineTo(x, y); }
  void curveTo(double x1, double y1, double x2, double y2, double x3, double y3) { path_.curveTo(x1, y1, x2, y2, x3, y3); }
  void relCurveTo(double x1, double y1, double x2, double y2, double x3, double y3) { path_.relCurveTo(x1, y1, x2, y2, x3, y3); }
  void arc(double xc, double yc, double r, double a1, double a2) { path_.arc(xc, yc, r, a1, a2); }
  void arcNegative(double xc, double yc, double r, double a1, double a2) { path_.arcNegative(xc, yc, r, a1, a2); }
  void closePath() { path_.close(); }
  const Path& path() const { return path_; }
  // --- drawing (Context.zig:592-641)
  void fill() {
    painter::fill(*surface_, wrapDither(), path_.nodes, FillOptions{anti_aliasing_mode_, fill_rule_, operator_, precision_, tolerance_});
  }
  void stroke() {
    StrokeOptions o;
    o.anti_aliasing_mode = anti_aliasing_mode_;
    o.dashes = dashes_;
    o.dash_offset = dash_offset_;
    o.line_cap_mode = line_cap_mode_;
    o.line_join_mode = line_join_mode_;
    o.line_width = line_width_;
    o.miter_limit = miter_limit_;
    o.op = operator_;
    o.precision = precision_;
    o.tolerance = tolerance_;
    o.transformation = transformation_;
    o.hairline = hairline_;
    painter::stroke(*surface_, wrapDither(), path_.nodes, o);
  }

 private:
  Pattern wrapDither() const {  // Context.zig:679-699
    if (dither_ == DitherType::none) return pattern_;
    const Format f = surface_->getFormat();
    const uint32_t scale = f == Format::alpha1 ? 1 : f == Format::alpha2 ? 2 : f == Format::alpha4 ? 4 : 8;
    if (pattern_.isOpaquePixel()) return Pattern::ditherPixel(dither_, pattern_.pixel(), scale);
    if (pattern_.isGradient()) return Pattern::ditherGradient(dither_, *pattern_.gradientPtr(), scale);
    return pattern_;
  }
  Surface* surface_;
  Path path_;
  Pattern pattern_ = Pattern::opaque(Pixel::rgba(0, 0, 0, 255));  // Context.zig:95: opaque black
  AntiAliasMode anti_aliasing_mode_ = AntiAliasMode::default_;
  std::vector<double> dashes_;
  double dash_offset_ = 0;
  DitherType dither_ = DitherType::none;
  FillRule fill_rule_ = FillRule::non_zero;
  bool hairline_ = false;
  CapMode line_cap_mode_ = CapMode::butt;
  JoinMode line_join_mode_ = JoinMode::miter;
  double line_width_ = 2.0;
  double miter_limit_ = 10.0;
  Operator operator_ = Operator::src_over;
  Precision precision_ = Precision::integer;
  double tolerance_ = default_tolerance;
  Transformation transformation_;
};

}  // namespace z2d
