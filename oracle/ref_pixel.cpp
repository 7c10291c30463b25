// ORACLE -- test infrastructure only (see ref_internal.h).
// Pixel formats and surface primitives: src/pixel.zig, src/surface.zig.
#include "ref_internal.h"

namespace zref {

static int fmt_bits(uint32_t fmt) {
  switch (fmt) {
    case Z2D_FMT_ALPHA8: return 8;
    case Z2D_FMT_ALPHA4: return 4;
    case Z2D_FMT_ALPHA2: return 2;
    case Z2D_FMT_ALPHA1: return 1;
    default: return 32;
  }
}

size_t sfc_byte_len(uint32_t fmt, int64_t w, int64_t h) {
  // surface.zig:391-394 (w*h*4 / w*h) and 632 ((h*w*bits + 7) / 8)
  return (size_t)((w * h * fmt_bits(fmt) + 7) / 8);
}

// Alpha(T).shlr (pixel.zig:587-626): scale an alpha value between bit widths,
// bit-replicating on the way up, truncating on the way down.
static int scale_alpha(int val, int from_bits, int to_bits) {
  if (from_bits == 1) return val * ((1 << to_bits) - 1);  // pixel.zig:583
  if (val == 0) return 0;
  if (from_bits == to_bits) return val;
  if (from_bits > to_bits) return val >> (from_bits - to_bits);
  int diff = to_bits - from_bits;
  if (diff == 2 || diff == 4) return (val << diff) + val;  // u2->u4, u4->u8
  // diff == 6: u2 -> u8
  return (val << (to_bits - from_bits)) | (val << (to_bits - 2 * from_bits)) | (val << (to_bits - 3 * from_bits)) | val;
}

static int px_alpha_bits(uint32_t fmt) {
  switch (fmt) {
    case Z2D_FMT_ALPHA4: return 4;
    case Z2D_FMT_ALPHA2: return 2;
    case Z2D_FMT_ALPHA1: return 1;
    default: return 8;
  }
}

// pixel.RGBA.fromPixel (pixel.zig:399-433) widened to RGBA16 (compositor.zig:675-687).
RGBA16 px_to_rgba16(const z2d_pixel& px) {
  switch (px.format) {
    case Z2D_FMT_XRGB:
    case Z2D_FMT_RGB: return {px.r, px.g, px.b, 255};
    case Z2D_FMT_ARGB:
    case Z2D_FMT_RGBA: return {px.r, px.g, px.b, px.a};
    default: return {0, 0, 0, scale_alpha(px.a, px_alpha_bits(px.format), 8)};
  }
}

bool px_is_opaque(const z2d_pixel& px) {  // pixel.zig:128-137
  switch (px.format) {
    case Z2D_FMT_XRGB:
    case Z2D_FMT_RGB: return true;
    case Z2D_FMT_ARGB:
    case Z2D_FMT_RGBA:
    case Z2D_FMT_ALPHA8: return px.a == 255;
    case Z2D_FMT_ALPHA4: return px.a == 15;
    case Z2D_FMT_ALPHA2: return px.a == 3;
    default: return px.a == 1;
  }
}

bool px_can_demultiply(const z2d_pixel& px) {  // pixel.zig:504-514
  if (px.format != Z2D_FMT_ARGB && px.format != Z2D_FMT_RGBA) return true;
  if (px.a == 0) return true;
  const int c[3] = {px.r, px.g, px.b};
  for (int v : c)
    if (v * 255 / px.a > 255) return false;
  return true;
}

// mem.readPackedInt / writePackedInt, little endian bit order (pixel.zig:664-677)
static int packed_get(const uint8_t* buf, size_t idx, int bits) {
  size_t bit = idx * (size_t)bits;
  return (buf[bit >> 3] >> (bit & 7)) & ((1 << bits) - 1);
}
static void packed_set(uint8_t* buf, size_t idx, int bits, int v) {
  size_t bit = idx * (size_t)bits;
  int sh = (int)(bit & 7);
  int mask = ((1 << bits) - 1) << sh;
  buf[bit >> 3] = (uint8_t)((buf[bit >> 3] & ~mask) | ((v << sh) & mask));
}

RGBA16 sfc_load(const Sfc& s, size_t idx) {
  const uint8_t* p = s.buf + idx * 4;
  switch (s.fmt) {
    case Z2D_FMT_ARGB: return {p[2], p[1], p[0], p[3]};  // packed struct {b,g,r,a}
    case Z2D_FMT_XRGB: return {p[2], p[1], p[0], 255};
    case Z2D_FMT_RGB: return {p[0], p[1], p[2], 255};
    case Z2D_FMT_RGBA: return {p[0], p[1], p[2], p[3]};
    case Z2D_FMT_ALPHA8: return {0, 0, 0, s.buf[idx]};
    default: {
      int bits = fmt_bits(s.fmt);
      return {0, 0, 0, scale_alpha(packed_get(s.buf, idx, bits), bits, 8)};
    }
  }
}

void sfc_store(Sfc& s, size_t idx, RGBA16 v) {
  uint8_t* p = s.buf + idx * 4;
  switch (s.fmt) {
    case Z2D_FMT_ARGB: p[0] = (uint8_t)v.b; p[1] = (uint8_t)v.g; p[2] = (uint8_t)v.r; p[3] = (uint8_t)v.a; break;
    case Z2D_FMT_XRGB: p[0] = (uint8_t)v.b; p[1] = (uint8_t)v.g; p[2] = (uint8_t)v.r; p[3] = 0; break;
    case Z2D_FMT_RGB: p[0] = (uint8_t)v.r; p[1] = (uint8_t)v.g; p[2] = (uint8_t)v.b; p[3] = 0; break;
    case Z2D_FMT_RGBA: p[0] = (uint8_t)v.r; p[1] = (uint8_t)v.g; p[2] = (uint8_t)v.b; p[3] = (uint8_t)v.a; break;
    case Z2D_FMT_ALPHA8: s.buf[idx] = (uint8_t)v.a; break;
    default: {
      int bits = fmt_bits(s.fmt);
      packed_set(s.buf, idx, bits, ((uint8_t)v.a) >> (8 - bits));  // compositor.zig:693-696
    }
  }
}

// T.fromPixel(px) written to buf[idx] (pixel.zig:399-433, 569-577)
void sfc_paint(Sfc& s, size_t idx, const z2d_pixel& px) {
  if (s.fmt <= Z2D_FMT_RGBA) {
    RGBA16 v = px_to_rgba16(px);
    sfc_store(s, idx, v);  // rgb/xrgb drop alpha, padding 0
    return;
  }
  int to_bits = fmt_bits(s.fmt);
  int a;
  if (px.format == Z2D_FMT_XRGB || px.format == Z2D_FMT_RGB)
    a = (1 << to_bits) - 1;  // pixel.zig:571-574: RGB is always opaque
  else
    a = scale_alpha(px.a, px_alpha_bits(px.format), to_bits);
  if (s.fmt == Z2D_FMT_ALPHA8)
    s.buf[idx] = (uint8_t)a;
  else
    packed_set(s.buf, idx, to_bits, a);
}

void sfc_paint_stride(Sfc& s, int x, int y, size_t len, const z2d_pixel& px) {
  if (x < 0 || y < 0 || x >= s.w || y >= s.h) return;  // surface.zig:538,794
  size_t start = (size_t)s.w * (size_t)y + (size_t)x;
  for (size_t i = 0; i < len; i++) sfc_paint(s, start + i, px);
}

void sfc_clear_stride(Sfc& s, int x, int y, size_t len) {
  z2d_pixel zero{Z2D_FMT_RGBA, 0, 0, 0, 0};
  sfc_paint_stride(s, x, y, len, zero);
}

void sfc_composite_stride(Sfc& s, int x, int y, size_t len, const z2d_pixel& px, uint32_t op, uint8_t opacity) {
  if (x < 0 || y < 0 || x >= s.w || y >= s.h) return;  // surface.zig:566
  size_t start = (size_t)s.w * (size_t)y + (size_t)x;
  RGBA16 src = px_to_rgba16(px);
  if (opacity < 255) src = int_op(Z2D_OP_DST_IN, src, RGBA16{0, 0, 0, opacity});  // surface.zig:573-576
  for (size_t i = 0; i < len; i++) sfc_store(s, start + i, int_op(op, sfc_load(s, start + i), src));
}

// Surface.downsample: 4x box average with truncation, compacted in place
// (surface.zig:447-469, 687-709; pixel.zig:435-464, 633-646).  Only the mask
// types (alpha8/4/2/1) are ever down-sampled on the hot path, but RGB(A) is
// handled too.
void sfc_downsample(Sfc& s) {
  const int scale = 4;
  if (s.w < scale || s.h < scale) return;
  size_t height = (size_t)(s.h / scale), width = (size_t)(s.w / scale), worig = (size_t)s.w;
  int bits = fmt_bits(s.fmt);
  for (size_t y = 0; y < height; y++) {
    for (size_t x = 0; x < width; x++) {
      int acc[4] = {0, 0, 0, 0};
      for (int i = 0; i < scale; i++)
        for (int j = 0; j < scale; j++) {
          size_t idx = (y * scale + i) * worig + (x * scale + j);
          if (bits == 32) {
            const uint8_t* p = s.buf + idx * 4;
            for (int k = 0; k < 4; k++) acc[k] += p[k];
          } else if (bits == 8) {
            acc[0] += s.buf[idx];
          } else {
            acc[0] += packed_get(s.buf, idx, bits);
          }
        }
      size_t o = y * width + x;
      if (bits == 32) {
        uint8_t* p = s.buf + o * 4;
        for (int k = 0; k < 4; k++) p[k] = (uint8_t)(acc[k] / 16);
        if (s.fmt == Z2D_FMT_RGB || s.fmt == Z2D_FMT_XRGB) p[3] = 0;
      } else if (bits == 8) {
        s.buf[o] = (uint8_t)(acc[0] / 16);
      } else {
        packed_set(s.buf, o, bits, acc[0] / 16);
      }
    }
  }
  s.h = (int32_t)height;
  s.w = (int32_t)width;
}

}  // namespace zref
