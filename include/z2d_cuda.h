/*
 * z2d_cuda.h -- C ABI of the B200-native fill/stroke rasterise-and-composite
 * library (libz2d_cuda.so).
 *
 * This is the drop-in boundary for the one hot path of vancluever/z2d that this
 * repository re-implements: everything below z2d's unmanaged painter /
 * compositor entry points.  Each entry point names the reference interface it
 * replaces (file:line relative to the z2d tree).  All types are plain C PODs;
 * enum values follow the declaration order of the reference's Zig enums so a
 * Zig shim can pass `@intFromEnum(x)` straight through.
 *
 * Calls are stream-ordered and asynchronous unless stated otherwise; draw
 * calls are recorded into a per-context command batch and executed, in
 * submission order, at the next flush point (z2d_flush, z2d_sync,
 * z2d_surface_download, z2d_composite on the same context, or when the batch
 * is full).  Results are identical to executing every call immediately.
 *
 * Threading: a z2d_ctx is single-threaded (as z2d's Context is); distinct
 * contexts may be used from distinct threads.
 */
#ifndef Z2D_CUDA_H
#define Z2D_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (reference: painter.zig:54-63,204-209; surface.zig:85-91;
 *      Transformation.zig:26-28; internal/InternalError.zig) ---------------- */
enum {
  Z2D_OK = 0,
  Z2D_E_PATH_NOT_CLOSED = -1,               /* FillError.PathNotClosed */
  Z2D_E_PIXEL_SOURCE_NOT_PREMULTIPLIED = -2, /* PixelSourceNotPreMultiplied */
  Z2D_E_INVALID_WIDTH = -3,                 /* Surface.Error.InvalidWidth */
  Z2D_E_INVALID_HEIGHT = -4,                /* Surface.Error.InvalidHeight */
  Z2D_E_INVALID_STATE = -5,                 /* InternalError.InvalidState */
  Z2D_E_OUT_OF_MEMORY = -6,                 /* mem.Allocator.Error */
  Z2D_E_INVALID_MATRIX = -7,                /* Transformation.Error.InvalidMatrix */
  Z2D_E_DEVICE = -8,                        /* CUDA / NCCL failure (new) */
  Z2D_E_INVALID_ARG = -9                    /* NULL handle, bad enum (new) */
};

/* pixel.Format (pixel.zig:47-56) */
enum {
  Z2D_FMT_ARGB = 0, Z2D_FMT_XRGB = 1, Z2D_FMT_RGB = 2, Z2D_FMT_RGBA = 3,
  Z2D_FMT_ALPHA8 = 4, Z2D_FMT_ALPHA4 = 5, Z2D_FMT_ALPHA2 = 6, Z2D_FMT_ALPHA1 = 7
};

/* compositor.Operator (compositor.zig:46-155) */
enum {
  Z2D_OP_CLEAR = 0, Z2D_OP_SRC, Z2D_OP_DST, Z2D_OP_SRC_OVER, Z2D_OP_DST_OVER,
  Z2D_OP_SRC_IN, Z2D_OP_DST_IN, Z2D_OP_SRC_OUT, Z2D_OP_DST_OUT, Z2D_OP_SRC_ATOP,
  Z2D_OP_DST_ATOP, Z2D_OP_XOR, Z2D_OP_PLUS, Z2D_OP_MULTIPLY, Z2D_OP_SCREEN,
  Z2D_OP_OVERLAY, Z2D_OP_DARKEN, Z2D_OP_LIGHTEN, Z2D_OP_COLOR_DODGE,
  Z2D_OP_COLOR_BURN, Z2D_OP_HARD_LIGHT, Z2D_OP_SOFT_LIGHT, Z2D_OP_DIFFERENCE,
  Z2D_OP_EXCLUSION, Z2D_OP_HUE, Z2D_OP_SATURATION, Z2D_OP_COLOR,
  Z2D_OP_LUMINOSITY, Z2D_OP_COUNT
};

/* compositor.Precision (compositor.zig:214-217) */
enum { Z2D_PRECISION_INTEGER = 0, Z2D_PRECISION_FLOAT = 1 };

/* options.zig:23-91 */
enum { Z2D_FILL_NON_ZERO = 0, Z2D_FILL_EVEN_ODD = 1 };
enum { Z2D_JOIN_MITER = 0, Z2D_JOIN_ROUND = 1, Z2D_JOIN_BEVEL = 2 };
enum { Z2D_CAP_BUTT = 0, Z2D_CAP_ROUND = 1, Z2D_CAP_SQUARE = 2 };
enum { Z2D_AA_NONE = 0, Z2D_AA_DEFAULT = 1, Z2D_AA_MULTISAMPLE_4X = 2, Z2D_AA_SUPERSAMPLE_4X = 3 };

/* internal/path_nodes.zig:9-21 -- PathNodeTag.  Points are DEVICE space
 * (the CTM was already applied by Path.zig:124-145). */
enum { Z2D_NODE_MOVE_TO = 0, Z2D_NODE_LINE_TO = 1, Z2D_NODE_CURVE_TO = 2, Z2D_NODE_CLOSE_PATH = 3 };

typedef struct z2d_node {
  uint32_t tag;
  uint32_t _pad;
  double p[6]; /* move/line: p[0..1]; curve: p1=(p[0],p[1]) p2=(p[2],p[3]) p3=(p[4],p[5]) */
} z2d_node;   /* 56 bytes */

/* pixel.Pixel (pixel.zig:100-140): channel values as stored by the format
 * (alpha4: a in 0..15, alpha2: 0..3, alpha1: 0..1). */
typedef struct z2d_pixel {
  uint32_t format;
  uint8_t r, g, b, a;
} z2d_pixel;

/* color.Color (color.zig:40-70): de-multiplied colour in one of three spaces. */
enum { Z2D_COLOR_LINEAR_RGB = 0, Z2D_COLOR_SRGB = 1, Z2D_COLOR_HSL = 2 };
typedef struct z2d_color {
  uint32_t space;
  float c[4]; /* r,g,b,a or h,s,l,a -- already clamped as Color.init does */
} z2d_color;

/* gradient.Stop (gradient.zig:776-781); the list is sorted as Stop.List keeps
 * it (offset ascending, ties by insertion index, gradient.zig:797-811). */
typedef struct z2d_stop {
  float offset;
  z2d_color color;
} z2d_stop;

/* gradient.GradientType; color.InterpolationMethod (+ Polar) */
enum { Z2D_GRADIENT_LINEAR = 0, Z2D_GRADIENT_RADIAL = 1, Z2D_GRADIENT_CONIC = 2 };
enum { Z2D_INTERP_LINEAR_RGB = 0, Z2D_INTERP_SRGB = 1, Z2D_INTERP_HSL = 2 };
enum { Z2D_POLAR_SHORTER = 0, Z2D_POLAR_LONGER = 1, Z2D_POLAR_INCREASING = 2, Z2D_POLAR_DECREASING = 3 };

/* gradient.Gradient (gradient.zig:31-160).
 *   linear: geom = {x0,y0,x1,y1}
 *   radial: geom = {inner_x,inner_y,inner_r,outer_x,outer_y,outer_r}
 *   conic : geom = {x,y,angle}
 * inv_ctm is the *stored* transformation of the gradient, i.e. the inverse of
 * the CTM passed to setTransformation (gradient.zig:201-203), laid out as
 * {ax,by,cx,dy,tx,ty}; identity when no transformation was set. */
typedef struct z2d_gradient {
  uint32_t type;
  uint32_t method;
  uint32_t polar;
  uint32_t n_stops;
  double geom[6];
  double inv_ctm[6];
  const z2d_stop* stops;
} z2d_gradient;

/* Dither (Dither.zig:28-58) */
enum { Z2D_DITHER_NONE = 0, Z2D_DITHER_BAYER = 1, Z2D_DITHER_BLUE_NOISE = 2 };
enum { Z2D_DITHER_SRC_PIXEL = 0, Z2D_DITHER_SRC_COLOR = 1, Z2D_DITHER_SRC_GRADIENT = 2 };

/* pattern.Pattern (pattern.zig:32-44) */
enum { Z2D_PATTERN_OPAQUE = 0, Z2D_PATTERN_GRADIENT = 1, Z2D_PATTERN_DITHER = 2 };
typedef struct z2d_pattern {
  uint32_t kind;
  z2d_pixel pixel;              /* OPAQUE; DITHER with SRC_PIXEL */
  const z2d_gradient* gradient; /* GRADIENT; DITHER with SRC_GRADIENT */
  uint32_t dither_type;         /* DITHER */
  uint32_t dither_source;
  uint32_t dither_scale;        /* Dither.scale (u4), Context.zig:690-695 */
  z2d_color dither_color;       /* DITHER with SRC_COLOR */
} z2d_pattern;

/* painter.FillOptions (painter.zig:28-48) */
typedef struct z2d_fill_opts {
  uint32_t anti_aliasing_mode;
  uint32_t fill_rule;
  uint32_t op;
  uint32_t precision;
  double tolerance;
} z2d_fill_opts;

/* painter.StrokeOptions (painter.zig:145-198) */
typedef struct z2d_stroke_opts {
  uint32_t anti_aliasing_mode;
  uint32_t line_cap_mode;
  uint32_t line_join_mode;
  uint32_t op;
  uint32_t precision;
  uint32_t hairline;
  double line_width;
  double miter_limit;
  double tolerance;
  double dash_offset;
  const double* dashes;
  size_t n_dashes;
  double ctm[6]; /* {ax,by,cx,dy,tx,ty} */
} z2d_stroke_opts;

typedef struct z2d_ctx z2d_ctx; /* device + stream + command batch */
typedef struct z2d_sfc z2d_sfc; /* device-resident surface */

/* compositor.SurfaceCompositor.Operation.Param (compositor.zig:232-283) */
enum { Z2D_PARAM_NONE = 0, Z2D_PARAM_DITHER = 1, Z2D_PARAM_GRADIENT = 2, Z2D_PARAM_PIXEL = 3, Z2D_PARAM_SURFACE = 4 };
typedef struct z2d_comp_param {
  uint32_t kind;
  z2d_pattern pattern;   /* PIXEL / GRADIENT / DITHER expressed as a pattern */
  const void* surface;   /* SURFACE: a z2d_sfc* (device library) */
} z2d_comp_param;

/* compositor.SurfaceCompositor.Operation (compositor.zig:220-230) */
typedef struct z2d_comp_op {
  uint32_t op;
  z2d_comp_param dst;
  z2d_comp_param src;
} z2d_comp_op;

/* ------------------------------------------------------------------------- */

/* Library / context ------------------------------------------------------- */
int32_t z2d_version(void);                      /* ABI version, currently 1 */
const char* z2d_last_error(const z2d_ctx* ctx); /* text of the last Z2D_E_DEVICE */

/* One context per device/thread.  `stream` is a cudaStream_t or NULL for a
 * private non-blocking stream.  (No reference equivalent: z2d has no device.) */
int32_t z2d_ctx_create(int32_t device, void* stream, z2d_ctx** out);
void z2d_ctx_destroy(z2d_ctx* ctx);
/* Recording is pipelined: every `max_draws` recorded draws the batch is handed to a worker thread that uploads
 * and executes it while the caller keeps recording (results are unchanged: batches run in order on one stream).
 * 0 disables the hand-over (one batch per flush point).  Default 32768.  Flushes first. */
int32_t z2d_ctx_set_chunk(z2d_ctx* ctx, uint32_t max_draws);
int32_t z2d_flush(z2d_ctx* ctx); /* enqueue everything recorded so far */
int32_t z2d_sync(z2d_ctx* ctx);  /* flush + wait */

/* Surface (surface.zig:97-186 init/initPixel/initBuffer, 188 deinit).
 * Byte layout of upload/download == the reference's `buf` slices: tightly
 * packed w*h pixels, 4 B for argb/xrgb/rgb/rgba, 1 B alpha8, and
 * bit-contiguous LSB-first alpha4/2/1 with rows NOT byte aligned
 * (surface.zig:632,756-765).  initial_px may be NULL (zeroed). */
int32_t z2d_surface_create(z2d_ctx* ctx, uint32_t format, int32_t width, int32_t height,
                           const z2d_pixel* initial_px, z2d_sfc** out);
/* Band of a larger canvas (one very large canvas split into horizontal bands, one per GPU).  The surface stores
 * rows [band_y0, band_y0 + band_rows) of a canvas `canvas_height` rows high; draw calls, put_pixel and pattern coordinates are
 * given in CANVAS space and only the rows held here are touched, so replaying the same calls on every band and stacking the
 * bands gives exactly the full-canvas result.  band_y0 must be a multiple of 16 (the tile height).  z2d_surface_height,
 * byte_len, upload and download refer to the band's own rows.  z2d_composite on a band: offsets 0 and no surface params. */
int32_t z2d_surface_create_band(z2d_ctx* ctx, uint32_t format, int32_t width, int32_t canvas_height, int32_t band_y0,
                                int32_t band_rows, const z2d_pixel* initial_px, z2d_sfc** out);
int32_t z2d_surface_band(const z2d_sfc* sfc, int32_t* band_y0, int32_t* canvas_height);
/* Band VIEWS: the band's rows live inside a full canvas instead of a buffer of their own, so a finished band IS its part of
 * the canvas and the "gather bands to one rank" step of SURVEY 8e disappears into the raster kernel's tile write-back.
 *   z2d_surface_band_view       band over a canvas of the same context (rank 0's own band);
 *   z2d_surface_ipc_export      64-byte CUDA IPC handle of a canvas, to be sent to the other ranks of the node (any host channel);
 *   z2d_surface_open_peer_band  band over a canvas that lives on ANOTHER process / GPU of the node: tile loads and the one
 *                               write-back per touched tile go over NVLink (peer mapping) straight into that canvas.
 * The owner of the canvas reads it after the writers have synchronised (z2d_sync on every rank, then a host barrier).  For the
 * packed formats a band must start on a 16-byte boundary of the canvas.  The views do not own the pixels: destroy them before
 * the canvas.  (No reference counterpart: z2d is single-process; the seam is Surface.buf being a slice of a larger buffer.) */
int32_t z2d_surface_band_view(z2d_sfc* canvas, int32_t band_y0, int32_t band_rows, z2d_sfc** out);
int32_t z2d_surface_ipc_export(z2d_sfc* canvas, void* handle64);
int32_t z2d_surface_open_peer_band(z2d_ctx* ctx, const void* handle64, uint32_t format, int32_t width, int32_t canvas_height,
                                   int32_t band_y0, int32_t band_rows, z2d_sfc** out);
void z2d_surface_destroy(z2d_sfc* sfc);
size_t z2d_surface_byte_len(const z2d_sfc* sfc);
int32_t z2d_surface_width(const z2d_sfc* sfc);
int32_t z2d_surface_height(const z2d_sfc* sfc);
uint32_t z2d_surface_format(const z2d_sfc* sfc);
int32_t z2d_surface_upload(z2d_sfc* sfc, const void* host, size_t n);
int32_t z2d_surface_download(z2d_sfc* sfc, void* host, size_t n); /* flushes + syncs */
/* The same read-back without blocking the caller: everything recorded so far is enqueued, the copy into `host` (pinned memory
 * for a truly asynchronous copy) runs on a separate stream after it and overlaps the calls recorded afterwards.  `host` is
 * valid after the next z2d_sync.  The reference has no counterpart (its buffers ARE host memory, surface.zig:373): this is the
 * boundary's replacement for reading `Surface.buf` of many surfaces (a batch of scenes) without one stall per surface. */
int32_t z2d_surface_download_async(z2d_sfc* sfc, void* host, size_t n);
void* z2d_surface_device_ptr(z2d_sfc* sfc); /* raw device pointer (interop) */

/* Replaces the pixel transform of export_png.writePNGIDATStream / encodeRGBAVec (src/export_png.zig:150-373): the surface as the
 * scanline bytes a PNG holds before zlib, produced on the device so the read-back is export-ready.  argb/rgba: R,G,B,A with the
 * colour channels de-multiplied in integer space (src/internal/pixel_vector.zig:27-49); xrgb/rgb: R,G,B; alpha8: one grey byte;
 * alpha4/2/1: grey samples packed most-significant-first, each row padded to a byte ((w*bits+7)/8 bytes).  Z2D_EXPORT_SRGB applies
 * WriteToPNGFileOptions.color_profile = .srgb (round(255*pow(c/255, 1/2.2)) per colour channel; ignored for alpha formats),
 * Z2D_EXPORT_FILTER_BYTE puts the filter-type byte 0 in front of every row.  n must equal z2d_surface_export_size().  A band
 * surface exports the rows it holds.  Flushes + syncs like z2d_surface_download. */
#define Z2D_EXPORT_SRGB 1u
#define Z2D_EXPORT_FILTER_BYTE 2u
size_t z2d_surface_export_size(const z2d_sfc* sfc, uint32_t flags);
int32_t z2d_surface_export(z2d_sfc* sfc, uint32_t flags, void* host, size_t n);
/* Surface.paintPixel (surface.zig:295) */
int32_t z2d_surface_paint_pixel(z2d_sfc* sfc, const z2d_pixel* px);
/* Surface.putPixel (surface.zig:288): out-of-bounds coordinates are ignored */
/* Surface.downsample (src/surface.zig:447-490, 687-709): 4x4 box average with truncation of every channel, in the surface's own
 * format; width and height become width/4, height/4 (surfaces smaller than 4 pixels in either direction are left alone).  Whole
 * surfaces only (not bands / views).  Flushes + syncs (the pixel buffer is replaced). */
int32_t z2d_surface_downsample(z2d_sfc* sfc);
int32_t z2d_surface_put_pixel(z2d_sfc* sfc, int32_t x, int32_t y, const z2d_pixel* px);
/* Surface.getPixel (surface.zig:280): the pixel in the surface's own format (channel values as stored).  Returns 1 and leaves
 * *out untouched where the reference returns null (coordinates outside the surface, or outside the rows a band holds).
 * Flushes + syncs (a 4-byte read-back). */
int32_t z2d_surface_get_pixel(z2d_sfc* sfc, int32_t x, int32_t y, z2d_pixel* out);

/* painter.fill (painter.zig:66-143).  Node coordinates must be finite: with a NaN or infinite control point the reference's
 * Spline.decompose recurses without bound, and so, in effect, does the subdivision here (known limits). */
int32_t z2d_fill(z2d_ctx* ctx, z2d_sfc* sfc, const z2d_pattern* pattern,
                 const z2d_node* nodes, size_t n_nodes, const z2d_fill_opts* opts);

/* painter.stroke (painter.zig:214-344) */
int32_t z2d_stroke(z2d_ctx* ctx, z2d_sfc* sfc, const z2d_pattern* pattern,
                   const z2d_node* nodes, size_t n_nodes, const z2d_stroke_opts* opts);

/* compositor.SurfaceCompositor.run (compositor.zig:302-440).  Infallible in
 * the reference (invalid combinations are silent no-ops); the status only
 * reports device/argument errors.  Z2D_E_INVALID_ARG also for a surface
 * parameter that does not cover the composited rectangle (the clip of
 * compositor.zig:347-374 only looks at ops[0].src; the reference would read
 * past the smaller surface's strides) and for the destination used as a
 * parameter of itself at a non-zero offset (there the reference's result
 * depends on its scanline / vector order). */
int32_t z2d_composite(z2d_ctx* ctx, z2d_sfc* dst, int32_t dst_x, int32_t dst_y,
                      const z2d_comp_op* ops, size_t n_ops, uint32_t precision);

/* Batched, ordered submission: equivalent to calling z2d_fill / z2d_stroke once
 * per element, in order (the loop a caller of painter.fill/stroke would run;
 * BASELINE configs 2, 3 and 5 issue 10^5 calls).  statuses may be NULL; the
 * return value is the first non-OK status (later commands are still recorded). */
typedef struct z2d_draw_cmd {
  uint32_t kind;                 /* 0 = fill, 1 = stroke */
  uint32_t _pad;
  z2d_sfc* surface;
  const z2d_pattern* pattern;
  const z2d_node* nodes;
  size_t n_nodes;
  const z2d_fill_opts* fill;     /* kind 0 */
  const z2d_stroke_opts* stroke; /* kind 1 */
} z2d_draw_cmd;
int32_t z2d_submit(z2d_ctx* ctx, const z2d_draw_cmd* cmds, size_t n, int32_t* statuses);

/* ---- text runs from device-resident glyph outlines (text.show, src/text.zig:73-195).
 * text.show builds ONE path for a run: per glyph it sets the path transformation to translate(x + advance + pp1, y).scale(s, s)
 * (text.zig:165-172) and replays the glyph's outline through Path.moveTo / lineTo / curveTo / close (Glyph.Outline.appendToPath,
 * src/internal/Glyph.zig:845-869), i.e. every outline point goes through Transformation.userToDevice
 * (src/Transformation.zig:194-206), and fills the result.  Here the outline is uploaded ONCE per glyph and the per-point
 * transformation runs on the device with the same two multiplies and two adds in the same order, so a run costs 56 bytes per
 * GLYPH of host work and upload instead of 56 bytes per NODE:
 *   z2d_glyph_cache_add   `nodes` = what appendToPath produces under the identity transformation: closed sub-paths, each
 *                         close_path followed by the move_to Path.close leaves behind (src/Path.zig:453-476); coordinates
 *                         already clamped as Path does.  Errors as painter.fill (PathNotClosed, InvalidState).
 *   z2d_fill_glyphs       one painter.fill whose node list is the concatenation, in order, of the cached outlines
 *                         transformed by m = {ax, by, cx, dy, tx, ty}; options and errors as z2d_fill. */
typedef struct z2d_glyph_instance {
  uint32_t glyph; /* id returned by z2d_glyph_cache_add */
  uint32_t _pad;
  double m[6];
} z2d_glyph_instance; /* 56 bytes */
int32_t z2d_glyph_cache_add(z2d_ctx* ctx, const z2d_node* nodes, size_t n_nodes, uint32_t* glyph_out);
int32_t z2d_fill_glyphs(z2d_ctx* ctx, z2d_sfc* sfc, const z2d_pattern* pattern, const z2d_glyph_instance* glyphs, size_t n_glyphs,
                        const z2d_fill_opts* opts);

/* Re-executes the device pipeline of the most recently flushed batch from its
 * device-resident inputs (nodes, draw table); nothing is read from the host.
 * Benchmarking aid: separates kernel time from host recording and H2D copies. */
int32_t z2d_replay(z2d_ctx* ctx);

/* Statistics of the last executed batch (counters and device timings the
 * benchmark reports; times are CUDA-event milliseconds on the context stream). */
typedef struct z2d_stats {
  uint64_t draws;         /* fill/stroke calls executed */
  uint64_t nodes;         /* path nodes uploaded */
  uint64_t edges;         /* flattened polygon edges (32 B each) */
  uint64_t band_edges;    /* edge references after tile-row binning */
  uint64_t tile_items;    /* entries of the per-tile-row ordered draw lists */
  uint64_t tiles;         /* tiles the raster kernel was launched over */
  uint64_t covered_px;    /* sum over draws of pixels with coverage > 0 (composited pixels) */
  uint64_t region_px;     /* sum over draws of the evaluated bounding-region pixels */
  uint64_t kernel_launches;
  uint64_t h2d_bytes;     /* bytes uploaded for the batch */
  uint64_t tile_pairs;    /* (draw, 16x16 tile) pairs the raster kernel evaluated coverage for */
  uint64_t crossings;     /* (edge, sub-scanline) crossings evaluated exactly in f64 (Polygon.zig:305) */
  float ms_flatten, ms_bin, ms_lists, ms_raster, ms_total;
  float _pad;
} z2d_stats;
int32_t z2d_get_stats(const z2d_ctx* ctx, z2d_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* Z2D_CUDA_H */
