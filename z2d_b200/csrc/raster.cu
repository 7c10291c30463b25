// K4 (raster.cuh) as its own translation unit.  The two raster kernels are ~300 / ~540 KB of SASS and instruction-fetch
// sensitive, and their code depends on how NVVM partitions a unit's functions under --split-compile: built together with the
// other kernels and "--split-compile 0" they came out differently (6.4 ... 7.9 ms on config 3 for byte-identical raster source)
// whenever an unrelated kernel or flag changed.  On their own, compiled unsplit (z2d_b200/build.py), their code only changes
// when they do.
#define Z2D_RASTER_TU 1
#include "kernels.cuh"
#include "pattern.cuh"
#include "raster.cuh"

namespace z2d {

// Load the two kernels at context creation instead of at the first batch (lazy module loading would otherwise add the load of
// ~840 KB of code to the first draw call's latency).
void raster_preload() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, k_raster_tiles);
  cudaFuncGetAttributes(&a, k_raster_tiles_rich);
}

void launch_raster(const RasterArgs& A, bool rich, cudaStream_t st) {
  if (!A.n_tiles) return;
  const unsigned blocks = (unsigned)((A.n_tiles + kRasterThreads / 32 - 1) / (kRasterThreads / 32));
  if (rich) k_raster_tiles_rich<<<blocks, kRasterThreads, 0, st>>>(A);
  else k_raster_tiles<<<blocks, kRasterThreads, 0, st>>>(A);
}

}  // namespace z2d
