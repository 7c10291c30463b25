// K4 (raster.cuh) as its own translation unit.  The two raster kernels are ~300 / ~530 KB of SASS and instruction-fetch
// sensitive: compiled together with the other kernels, ptxas scheduled them differently whenever an unrelated kernel changed
// (6.85 vs 7.87 ms on config 3 for byte-identical raster source).  On their own, their code only changes when they do.
#define Z2D_RASTER_TU 1
#include "kernels.cuh"
#include "pattern.cuh"
#include "raster.cuh"

namespace z2d {

// Load the two kernels before anything else of the library.  With lazy module loading a kernel's code is placed when it is
// first used, i.e. behind whatever was launched before it, and K4's speed was seen to depend on that placement (6.85 vs 7.8 ms on
// config 3 for the same SASS, flipping whenever an unrelated kernel changed size).  Loaded first, at context creation, its
// placement no longer depends on the other kernels.
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
