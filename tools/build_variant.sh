# build a tuning variant of the library: tools/build_variant.sh NAME -DZ2D_RASTER_MIN_CTAS=3 ...  -> z2d_b200/variants/NAME.so
name=$1; shift
mkdir -p z2d_b200/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false --expt-relaxed-constexpr \
  -Xcompiler -fPIC -shared -cudart static --split-compile 0 -t 0 "$@" -o z2d_b200/variants/$name.so z2d_b200/csrc/kernels.cu z2d_b200/csrc/raster.cu z2d_b200/csrc/z2d_lib.cu
