# stroke workload (BASELINE config 3 shape) on every variant library under z2d_b200/variants, then on the in-tree build
for so in z2d_b200/variants/*.so z2d_b200/libz2d_cuda.so; do
  echo "$so $(Z2D_CUDA_LIB=$PWD/$so python tools/stroke_timing.py 2>/dev/null | head -1 | grep -o "'ms_flatten': [0-9.]*, .*")"
done
