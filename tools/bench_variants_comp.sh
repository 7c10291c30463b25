for so in z2d_b200/variants/*.so z2d_b200/variants/libz2d_base.so; do Z2D_CUDA_LIB=$PWD/$so python tools/bench_composite_variants.py 2>&1 | tail -1; done
