for so in z2d_b200/variants/*.so; do
  for cell in "rgba linear none src_over integer" "rgba radial bayer src_over float" "alpha8 linear none src_over integer"; do
    echo -n "$so $cell: "; Z2D_CUDA_LIB=$PWD/$so python tools/c4_time.py $cell
  done
done
