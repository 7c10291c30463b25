# variants on three shapes: config 2 (fills), config 3 (strokes), 512 config-5 scenes
for so in z2d_b200/variants/*.so; do
  echo "== $so"
  Z2D_CUDA_LIB=$PWD/$so python tools/warmup_probe.py 12 2>/dev/null | tail -1
  Z2D_CUDA_LIB=$PWD/$so python tools/stroke_timing.py 2>/dev/null | tail -1
  Z2D_CUDA_LIB=$PWD/$so python tools/c5_split.py 256 2>/dev/null | tail -1 | cut -c1-140
done
