# bench every variant library under z2d_b200/variants (tuning experiments); device-resident timing only, 20 replays each
for so in z2d_b200/variants/*.so z2d_b200/variants/*.so; do
  Z2D_CUDA_LIB=$PWD/$so python tools/warmup_probe.py 30 2>/dev/null | awk -v so="$so" '{r+=$6; t+=$8; n++} END {printf "%s raster %.3f total %.3f (mean of %d)\n", so, r/n, t/n, n}'
done
