# bench every variant library under z2d_b200/variants (tuning experiments)
for so in z2d_b200/libz2d_cuda.so z2d_b200/variants/*.so; do
  Z2D_CUDA_LIB=$PWD/$so python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > /tmp/v.json
  python - "$so" <<'PY'
import json, sys
b = json.load(open("/tmp/v.json"))
print(sys.argv[1], "ms_per_step", round(b["ms_per_step"], 3), "raster", round(b["stages_ms"]["raster"], 3), "e2e", round(b["e2e"]["ms_per_step"], 2))
PY
done
