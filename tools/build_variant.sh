# build a tuning variant of the library: tools/build_variant.sh NAME -DZ2D_RASTER_MIN_CTAS=3 ...  -> z2d_b200/variants/NAME.so
# (same per-unit compile commands as the product build, z2d_b200/build.py)
name=$1; shift
python -m z2d_b200.build --variant "$name" "$@"
