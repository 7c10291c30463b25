"""Build libz2d_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libz2d_cuda.so")
SOURCES = ["kernels.cu", "raster.cu", "z2d_lib.cu"]
HEADERS = ["kernels.cuh", "z2d_batch.cuh", "z2d_device.cuh", "raster.cuh", "pattern.cuh", "composite.cuh", "stroke.cuh", "stroke_units.cuh", "geom.cuh", "smallbatch.cuh", "slowpath.cuh", "blue_noise_table.h",
           "../../include/z2d_cuda.h"]
# -fmad=false: the reference never fuses a*b+c and its results (f64 edge crossings,
# f32 blend arithmetic) are reproduced bit-exactly only without contraction.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "--split-compile", "0", "-t", "0"]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libz2d_cuda.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
