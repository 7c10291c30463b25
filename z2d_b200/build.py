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
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"]
# Every translation unit is compiled on its own, UNSPLIT (--split-compile 1).  The split count decides how NVVM partitions a
# unit's functions before optimising, and with it the generated code: "0" (= the build machine's core count, shared between the
# units of one nvcc command) made the K4 kernels come out differently -- and up to 14 % slower -- whenever an unrelated file or
# flag changed, and with a fixed count > 1 two builds of the same source still differed in kernels.cu (the partition depends on
# thread timing).  Unsplit builds are reproducible bit for bit (checked: SASS of two builds identical) and were the fastest K4
# measured (config 3 raster 6.40 ms against 6.88 ms with a 4-way split, config 2 2.29 against 2.34 ms); kernels.cu takes 3.5
# minutes this way.  Z2D_*_SPLIT override the counts for experiments.
SPLIT = {"kernels.cu": os.environ.get("Z2D_KERNELS_SPLIT", "1"), "raster.cu": os.environ.get("Z2D_RASTER_SPLIT", "1"), "z2d_lib.cu": "1"}
BUILD_DIR = os.path.join(HERE, "_build")


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, extra=(), so=None):
    so = so or SO
    if not force and not extra and so == SO and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(BUILD_DIR, exist_ok=True)
    tag = os.path.splitext(os.path.basename(so))[0]
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(BUILD_DIR, f"{tag}.{os.path.splitext(src)[0]}.o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["--split-compile", SPLIT[src]] + list(extra) + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    failed = False
    for src, pr in procs:
        out, err = pr.communicate()
        if pr.returncode != 0:
            sys.stderr.write(out + err)
            failed = True
        elif verbose:
            sys.stderr.write(err)
    if failed:
        raise RuntimeError("nvcc failed building libz2d_cuda.so")
    res = subprocess.run([nvcc, "-shared", "-cudart", "static", "-o", so] + objs, capture_output=True, text=True)
    for obj in objs:  # (nothing is incremental here, and the objects would travel with every gpurun snapshot)
        if os.path.exists(obj):
            os.remove(obj)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libz2d_cuda.so")
    return so


if __name__ == "__main__":
    # python -m z2d_b200.build [--force] [-v] [--variant NAME -DFOO=1 ...]  (variants: z2d_b200/variants/NAME.so, tools/build_variant.sh)
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        os.makedirs(os.path.join(HERE, "variants"), exist_ok=True)
        print(build(force=True, extra=sys.argv[i + 2:], so=os.path.join(HERE, "variants", sys.argv[i + 1] + ".so")))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
