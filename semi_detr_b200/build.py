"""Build semi_detr_b200/lib/libsemidetr_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m semi_detr_b200.build [--force] [--verbose]

The library is plain CUDA behind the C ABI of include/semidetr_b200.h -- no torch headers, so it
builds in seconds and the same .so serves ctypes (this package), cgo, JNI or a pybind shim.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsemidetr_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I" + INCLUDE]
# per-file extra flags
EXTRA = {
    # the matching cost must round like torch's separately-rounded elementwise kernels
    "hungarian.cu": ["-fmad=false"],
}
SOURCES = ["abi.cu", "msda_forward.cu", "msda_forward_tma.cu", "msda_backward.cu", "msda_backward_tile.cu", "msda_prologue.cu", "hungarian.cu", "ema.cu", "layernorm.cu", "optimizer.cu", "colsum.cu", "detr_loss.cu", "ssod.cu", "attention.cu", "exchange.cu", "gemm_tf32.cu"]
# instrumentation library (include/semidetr_b200_debug.h): microbenchmarks + the GEMM with %globaltimer stamps
DEBUG_LIB = os.path.join(LIBDIR, "libsemidetr_b200_debug.so")
DEBUG_SOURCES = [("abi.cu", []), ("umma_rate.cu", []), ("gemm_tf32.cu", ["-DSDB_GEMM_TRACE=1"])]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "semidetr_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *ARCH, *COMMON, *EXTRA.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stdout.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building semi_detr_b200 (see output above)")
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs]
    subprocess.run(cmd, check=True)
    return LIB


def build_debug(verbose=False):
    """lib/libsemidetr_b200_debug.so -- only tools/ load it; never built by __graft_entry__.build()."""
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(LIBDIR, "obj_debug")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src, extra in DEBUG_SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    subprocess.run([nvcc, *ARCH, "-shared", "-o", DEBUG_LIB, *objs], check=True)
    return DEBUG_LIB


if __name__ == "__main__":
    if "--debug" in sys.argv:
        print(build_debug(verbose="--verbose" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
