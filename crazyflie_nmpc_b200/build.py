"""Build the in-tree CUDA library crazyflie_nmpc_b200/libcfnmpc.so for sm_100a.

nvcc cross-compiles without a GPU; the resulting .so travels with the repo snapshot
to the GPU box (it is git-ignored, not gpurun-ignored)."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libcfnmpc.so")
SOURCES = ["cfnmpc_api.cu", "acados_shim.cpp", "cfnmpc_multi.cpp"]
HEADERS = ["cf_simt.h", "cf_model.h", "cf_spec_generated.h", "cf_rti_warp.h", "cf_pcond_warp.h", "cf_loop_kernels.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS if os.path.exists(os.path.join(CSRC, f))]
    deps += [os.path.join(PKG, "..", "include", f) for f in os.listdir(os.path.join(PKG, "..", "include")) if f.endswith(".h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, f) for f in SOURCES if os.path.exists(os.path.join(CSRC, f))]
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", os.path.join(PKG, "..", "include")] + srcs + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libcfnmpc.so")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(LIB)
