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
HEADERS = ["cf_simt.h", "cf_model.h", "cf_spec_generated.h", "cf_rti_warp.h", "cf_pcond_warp.h", "cf_loop_kernels.h", "cf_kernels.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


# Generic-model libraries: the same kernel sources compiled against another generated OCP description.
# libcfnmpc_<name>.so <- csrc/cfnmpc_generic.cu with -DCF_SPEC_HEADER.  "crazyflie_generic" is the Crazyflie OCP on the
# generic path (preparation kernel + dense-stage feedback program), the cross-check of that path against the tuned one.
MODEL_LIBS = {"pendulum": "cf_spec_pendulum.h", "crazyflie_generic": "cf_spec_generated.h"}


def model_lib(name):
    return os.path.join(PKG, f"libcfnmpc_{name}.so")


def build_models(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in ["cfnmpc_generic.cu"] + HEADERS if os.path.exists(os.path.join(CSRC, f))]
    deps.append(os.path.join(PKG, "..", "include", "cfnmpc.h"))
    for name, spec in MODEL_LIBS.items():
        lib = model_lib(name)
        d = deps + [os.path.join(CSRC, spec)]
        if not force and os.path.exists(lib) and all(os.path.getmtime(x) <= os.path.getmtime(lib) for x in d):
            continue
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f'-DCF_SPEC_HEADER="{spec}"', "-I", os.path.join(PKG, "..", "include"),
                                                                              os.path.join(CSRC, "cfnmpc_generic.cu"), "-o", lib]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed building {lib}")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS if os.path.exists(os.path.join(CSRC, f))]
    deps += [os.path.join(PKG, "..", "include", f) for f in os.listdir(os.path.join(PKG, "..", "include")) if f.endswith(".h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    build_models(force, verbose)
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
