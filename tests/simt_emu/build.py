"""Build tests/simt_emu/libcfemu.so: the CUDA warp program compiled for the host with its 32
lanes run as lock-step fibers (TEST HARNESS ONLY, see emu.cpp)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "crazyflie_nmpc_b200", "csrc")
LIB = os.path.join(HERE, "libcfemu.so")


def build(force=False):
    """libcfemu.so (Crazyflie OCP) and libcfemu_pendulum.so (the same sources with the pendulum's generated description)."""
    for lib, spec in ((LIB, "cf_spec_generated.h"), (os.path.join(HERE, "libcfemu_pendulum.so"), "cf_spec_pendulum.h")):
        deps = [os.path.join(HERE, "emu.cpp")] + [os.path.join(CSRC, f) for f in ("cf_simt.h", "cf_model.h", "cf_rti_warp.h", "cf_pcond_warp.h", spec)]
        if not force and os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps):
            continue
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", CSRC, f'-DCF_SPEC_HEADER="{spec}"',
                        os.path.join(HERE, "emu.cpp"), "-o", lib, "-lpthread"], check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
