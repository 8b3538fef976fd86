"""Compile the reference's UNMODIFIED ROS node source (crazyflie_controller/src/acados_mpc.cpp, read from where it lies
under /root/reference -- never copied) into two stand-alone programs, with the stand-in ROS / boost / Eigen headers of
tests/dropin/stubs/ and the driver tests/dropin/node_driver.cpp:

  tests/dropin/_build/node_ref    linked against the reference's own acados/HPIPM/BLASFEO build (oracle/_ref/libcfref.so)
                                  through tests/dropin/refglue/ -> mints tests/golden/node_loop_golden.npz
  tests/dropin/_build/node_ours   compiled against include/ (the drop-in headers) and linked against libcfnmpc.so: the
                                  demonstration that the node builds unchanged on this library; runs on the GPU box

Both binaries are git-ignored build products that travel to the GPU box.  Needs /root/reference; a no-op where it is absent
(the prebuilt binaries are used)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CF_REFERENCE", "/root/reference")
NODE = os.path.join(REF, "crazyflie_controller", "src", "acados_mpc.cpp")
OUT = os.path.join(HERE, "_build")
AC = os.path.join(REF, "acados")


def available():
    return os.path.exists(NODE)


def build(verbose=False):
    if not available():
        return False
    os.makedirs(OUT, exist_ok=True)
    common = ["-std=c++17", "-O2", "-w", "-I", os.path.join(HERE, "stubs"), f'-DCF_NODE_SOURCE="{NODE}"']
    drv = os.path.join(HERE, "node_driver.cpp")
    # -- our backend: exactly the include directory INTEGRATION.md tells the node's CMakeLists to use
    pkg = os.path.join(ROOT, "crazyflie_nmpc_b200")
    cmd = ["g++"] + common + ["-I", os.path.join(ROOT, "include"), drv, "-o", os.path.join(OUT, "node_ours"),
                              "-L", pkg, "-lcfnmpc", "-Wl,-rpath,$ORIGIN/../../../crazyflie_nmpc_b200"]
    subprocess.run(cmd, check=True)
    # -- reference backend: the reference's own headers + the glue that stands in for the generated solver
    ref_so = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(ref_so, "libcfref.so")):
        inc = [os.path.join(HERE, "refglue"), AC, os.path.join(AC, "interfaces"), os.path.join(AC, "external"),
               os.path.join(AC, "external", "hpipm", "include"), os.path.join(ref_so, "include"),
               os.path.join(AC, "external", "blasfeo", "include")]
        iflags = [f for d in inc for f in ("-I", d)]
        glue_o = os.path.join(OUT, "ref_node_glue.o")
        subprocess.run(["gcc", "-std=c99", "-O2", "-w", "-c", os.path.join(HERE, "refglue", "ref_node_glue.c"), "-o", glue_o] + iflags,
                       check=True)
        # The node reads nlp_out->total_time (acados_mpc.cpp:616), a field the 2020 acados had and the vendored acados
        # dropped (acados/acados/ocp_nlp/ocp_nlp_common.h:228-242); the value only feeds a message that is compiled out
        # (PUB_OPENLOOP_TRAJ 0).  Renamed at compile time for this translation unit only -- the source stays unmodified.
        cmd = ["g++"] + common + ["-Dtotal_time=inf_norm_res"] + iflags + [drv, glue_o, "-o", os.path.join(OUT, "node_ref"), "-L", ref_so, "-lcfref",
                                           "-Wl,-rpath,$ORIGIN/../../../oracle/_ref", "-lm", "-lpthread"]
        subprocess.run(cmd, check=True)
    if verbose:
        print("built", os.listdir(OUT))
    return True


if __name__ == "__main__":
    ok = build(verbose=True)
    if not ok:
        print("reference tree absent: nothing built")
    sys.exit(0)
