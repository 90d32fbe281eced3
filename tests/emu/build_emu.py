"""Build the CUDA-semantics emulator flavour of the kernel sources (TEST INFRASTRUCTURE)."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "vaesne-dev_b200", "csrc")
OUT = os.path.join(HERE, "libvaesne_emu.so")
# tcgen05 / TMA kernels (attn_tc*.cu) are Blackwell-only and are not part of the emulated build
PORTABLE = ["api.cu", "lin.cu", "attn.cu", "attn_small.cu", "attn_mid.cu", "misc.cu", "loss.cu", "extra.cu"]


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in PORTABLE] + [os.path.join(HERE, "emu_cuda.cpp")]
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "emu_cuda.h"), os.path.join(ROOT, "include", "vaesne_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DVAESNE_EMU", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-I" + HERE]
    for s in srcs:
        cmd += ["-x", "c++", s]
    cmd += ["-o", OUT, "-lpthread"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
