"""Build lib/libvaesne_b200.so: every .cu under csrc/ compiled by nvcc for sm_100a (cross-compiles
without a GPU) and linked into ONE C-ABI shared library kept in-tree."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OUT = os.path.join(LIBDIR, "libvaesne_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


if os.environ.get("VAESNE_TC_PROFILE"):      # probe build: per-phase clocks in attn_tc_dkv_kernel (tests/probe/attn_tc_check.py TC_PROF=1)
    FLAGS.append("-DVAESNE_TC_PROFILE")
if os.environ.get("B2_CHECK"):               # probe build: lin_tc_bwd2_kernel checks that a ring stage holds the expected tile
    FLAGS.append("-DB2_CHECK")                # (tests/probe/bwd2_acc_probe.py MARK=1)


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def source_hash():
    """sha256 over the kernel sources (sorted csrc/*.cu, *.cuh and the C-ABI header): identifies a build independently of
    link-time noise in the .so; profiles/r2_traffic.json is tied to it."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(ROOT, "include", "vaesne_b200.h")]:
        h.update(os.path.basename(f).encode()); h.update(open(f, "rb").read())
    return h.hexdigest()


def up_to_date():
    if not os.path.exists(OUT):
        return False
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(ROOT, "include", "vaesne_b200.h")]
    return all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs + ["-lcudart"])
    return OUT


if __name__ == "__main__":
    if "--source-hash" in sys.argv:
        print(source_hash()); sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
