"""Stage the UNMODIFIED reference package for the benchmark's reference arm.

    python baseline/install_ref.py

Runs, from a writable copy under /tmp (the reference tree is read-only and setuptools writes build/ next to setup.py),

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference/package>

`baseline/_ref/` is git-ignored (the reference's sources never enter this repository's history) but not gpurun-ignored, so the
installed package travels to the GPU box, where /root/reference does not exist.  Nothing under vaesne-dev_b200/ imports it:
only `bench.py --impl reference` (through baseline/ref_arm.py, in its own interpreter) does.  Outcome on this image: the
reference is pure Python (setup.py: name VAESNe, find_packages, no install_requires), the wheel builds and installs offline;
`--no-deps` is a no-op.  Its one missing import, matplotlib (training_util.py:6, plot_util.py), is satisfied by the inert stub
in baseline/stubs/ — plotting is never called on the measured path."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VAESNE_REFERENCE", "/root/reference/package")
TARGET = os.path.join(HERE, "_ref")


def installed() -> bool:
    return os.path.exists(os.path.join(TARGET, "VAESNe", "mmVAE.py"))


def install(force: bool = False) -> bool:
    """True if baseline/_ref holds the reference afterwards."""
    if installed() and not force:
        return True
    if not os.path.exists(os.path.join(REF, "setup.py")):
        return False
    tmp = tempfile.mkdtemp(prefix="vaesne_ref_")
    try:
        src = os.path.join(tmp, "package")
        shutil.copytree(REF, src)
        if os.path.exists(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
               "--target", TARGET, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-2000:] + r.stderr[-2000:])
            return False
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return installed()


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print(f"[baseline] reference {'installed in' if ok else 'NOT available for'} {TARGET}")
    sys.exit(0 if ok else 1)
