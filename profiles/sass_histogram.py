"""Per-kernel SASS opcode histogram of the built library (evidence that the hot kernels are tcgen05 / TMEM / TMA code).

    python profiles/sass_histogram.py [out.txt]

Runs `cuobjdump -sass` on vaesne-dev_b200/lib/libvaesne_b200.so (no GPU needed) and counts, per kernel, the opcodes that
matter: UTCHMMA / UTCQMMA / UTCMMA-family (tcgen05.mma), LDTM / STTM (tcgen05.ld / .st), UTMALDG / UTMASTG (TMA),
UTCBAR / SYNCS (mbarrier), MUFU.EX2, FFMA2 / FADD2 / FMUL2 (packed fp32), RED / ATOM, LDS / STS, LDG / STG, spills (LDL / STL)."""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vaesne-dev_b200", "lib", "libvaesne_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "MUFU.EX2", "MUFU.RCP", "MUFU.LG2",
         "FFMA2", "FADD2", "FMUL2", "F2FP", "RED", "ATOM", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR.SYNC", "ELECT"]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w in ("MUFU.EX2", "MUFU.RCP", "MUFU.LG2", "BAR.SYNC") and op.startswith(w)):
                kernels[cur][w] += 1
                break
    try:
        names = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True, check=True).stdout.splitlines()
    except Exception:      # noqa: BLE001
        names = list(kernels)
    sha = hashlib.sha256(open(LIB, "rb").read()).hexdigest()[:16]
    lines = [f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (sha256 {sha}); cuobjdump -sass, sm_100a", ""]
    tot = collections.Counter()
    for (k, c), nm in zip(kernels.items(), names):
        nm = re.sub(r"\(.*", "", nm)
        hits = "  ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        lines.append(f"{nm:60s} instrs={c['_total']:6d}  {hits}")
        tot.update(c)
    lines += ["", "TOTAL  " + "  ".join(f"{w}={tot[w]}" for w in WATCH if tot[w])]
    text = "\n".join(lines) + "\n"
    if out_path:
        open(out_path, "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
