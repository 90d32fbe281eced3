"""Build profiles/r2_traffic.json — measured DRAM traffic per batch row of the attention kernels — from `ncu --set full`
captures of the CURRENT build (run on the CPU box after the GPU session brought the .ncu-rep files back):

    python profiles/make_traffic.py gpurun_out/r2_attn_fwd.ncu-rep gpurun_out/r2_attn_bwd.ncu-rep --rows 1024 --lq 982 --lk 982

traffic = dram__bytes_read.sum + dram__bytes_write.sum of one launch; the kernels stream each (row, head) once, so it is stored
per batch row and bench.py scales it by N.  The table carries the sha256 of the kernel sources the captured library was built
from (tests/probe/gpu_ncu_r2.sh writes it next to the reports; the .so itself is not bit-reproducible across links); bench.py
refuses a table whose hash is not that of the sources in the tree."""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dram_bytes(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    res = []
    for v in vals:
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(key)
            x = float(v[i].replace(",", ""))
            tot += x * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        res.append((v[hdr.index("Kernel Name")], tot, float(v[hdr.index("gpu__time_duration.sum")].replace(",", ""))))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("fwd"); ap.add_argument("bwd")
    ap.add_argument("--rows", type=int, default=1024); ap.add_argument("--lq", type=int, default=982); ap.add_argument("--lk", type=int, default=982)
    ap.add_argument("--sha", default=os.path.join(ROOT, "gpurun_out", "r2_ncu_src.sha256"))
    ap.add_argument("--lin-bwd-ln", default=None, help="ncu --set full report of lin_tc_bwd2_kernel<LayerNorm> from tests/probe/lin_bench.py")
    ap.add_argument("--lin-fwd-ln", default=None, help="same for lin_tc_fwd_kernel<LayerNorm>")
    ap.add_argument("--tokens", type=int, default=1005568, help="token count of tests/probe/lin_bench.py")
    a = ap.parse_args()
    sha = open(a.sha).read().split()[0]
    tab = {"_source": f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none on tests/probe/attn_tc_check.py "
                      f"(N={a.rows} rows x 4 heads, Lq={a.lq}, Lk={a.lk}, p_drop 0.1); per batch row, scaled by N in bench.py",
           "src_sha256": sha, "per_row": {}, "kernels": {}}
    for name, rep in (("attn_fwd", a.fwd), ("attn_bwd", a.bwd)):
        kname, tot, dur = dram_bytes(rep)[0]
        tab["per_row"][f"{name}|{a.lq}|{a.lk}"] = tot / a.rows
        tab["kernels"][name] = {"kernel": kname, "dram_bytes": tot, "ncu_duration": dur, "report": os.path.basename(rep)}
    tab["per_token"] = {}
    for name, rep in (("lin_bwd_ln", a.lin_bwd_ln), ("lin_fwd_ln", a.lin_fwd_ln)):
        if rep:
            kname, tot, dur = dram_bytes(rep)[0]
            tab["per_token"][name] = tot / a.tokens
            tab["kernels"][name] = {"kernel": kname, "dram_bytes": tot, "ncu_duration": dur, "tokens": a.tokens, "report": os.path.basename(rep)}
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    json.dump(tab, open(path, "w"), indent=1)
    print(open(path).read())


if __name__ == "__main__":
    main()
