"""Summarise ncu outputs into small text files that are committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/r1_launches_v1.txt
    python profiles/summarize.py report   gpurun_out/prof.ncu-rep     profiles/r1_attn_v1_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = r"gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |dram__bytes_read.sum$|dram__bytes_write.sum$|" \
       r"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor|sm__warps_active.avg.pct_of_peak_sustained_active|" \
       r"launch__registers_per_thread|launch__grid_size|launch__block_size|sm__throughput.avg.pct|sm__inst_executed_pipe_(alu|fma|xu|lsu|uniform|tensor)|" \
       r"smsp__inst_executed.sum$|sm__cycles_elapsed.max|l1tex__data_bank_conflicts|smsp__warp_issue_stalled.*_per_warp_active|launch__occupancy_limit|" \
       r"sm__inst_executed_pipe_xu|smsp__issue_active.avg.pct|launch__shared_mem_per_block"


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:110]


def launches(src, dst):
    agg = collections.OrderedDict()
    with open(src) as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    total = 0.0
    for r in rows[1:]:
        ns = float(r[vi].replace(",", ""))
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1; a[1] += ns; total += ns
    with open(dst, "w") as f:
        f.write(f"# source: {src}  ({len(rows) - 1} launches, {total / 1e6:.3f} ms summed device time; cold-cache, serialised: compare SHARES)\n")
        f.write(f"{'share':>7} {'total_ms':>10} {'calls':>6} {'avg_us':>9}  kernel\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100 * ns / total:6.2f}% {ns / 1e6:10.3f} {n:6d} {ns / n / 1e3:9.1f}  {k}\n")
    print(open(dst).read()[:3000])


def report(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    pat = re.compile(KEYS)
    with open(dst, "w") as f:
        f.write(f"# source: {src}\n")
        for v in vals:
            f.write(f"\n## {v[hdr.index('Kernel Name')]}  grid={v[hdr.index('Grid Size')]} block={v[hdr.index('Block Size')]}\n")
            for h, u, x in zip(hdr, units, v):
                if pat.search(h):
                    f.write(f"{h:75s} {x:>18s} {u}\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
