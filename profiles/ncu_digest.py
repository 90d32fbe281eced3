"""Digest of an `ncu --set full` report (run on the CPU box): `python profiles/ncu_digest.py <file.ncu-rep> [--src N]`.
Prints the handful of raw metrics the roofline discussion in DESIGN.md uses, the warp-stall breakdown, and
(with --src) the N source lines with the most stall samples."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor", "sm__inst_executed_pipe_xu.avg.pct", "sm__pipe_xu_cycles_active.avg.pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct", "sm__inst_executed_pipe_uniform",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
    "smsp__average_warp", "smsp__average_warps_issue_stalled",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    path = sys.argv[1]
    hdr, units, launches = raw(path)
    for vals in launches:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"=== {name}")
        stalls = []
        for h, u, v in zip(hdr, units, vals):
            if "issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                try:
                    stalls.append((float(v.replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_warp_active.pct", "")))
                except ValueError:
                    pass
                continue
            if any(h.startswith(k.strip()) for k in KEYS) and "issue_stalled" not in h and ".max" not in h and ".min" not in h:
                if "per_second" in h or "peak_sustained_elapsed" in h and "throughput" not in h:
                    continue
                print(f"  {h:86s} {v:>16s} {u}")
        for pct, nm in sorted(stalls, reverse=True)[:8]:
            print(f"  stall {nm:40s} {pct:8.1f} %")
    if "--src" in sys.argv:
        n = int(sys.argv[sys.argv.index("--src") + 1])
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        while rows and rows[0] and rows[0][0] == "Kernel Name":
            rows = rows[1:]
        h = rows[0]
        try:
            si = h.index("# Samples") if "# Samples" in h else [i for i, x in enumerate(h) if "Sampl" in x][0]
        except IndexError:
            print("no sampling column:", h[:12]); return
        srci = h.index("Source") if "Source" in h else 1
        body = []
        for r in rows[1:]:
            try:
                body.append((int(r[si].replace(",", "")), r[srci][:150]))
            except (ValueError, IndexError):
                pass
        tot = sum(b[0] for b in body) or 1
        top = sorted(range(len(body)), key=lambda i: -body[i][0])[:n]
        for i in sorted(top):
            cnt, src = body[i]
            print(f"  {100.0 * cnt / tot:5.1f}%  [{i:5d}] {src}")


if __name__ == "__main__":
    main()
