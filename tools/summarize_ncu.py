"""Summarise ncu outputs into small text files under profiles/ (the .ncu-rep files stay in gpurun_out/, scratch).

    python tools/summarize_ncu.py launches <launches.csv> <out.txt> "<title>"
    python tools/summarize_ncu.py full <report.ncu-rep> <out.txt> "<title>"
"""
import csv
import subprocess
import sys
from collections import OrderedDict


def launches(path, out, title):
    rows = list(csv.reader(open(path, errors="replace")))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[start]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        us = v / 1e3 if r[mu] in ("ns", "nsecond") else (v if r[mu].startswith("us") else v * 1e3)
        name = r[kn].split("(")[0].replace("void ", "").replace("pb::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n# total kernel time {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches "
                f"(ncu: cold-cache, serialised -- compare SHARES, not absolutes)\n")
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{us:12.1f} us {n:5d} launches {100 * us / tot:6.2f}%  {name}\n")


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def full(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    with open(out, "w") as f:
        f.write(f"# {title}\n# source: ncu --set full --clock-control none --import-source on ({rep.split('/')[-1]})\n")
        for r in rows[2:]:
            f.write(f"kernel: {r[h.index('Kernel Name')]}\n")
            for w in WANT:
                if w in h:
                    f.write(f"  {w:66s} {r[h.index(w)]} {u[h.index(w)]}\n")
        hi = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
        if hi:
            hh = srows[hi[0]]
            body = srows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(srows))]
            si, sc = hh.index("# Samples"), hh.index("Source")
            stalls = [i for i, c in enumerate(hh) if c.startswith("stall_") and "Not Issued" not in c]
            tot = sum(int(r[si]) for r in body if r[si].isdigit()) or 1
            agg = {}
            for r in body:
                for i in stalls:
                    if r[i].isdigit():
                        agg[hh[i]] = agg.get(hh[i], 0) + int(r[i])
            f.write("warp-state samples by stall reason: " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]) + "\n")
            f.write("hottest SASS instructions (samples, instruction):\n")
            for r in sorted(body, key=lambda r: -int(r[si]) if r[si].isdigit() else 0)[:12]:
                f.write(f"  {r[si]:>6s}  {r[sc].strip()[:90]}\n")
            ops = {}
            for r in body:
                op = [o for o in r[sc].strip().split() if not o.startswith("@")]
                if op:
                    ops[op[0].split(".")[0]] = ops.get(op[0].split(".")[0], 0) + 1
            keys = [k for k in ops if k.startswith(("UTC", "LDTM", "UBLKCP", "UTMA", "LDGSTS", "SYNCS"))]
            if keys:
                f.write("Blackwell-specific SASS present: " + ", ".join(f"{k} x{ops[k]}" for k in sorted(keys)) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
