"""Turn ncu outputs into the small text summaries committed under profiles/.

  python scripts/summarize_ncu.py full   gpurun_out/prof.ncu-rep        > profiles/rNN_ncu_full.md
  python scripts/summarize_ncu.py launch gpurun_out/launches.csv [skip] > profiles/rNN_launches.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("smsp__inst_executed.sum", "warp inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, n) for m, n in METRICS if m in idx]
    print("| kernel | " + " | ".join(f"{n} ({units[idx[m]]})" if units[idx[m]] else n for m, n in cols) + " |")
    print("|---|" + "---|" * len(cols))
    seen = collections.Counter()
    for d in data:
        name = re.sub(r"\(.*", "", d[idx["Kernel Name"]]).replace("void ", "").replace("tnf::<unnamed>::", "").replace("unnamed>::", "")
        seen[name] += 1
        if seen[name] > 1:
            continue
        vals = []
        for m, _ in cols:
            v = d[idx[m]]
            try:
                f = float(v)
                vals.append(f"{f:,.1f}" if abs(f) < 1e6 else f"{f:,.0f}")
            except ValueError:
                vals.append(v)
        print(f"| {name} | " + " | ".join(vals) + " |")


def launch(path, skip=0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))[skip:]
    agg = {}
    tot = 0.0
    for r in rows:
        name = r["Kernel Name"].replace("tnf::<unnamed>::", "tnf::").replace("void ", "")
        m = re.search(r"(tnf::\w+)", name)
        if m:
            name = m.group(1)
        else:
            f = re.search(r"(\w+Functor\w*|\w+_kernel_cuda\w*|CatArrayBatchedCopy\w*|reduce_kernel|distribution_elementwise\w*|RadixSort\w+)", name)
            name = "torch:" + (f.group(1) if f else re.sub(r"[<(].*", "", name)[-50:])
        v = float(r["Metric Value"])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"{len(rows)} launches, {tot / 1e6:.3f} ms of kernel time (ncu: serialised, cold cache -- compare SHARES)\n")
    print("| kernel | launches | total us | share |")
    print("|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| {n} | {c} | {t / 1e3:.1f} | {100 * t / tot:.1f}% |")


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2])
    else:
        launch(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
