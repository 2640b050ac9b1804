"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/X_launches.csv profiles/NAME_launches.txt
    python tools/ncu_summary.py full gpurun_out/X_prof.ncu-rep profiles/NAME_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit.startswith("n") else (v * 1e3 if unit.startswith("m") else v)
        key = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        key += " grid=" + row.get("Grid Size", "").replace(" ", "")
        agg.setdefault(key, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {src}\n")
        f.write(f"# {sum(len(v) for v in agg.values())} launches, total {tot:.1f} us\n")
        f.write(f"{'kernel':70s} {'n':>4s} {'avg_us':>9s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:70s} {len(v):4d} {sum(v) / len(v):9.1f} {100 * sum(v) / tot:6.1f}%\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on: {src} (first launch of each kernel)\n")
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
            if name in seen:
                continue
            seen.add(name)
            f.write(f"== {name}\n")
            for m in METRICS:
                if m in idx:
                    f.write(f"   {m:72s} {r[idx[m]]:>16s} {units[idx[m]]}\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
