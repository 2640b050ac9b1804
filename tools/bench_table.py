"""Markdown table of the bench lines kept under profiles/round2_bench/ (one JSON line per file, named
<config>[_m<n_M>]_n<gpus>.json).   python tools/bench_table.py"""
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = {}
for f in glob.glob(os.path.join(ROOT, "profiles", "round2_bench", "*.json")):
    name = os.path.basename(f)[:-5]
    m = re.match(r"(c\d)(?:_m(\d))?_n(\d)(_det)?$", name)
    if not m:
        continue
    d = json.loads(open(f).read().strip().splitlines()[-1])
    cfg, nm, n, det = m.group(1), m.group(2), int(m.group(3)), m.group(4)
    key = cfg + (f" n_M={nm}" if nm else "") + (" deterministic" if det else "")
    rows.setdefault(key, {})[n] = d
print("| configuration | GPUs | slices/s (HBM-resident) | slices/s end to end (host buffers) | ms per step | steps x slices | ms per slice-iteration |")
print("|---|---|---|---|---|---|---|")
for key in sorted(rows):
    base = rows[key].get(1)
    for n in sorted(rows[key]):
        d = rows[key][n]
        per = d["config"]["slices_per_step"]
        scale = f" ({d['value'] / base['value']:.2f} x)" if base and n > 1 else ""
        print(f"| {key} | {n} | {d['value']:.3f}{scale} | {d['e2e']['value']:.3f} | {d['ms_per_step']:.1f} | {d['steps']} x {per} | "
              f"{d.get('ms_per_iter', float('nan')):.4f} |")
