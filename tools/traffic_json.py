"""profiles/traffic.json: measured DRAM bytes per launch (ncu --set full: dram__bytes_read.sum +
dram__bytes_write.sum, mean over the captured launches) for every profile slot of bench.py.

    python tools/traffic_json.py gpurun_out/X_prof.ncu-rep profiles/traffic.json
"""
import collections, csv, json, re, subprocess, sys

SLOT_OF = [  # (regex on the kernel name incl. template args, slots it feeds)
    (r"hashgrid_fwd_pair_kernel<2>|hashgrid_fwd_taps_kernel", ["hashgrid_fwd_image"]),
    (r"hashgrid_fwd_pair_kernel<3>|hashgrid_fwd_bundle_kernel", ["hashgrid_fwd_motion"]),
    (r"hashgrid_bwd_pair_kernel<2>|hashgrid_bwd_kernel<2>|hashgrid_bwd_taps_kernel", ["hashgrid_bwd_image"]),
    (r"hashgrid_bwd_pair_kernel<3>|hashgrid_bwd_kernel<3>|hashgrid_bwd_bundle_kernel", ["hashgrid_bwd_motion"]),
    (r"mlp_fwd_tc_kernel<256", ["mlp_fwd_image"]), (r"mlp_fwd_tc_kernel<64", ["mlp_fwd_motion"]),
    (r"mlp_bwd_tc_kernel<256", ["mlp_bwd_image"]), (r"mlp_bwd_tc64_kernel", ["mlp_bwd_motion"]),
    (r"fft_rows_kernel<0>|fft_rows_kernel<false>", ["fft_rows"]), (r"fft_rows_kernel<1>|fft_rows_kernel<true>", ["fft_rows_adj"]),
    (r"motion_rows_fwd_kernel|rows_fwd_fused_kernel", ["motion_rows_fwd"]),
    (r"motion_rows_bwd_kernel|rows_bwd_fused_kernel", ["motion_rows_bwd"]),
    (r"colpass_loss_kernel", ["colpass_loss"]), (r"grad_entropy_kernel", ["grad_entropy"]),
    (r"adam_kernel", ["adam_motion", "adam_image"]),
]
N_MOTION, N_IMAGE = 14232576, 6230672      # image: MLP block + the live rows of the tap-indexed table at 320 x 320


def main(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    per = collections.defaultdict(list)

    def to_bytes(v, unit):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]

    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
        b = sum(to_bytes(r[idx[m]], units[idx[m]]) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        per[name].append(b)
    out = collections.defaultdict(float)
    for name, vals in per.items():
        mean = sum(vals) / len(vals)
        for pat, slots in SLOT_OF:
            if re.search(pat, name):
                if slots == ["adam_motion", "adam_image"]:   # one kernel, two launches per iteration
                    out["adam_motion"] = 2 * mean * N_MOTION / (N_MOTION + N_IMAGE)
                    out["adam_image"] = 2 * mean * N_IMAGE / (N_MOTION + N_IMAGE)
                else:
                    out[slots[0]] += mean                     # dense + hashed launches add up
                break
    json.dump({"source": src, "unit": "bytes per launch (dram read + write, ncu --set full, cold cache)",
               **{k: round(v) for k, v in sorted(out.items())}}, open(dst, "w"), indent=1)
    print(open(dst).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
