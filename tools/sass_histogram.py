"""Per-kernel SASS opcode histogram of libimmoco_b200.so (cuobjdump -sass): the evidence behind the
tcgen05 / TMEM / packed-fp32 / reduction claims of DESIGN.md.   python tools/sass_histogram.py > profiles/round2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "miccai24_immoco_b200", "libimmoco_b200.so")
# mnemonic prefixes worth counting (B200_PROFILING.md): tensor core + TMEM, TMA, reductions / atomics, packed fp32
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UTMALDG", "UTMASTG", "REDG", "RED", "ATOMG", "ATOM",
         "FFMA2", "FMUL2", "FADD2", "HMMA", "ELECT", "SYNCS", "LDGSTS", "MUFU", "DFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            kernels[name]["_total"] += 1
            for w in WATCH:
                if op.startswith(w):
                    kernels[name][w + ("" if op == w else "")] += 1
                    break
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a), one line per kernel")
    print("# columns: total instructions, then the watched mnemonic prefixes that occur\n")
    tot = collections.Counter()
    for (mangled, c), nice in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*$", "", nice.replace("(anonymous namespace)::", "")).replace("void ", "")
        if short.startswith("cub::"):
            short = "cub::" + re.sub(r"<.*", "", short.split("::")[-1])
        parts = [f"{k}={v}" for k, v in c.items() if k != "_total"]
        print(f"{short:64s} total={c['_total']:6d}  " + "  ".join(parts))
        tot.update(c)
    print("\n# whole library: " + "  ".join(f"{k}={v}" for k, v in tot.items()))
    print("# no UTMALDG / UTMASTG: the library does not use TMA -- the MLP kernels stage their E tiles with LDG -> registers ->"
          " STS because every element is split into its tf32 hi / lo parts on the way (DESIGN.md 4.1)")


if __name__ == "__main__":
    sys.exit(main())
