"""SASS opcode histogram per kernel of libjegal_b200.so (cuobjdump, no GPU needed): the evidence that the contraction
kernels are tcgen05 / TMEM / TMA code and what the HBM-bound kernels are made of.
Usage: python scripts/sass_histogram.py [jegal_b200/libjegal_b200.so] > profiles/sass_r02.md"""
import collections
import re
import subprocess
import sys

KEY = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "UTCATOMSWS", "SYNCS", "FHFMA", "HADD2", "FFMA", "FMNMX", "FMNMX3", "FADD", "FMUL",
       "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "RED", "ATOMG", "ATOMS", "BAR", "BRA", "CALL"]


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    kern, hist, order = None, {}, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = kern.replace("(anonymous namespace)::", "").replace("void ", "")
            kern = re.sub(r"\((?!int\)|bool\)).*", "", kern).replace("jegal::", "").replace("(int)", "").replace("(bool)", "") or m.group(1)[-40:]
            hist[kern] = collections.Counter()
            order.append(kern)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and kern:
            op, mods = m.group(1), m.group(2)
            hist[kern][op] += 1
            if op in ("UTCHMMA", "UTMALDG", "LDTM", "STG", "LDG", "LDS") and mods:
                hist[kern][op + mods] += 1
    print("# SASS opcode histograms (cuobjdump -sass jegal_b200/libjegal_b200.so, sm_100a)\n")
    print("`UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `LDTM` = tcgen05.ld, `UTMALDG` = cp.async.bulk.tensor (TMA), `UTCBAR` = tcgen05.commit,")
    print("`SYNCS` = mbarrier ops, `FHFMA` = mixed-precision fma.rn.f32.f16 (row norms fused into the load), `STG.E.EF.256` = 256-bit streaming store.\n")
    print("| kernel | instr | " + " | ".join(KEY) + " | notable |")
    print("|---|---|" + "---|" * (len(KEY) + 1))
    for k in order:
        h = hist[k]
        tot = sum(v for op, v in h.items() if "." not in op)
        notable = ", ".join(f"{op} x{v}" for op, v in sorted(h.items()) if "." in op and any(t in op for t in ("2CTA", "256", "x32", "x16", "128")))[:160]
        print(f"| `{k[:70]}` | {tot} | " + " | ".join(str(h.get(op, 0) or "") for op in KEY) + f" | {notable} |")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "jegal_b200/libjegal_b200.so")
