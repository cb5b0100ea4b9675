"""Stall samples per CUDA source line, PER KERNEL of a report (scripts/ncu_lines.py merges all kernels).
Usage: python scripts/ncu_kernel_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main(path, top=30):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    kern, hdr, fname, per = "?", None, "", {}
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "Kernel Name":
            kern = r[1].split("(")[0][-60:] + str(sum(1 for k in per if k[0].startswith(r[1].split("(")[0][-60:])) == 0 and "" or "")
            kern = r[1][:110]
            continue
        if r[0] == "File Path":
            fname = r[1].rsplit("/", 1)[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
            stall = [(i, k[6:]) for i, k in enumerate(hdr) if k.startswith("stall_") and "Not Issued" not in k]
            continue
        if hdr is None or len(r) <= si or not r[0]:
            continue
        try:
            n, ins = float(r[si]), float(r[ii])
        except ValueError:
            continue
        e = per.setdefault((kern, fname, int(r[0])), [0.0, 0.0, r[1].strip()[:90], {}])
        e[0] += n
        e[1] += ins
        for i, k in stall:
            e[3][k] = e[3].get(k, 0.0) + float(r[i] or 0)
    for K in sorted(set(k[0] for k in per)):
        sub = {k: v for k, v in per.items() if k[0] == K}
        tot = sum(v[0] for v in sub.values())
        agg = {}
        for v in sub.values():
            for k, x in v[3].items():
                agg[k] = agg.get(k, 0) + x
        print(f"===== {K}\n total samples {tot:.0f}, warp instructions {sum(v[1] for v in sub.values())/1e6:.1f} M")
        print("  " + ", ".join(f"{k}={100*x/tot:.1f}%" for k, x in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
        for (_, f, l), e in sorted(sub.items(), key=lambda kv: -kv[1][0])[:top]:
            st = sorted(e[3].items(), key=lambda kv: -kv[1])[:2]
            print(f"{e[0]:7.0f} {100*e[0]/tot:5.1f}% {e[1]/1e6:7.2f}M {f}:{l:<4d} {st[0][0]}={st[0][1]:.0f} {st[1][0]}={st[1][1]:.0f} | {e[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
