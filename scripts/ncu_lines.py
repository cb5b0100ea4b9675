"""Stall samples per CUDA source line of one profiled kernel (read here, no GPU needed).
Usage: python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main(path, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    fname, hdr, per = "", None, {}
    for r in rows:
        if not r:
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
            continue  # SASS rows have an empty line number; the CUDA row carries the totals of its SASS
        try:
            n, ins = float(r[si]), float(r[ii])
        except ValueError:
            continue
        key = (fname, int(r[0]))
        e = per.setdefault(key, [0.0, 0.0, r[1].strip()[:90], {}])
        e[0] += n
        e[1] += ins
        for i, k in stall:
            e[3][k] = e[3].get(k, 0.0) + float(r[i] or 0)
    tot = sum(e[0] for e in per.values())
    toti = sum(e[1] for e in per.values())
    print(f"total samples {tot:.0f}, warp instructions {toti/1e6:.1f} M")
    for (f, l), e in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        st = sorted(e[3].items(), key=lambda kv: -kv[1])[:2]
        print(f"{e[0]:7.0f} {100*e[0]/tot:5.1f}% {e[1]/1e6:7.1f}M {f}:{l:<4d} {st[0][0]}={st[0][1]:.0f} {st[1][0]}={st[1][1]:.0f} | {e[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
