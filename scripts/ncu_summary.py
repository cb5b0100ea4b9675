"""Summarise an .ncu-rep (read here, no GPU needed): one block per profiled launch with the
metrics the roofline uses.  Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/x.md]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor memory (TMEM) active % (elapsed)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "dram bytes read"),
    ("dram__bytes_write.sum", "dram bytes written"),
    ("dram__bytes_read.sum.per_second", "dram read rate"),
    ("dram__bytes_write.sum.per_second", "dram write rate"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__block_size", "block size"),
    ("launch__grid_size", "grid size"),
    ("launch__cluster_size", "cluster size"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for h, i in list(col.items()):  # some metrics carry a "TPC.TriageCompute." style prefix
        col.setdefault(h.split(".", 2)[-1] if h.startswith(("TPC.", "SM_", "LTS.")) else h, i)
    print(f"# ncu summary of `{path}`\n")
    print("Per profiled launch (`ncu --set full --clock-control none`, cold caches, serialised replays —")
    print("compare shares and ratios, not absolute times, with the CUDA-event numbers in bench.py).\n")
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        print(f"## {name[:150]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for key, label in KEYS:
            if key in col:
                print(f"| {label} (`{key}`) | {r[col[key]]} | {units[col[key]]} |")
        rd = r[col["dram__bytes_read.sum"]] if "dram__bytes_read.sum" in col else None
        print()


if __name__ == "__main__":
    main(sys.argv[1])
