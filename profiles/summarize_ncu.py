"""Turns gpurun_out/*.csv / *.ncu-rep captures into the small text summaries kept under profiles/.
  python profiles/summarize_ncu.py launches gpurun_out/launches_r01b.csv 146 > profiles/r01_launches.md
  python profiles/summarize_ncu.py full gpurun_out/prof_r01_textgemm.ncu-rep > profiles/r01_textgemm_full.md
"""
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def launches(path, per_step):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    step = rows[1:][-per_step:]
    tot = sum(float(r[vi]) for r in step) / 1e3
    print(f"# ncu launch list, last step ({per_step} launches), gpu__time_duration (cold-cache, serialised)\n")
    print(f"total {tot:.1f} us\n")
    agg = {}
    for r in step:
        n = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi]) / 1e3
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {n} | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
    print("\n| # | kernel | grid | block | us |\n|---|---|---|---|---|")
    for i, r in enumerate(step):
        print(f"| {i} | {r[ki].split('(')[0].replace('void ', '')} | {r[gi]} | {r[bi]} | {float(r[vi]) / 1e3:.1f} |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of {path}\n")
    for r in rows[2:]:
        print(f"## {r[hdr.index('Kernel Name')]}  (launch id {r[0]})\n")
        for i, h in enumerate(hdr):
            if h in KEEP:
                print(f"- {h} [{units[i]}] = {r[i]}")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]))
    else:
        full(sys.argv[2])
