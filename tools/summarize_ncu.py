"""Turn ncu CSV exports into the short tables committed under profiles/.

  python tools/summarize_ncu.py launches <launches.csv>          # --metrics gpu__time_duration.sum launch list
  python tools/summarize_ncu.py full <raw.csv>                   # ncu -i x.ncu-rep --page raw --csv
"""
import csv
import sys
from collections import OrderedDict


def short(name):
    name = name.replace("void ", "")
    for pre in ("ifb::", "(anonymous namespace)::"):
        name = name.replace(pre, "")
    return name.split("(")[0]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        k = short(r[ki])
        if k.startswith(("synth_", "set_scalars")) or "memset" in k.lower():
            continue     # model set-up, not part of a decode step
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1000.0
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.1f}% |")
    print(f"| all | {sum(a[0] for a in agg.values())} | {tot:.1f} | | 100% |")


FULL = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
]


def full(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for n, r in enumerate(rows[2:]):
        print(f"### launch {n}: `{short(r[idx['Kernel Name']])}`  grid {r[idx['launch__grid_size']]} x block {r[idx['launch__block_size']]}\n")
        print("| metric | value |")
        print("|---|---:|")
        for key, label in FULL:
            if key in idx:
                print(f"| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
