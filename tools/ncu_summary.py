"""Markdown summaries of ncu CSV exports for profiles/ (measurement aid, not product code).

  python tools/ncu_summary.py launches <launches.csv> [title]     per-kernel totals of a launch list
        (ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <cmd>)
  python tools/ncu_summary.py raw <raw.csv> [title]               one row per profiled launch of a full capture
        (ncu -i prof.ncu-rep --page raw --csv > raw.csv)
"""
import csv
import sys
from collections import OrderedDict


def read_rows(path):
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[start], rows[start + 1:]


def launches(path, title):
    hdr, rows = read_rows(path)
    ix = {n: i for i, n in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        key = (r[ix["Kernel Name"]], r[ix["Grid Size"]])
        ns = float(r[ix["Metric Value"]].replace(",", ""))
        if r[ix["Metric Unit"]].startswith("us"):
            ns *= 1e3
        a = agg.setdefault(key, [0.0, 0])
        a[0] += ns
        a[1] += 1
    total = sum(v[0] for v in agg.values())
    print(f"# {title}\n\nTotal device time of the {sum(v[1] for v in agg.values())} launches: {total / 1e6:.2f} ms "
          "(cold-cache, serialised: compare SHARES with the in-cycle numbers, not absolutes)\n")
    print("| share | total us | launches | avg us | grid | kernel |\n|---|---|---|---|---|---|")
    for (name, grid), (ns, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"| {100 * ns / total:.1f}% | {ns / 1e3:.0f} | {cnt} | {ns / 1e3 / cnt:.2f} | {grid} | `{name[:90]}` |")


def raw(path, title):
    hdr, rows = read_rows(path)
    ix = {n: i for i, n in enumerate(hdr)}
    cols = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs"),
            ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem)"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
            ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard")]
    cols = [(m, t) for m, t in cols if m in ix]
    units = rows[0]
    print(f"# {title}\n\n| kernel | grid | " + " | ".join(t for _, t in cols) + " |\n|---|---|" + "---|" * len(cols))
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        cells = []
        for m, _ in cols:
            v, u = r[ix[m]], units[ix[m]]
            try:
                cells.append(f"{float(v.replace(',', '')):.2f} {u}".strip())
            except ValueError:
                cells.append(v)
        print(f"| `{r[ix['Kernel Name']][:70]}` | {r[ix['Grid Size']]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else path
    (launches if mode == "launches" else raw)(path, title)
