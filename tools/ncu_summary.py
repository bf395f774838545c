"""Turn `ncu -i X.ncu-rep --page raw --csv` (stdin or a file) into the small JSON summaries kept under profiles/:
one dict per profiled launch with the metrics that the design notes quote.  usage:
  ncu -i gpurun_out/x.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/NAME.json
  python tools/ncu_summary.py --launches list.csv > profiles/NAME.md     (gpu__time_duration launch list -> table)"""
import csv
import io
import json
import sys
from collections import OrderedDict

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "launch__", "lts__t_bytes.sum", "lts__t_sector_hit_rate", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor", "sm__throughput.avg.pct", "smsp__cycles_active.avg",
        "sm__warps_active.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu", "dram__throughput.avg.pct", "lts__throughput.avg.pct", "gpc__cycles_elapsed.max",
        "sm__inst_executed_pipe_uniform", "sm__pipe_shared_cycles_active", "smsp__warp_issue_stalled")


def rows(text):
    lines = [l for l in text.splitlines() if l.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--launches":
        r = rows(open(sys.argv[2]).read())
        hdr = r[0]
        kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = OrderedDict()
        for x in r[1:]:
            if len(x) <= mv:
                continue
            try:
                v = float(x[mv].replace(",", ""))
            except ValueError:
                continue
            unit = x[hdr.index("Metric Unit")]
            us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
            a = agg.setdefault(x[kn][:80], [0, 0.0])
            a[0] += 1
            a[1] += us
        tot = sum(a[1] for a in agg.values())
        print(f"{sum(a[0] for a in agg.values())} launches, {tot:.0f} us total\n")
        print("| kernel | launches | total us | share | us/launch |\n|---|---|---|---|---|")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% | {a[1] / a[0]:.2f} |")
        return
    text = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
    r = rows(text)
    names, units = r[0], r[1]
    out = []
    for x in r[2:]:
        d = OrderedDict()
        for n, u, v in zip(names, units, x):
            if n in ("Kernel Name", "Grid Size", "Block Size") or any(n.startswith(k) for k in KEEP):
                d[f"{n} [{u}]" if n not in ("Kernel Name", "Grid Size", "Block Size") else n] = v
        out.append(d)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
