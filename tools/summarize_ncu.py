"""Summaries of ncu CSV launch lists for profiles/ (run here, no GPU needed).

    python tools/summarize_ncu.py launches gpurun_out/launches_X.csv  "header text"  > profiles/launches_X_summary.txt
    python tools/summarize_ncu.py last     gpurun_out/launches_rerank_X.csv "header text" [skip_regex]
    ncu -i X.ncu-rep --page raw --csv > raw.csv;  python tools/summarize_ncu.py raw raw.csv "header text"
`launches`: per-kernel call count, total device time and share.  `last`: the kernels of the last
iteration in launch order with time and DRAM bytes (needs dram__bytes_read/write.sum in the CSV).
`raw`: the metrics DESIGN.md quotes from a `--set full` capture, one block per captured launch."""
import csv
import re
import sys
from collections import OrderedDict, defaultdict


def rows(path):
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("(anonymous namespace)::", "<unnamed>::")
    return name[:70]


RAW_METRICS = [
    "Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.sum",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
]


def raw(path, header):
    with open(path, newline="") as f:
        table = list(csv.reader(l for l in f if l.startswith('"')))
    hdr, units = table[0], table[1]
    print("# " + header)
    for vals in table[2:]:
        print("----")
        for m in RAW_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("%s = %s %s" % (m, vals[i], units[i]))


def main():
    mode, path, header = sys.argv[1], sys.argv[2], sys.argv[3]
    if mode == "raw":
        return raw(path, header)
    rs = rows(path)
    print("# " + header)
    if mode == "launches":
        tot, cnt = defaultdict(float), defaultdict(int)
        for r in rs:
            if r["Metric Name"] == "gpu__time_duration.sum":
                v = float(r["Metric Value"].replace(",", ""))
                v = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"], 1e-6)
                tot[short(r["Kernel Name"])] += v
                cnt[short(r["Kernel Name"])] += 1
        total = sum(tot.values())
        print("# (launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes)")
        print("%-72s %6s %12s %7s" % ("kernel", "calls", "total_ms", "share"))
        for k in sorted(tot, key=lambda k: -tot[k]):
            print("%-72s %6d %12.3f %6.2f%%" % (k, cnt[k], tot[k], 100 * tot[k] / total))
        print("total %.3f ms over %d launches" % (total, sum(cnt.values())))
    else:
        per = OrderedDict()
        for r in rs:
            d = per.setdefault(int(r["ID"]), {"name": short(r["Kernel Name"])})
            v = float(r["Metric Value"].replace(",", ""))
            if r["Metric Name"] == "gpu__time_duration.sum":
                d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r["Metric Unit"], 1e-6)
            else:
                d["bytes"] = d.get("bytes", 0.0) + v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
        ids = list(per)
        # last iteration = launches after the last occurrence of the first kernel of the sequence
        first_name = sys.argv[4] if len(sys.argv) > 4 else "prep_rows"
        starts = [i for i in ids if first_name in per[i]["name"]]
        lo = starts[-1] if starts else ids[0]
        print("%-64s %9s %10s %9s" % ("kernel", "ms", "dramMB", "GB/s"))
        t = 0.0
        for i in ids:
            if i < lo:
                continue
            d = per[i]
            ms, b = d.get("ms", 0.0), d.get("bytes", 0.0)
            t += ms
            print("%-64s %9.4f %10.2f %9.1f" % (d["name"][:64], ms, b * 1e-6, b / ms * 1e-6 if ms else 0))
        print("total %.3f ms" % t)


if __name__ == "__main__":
    main()
