#!/usr/bin/env python
"""Turn ncu output into the small text summaries committed under profiles/.

  python profiles/summarize_ncu.py launches gpurun_out/launches.csv > profiles/rNN_launches.md
  python profiles/summarize_ncu.py full gpurun_out/prof_x.ncu-rep   > profiles/rNN_x_full.md
  python profiles/summarize_ncu.py traffic profiles/rNN_traffic.json WORKLOAD FRAMES_PER_LAUNCH rep1.ncu-rep [rep2 ...]

`launches` takes the CSV written by  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...
`full` takes a report written by      ncu --set full --clock-control none --import-source on -o ...
`traffic` reads dram__bytes_read.sum + dram__bytes_write.sum of the first launch in each report and writes the JSON that
bench.py turns into `roofline.traffic` (so that number always comes from a capture, keyed by kernel, workload and frames
per launch -- never from a literal in the bench).
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.split("::")[-1] if "<" not in name else name


def launches(path):
    rows = []
    with open(path, newline="") as f:
        text = f.read()
    start = text.find('"ID"')
    for r in csv.DictReader(io.StringIO(text[start:])):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            if r.get("Metric Unit") == "us":
                v *= 1e3
            elif r.get("Metric Unit") == "ms":
                v *= 1e6
            rows.append((short(r["Kernel Name"]), v, r["Grid Size"], r["Block Size"]))
    agg = collections.OrderedDict()
    for k, v, g, b in rows:
        a = agg.setdefault(k, [0, 0.0, g, b])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu launch list: {len(rows)} launches, {tot / 1e6:.3f} ms of kernel time (cold cache, serialised)\n")
    print("| kernel | launches | total ms | avg us | share | grid (first) | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {a[1] / a[0] / 1e3:.1f} | {100 * a[1] / tot:.1f}% | {a[2]} | {a[3]} |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full: {path}\n")
    for r in rows[2:]:
        print(f"## `{short(r[hdr.index('Kernel Name')])}`  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for m in KEEP:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {m} | {r[i]} | {units[i]} |")
        print()


def _to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def traffic(out_path, workload, frames_per_launch, *reports):
    import json
    import os
    recs = []
    for path in reports:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, r = rows[0], rows[1], rows[2]
        rd, wr, du = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        full_name = r[hdr.index("Kernel Name")]
        m = re.match(r"(?:void\s+)?([\w:]+)", full_name)            # up to the template / argument list; namespaces dropped
        name = m.group(1).split("::")[-1] if m else full_name
        recs.append({"kernel": name, "workload": workload, "frames_per_launch": float(frames_per_launch),
                     "dram_bytes_read": _to_bytes(r[rd], units[rd]), "dram_bytes_write": _to_bytes(r[wr], units[wr]),
                     "dram_bytes_per_launch": _to_bytes(r[rd], units[rd]) + _to_bytes(r[wr], units[wr]),
                     "duration": f"{r[du]} {units[du]} (under ncu)", "source": os.path.basename(path)})
    json.dump(recs, open(out_path, "w"), indent=1)
    print(json.dumps(recs, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
