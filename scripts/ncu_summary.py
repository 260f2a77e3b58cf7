#!/usr/bin/env python3
"""ncu_summary.py <report.ncu-rep> <out.json> [label]: compact, committable summary of an
`ncu --set full` capture (per kernel: duration, DRAM bytes, throughputs, occupancy limits,
warp-stall sampling)."""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
label = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.avg.per_second",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_requests_srcunit_tex.sum",
        "smsp__inst_executed.sum"]
res = []
for d in data:
    k = {"kernel": d[hdr.index("Kernel Name")][:160]}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            try:
                k[w] = [float(d[i]), units[i]]
            except ValueError:
                pass
    stalls = {}
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h:
            try:
                v = float(d[i])
            except ValueError:
                continue
            if v > 0:
                stalls[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = v
    tot = sum(stalls.values()) or 1
    k["warp_stall_share_pct"] = {a: round(100 * b / tot, 1)
                                 for a, b in sorted(stalls.items(), key=lambda x: -x[1])[:8]}
    rd = k.get("dram__bytes_read.sum", [0, ""])
    wr = k.get("dram__bytes_write.sum", [0, ""])
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    k["dram_bytes_per_launch"] = rd[0] * mult.get(rd[1], 1) + wr[0] * mult.get(wr[1], 1)
    res.append(k)
json.dump({"label": label, "source": rep.split("/")[-1], "command": "ncu --set full --clock-control none "
           "--import-source on (see scripts/gpu_round.sh / gpu_quick.sh)", "kernels": res},
          open(out, "w"), indent=1)
print("wrote", out, len(res), "kernel instance(s)")
