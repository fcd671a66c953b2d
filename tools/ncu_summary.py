"""Condenses an .ncu-rep into the JSON summaries kept under profiles/: usage python tools/ncu_summary.py rep.ncu-rep out.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "sm__inst_executed_pipe_tensor_op_dmma.sum": "dmma_instructions",
    "lts__t_bytes.sum": "l2_bytes",
}
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
out = []
for r in rows[2:]:
    e = {"kernel": r[idx["Kernel Name"]]}
    for k, nm in KEYS.items():
        if k in idx:
            v = r[idx[k]].replace(",", "")
            try:
                e[nm] = float(v)
            except ValueError:
                e[nm] = v
            if units[idx[k]] and nm in ("time_us", "dram_read", "dram_write", "l2_bytes"):
                e[nm + "_unit"] = units[idx[k]]
    stalls = {h.split("smsp__average_warps_issue_stalled_")[1].split("_per_issue_active")[0]: r[i]
              for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
    top = sorted(((float(v.replace(",", "")), k) for k, v in stalls.items() if v not in ("", "n/a")), reverse=True)[:4]
    e["top_stalls_warps_per_issue"] = {k: v for v, k in top}
    out.append(e)
json.dump({"source": sys.argv[1], "launches": out}, open(sys.argv[2], "w"), indent=1)
print(f"{len(out)} launches -> {sys.argv[2]}")
