#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu --set full) into a small JSON for profiles/.
    python tools/ncu_summarize.py out.json rep1.ncu-rep [rep2 ...]"""
import csv, json, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1tex_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__inst_executed_op_shared_atom.sum": "shared_atomic_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "ctas_per_sm_limit_registers",
    "launch__occupancy_limit_shared_mem": "ctas_per_sm_limit_smem",
    "sass__inst_executed_local_loads": "local_loads",
    "sass__inst_executed_local_stores": "local_stores",
}

def main():
    out = {"source": "ncu --set full --clock-control none --import-source on (cold caches, serialised: compare shares and ratios, "
                     "not absolute times)", "kernels": []}
    for rep in sys.argv[2:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
            k = {"report": rep.split("/")[-1], "kernel": d["Kernel Name"].split("(")[0]}
            for key, name in KEYS.items():
                if key in d and d[key] != "":
                    try:
                        v = float(d[key].replace(",", ""))
                    except ValueError:
                        continue
                    k[name] = v
                    if u.get(key) and name in ("duration", "dram_read", "dram_write"):
                        k[name + "_unit"] = u[key]
            st = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(d[h]) for h in hdr
                  if "issue_stalled" in h and "per_issue_active" in h and d[h] not in ("", "0")}
            k["stall_cycles_per_issue_top"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:6])
            out["kernels"].append(k)
    json.dump(out, open(sys.argv[1], "w"), indent=1)
    print(f"{len(out['kernels'])} kernels -> {sys.argv[1]}")

if __name__ == "__main__":
    main()
