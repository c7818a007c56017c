"""Summarise an `ncu --page raw --csv` export: one JSON record per captured launch with the metrics the
roofline discussion uses.  usage: python tools/ncu_summary.py <raw.csv> [<out.json>]"""
import csv
import json
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "registers",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active": "dmma_pipe_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_cycle",
}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    with open(sys.argv[1]) as f:
        r = csv.reader(f)
        hdr = next(r); units = next(r)
        out = []
        for row in r:
            d = {"kernel": row[hdr.index("Kernel Name")].split("(")[0]}
            for k, name in KEYS.items():
                if k in hdr:
                    i = hdr.index(k)
                    try:
                        v = float(row[i].replace(",", ""))
                    except ValueError:
                        continue
                    u = units[i]
                    if u in SCALE:
                        v *= SCALE[u]
                        name_u = name + ("_us" if u in ("ns", "us", "ms", "s") else "_bytes")
                    else:
                        name_u = name
                    d[name_u] = v
            out.append(d)
    txt = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")
    for d in out:
        print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in d.items()})


if __name__ == "__main__":
    main()
