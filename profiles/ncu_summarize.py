"""Condense `ncu --set full` captures (.ncu-rep) into the JSON summary committed under profiles/.

    python profiles/ncu_summarize.py OUT.json LABEL=path.ncu-rep [LABEL=path.ncu-rep ...]

Runs `ncu -i <rep> --page raw --csv` (no GPU needed) and keeps the metrics the roofline discussion in
DESIGN.md / profiles/README.md cites; bench.py reads dram__bytes_* from the entry whose label starts with
"rollout" and contains "precision <p>"."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg", "smsp__cycles_active.max",
    "smsp__cycles_active.min", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def summarize(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit = rows[0], rows[1]
    out = []
    for val in rows[2:]:
        d = {"kernel": val[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h):
                try:
                    d[h] = {"value": float(val[i]), "unit": unit[i]}
                except ValueError:
                    pass
        out.append(d)
    return out[0] if len(out) == 1 else out


def main():
    out_path, res = sys.argv[1], {}
    for spec in sys.argv[2:]:
        label, path = spec.split("=", 1)
        res[label] = summarize(path)
    json.dump(res, open(out_path, "w"), indent=1, sort_keys=True)
    print("wrote", out_path, list(res))


if __name__ == "__main__":
    main()
