"""Turn an `ncu --set full` report into the small csv summaries kept in this directory (and the stack-traffic sidecar).

    python profiles/summarize_ncu.py gpurun_out/stack.ncu-rep profiles/r2_v3_stack_ncu_summary.csv \
        --header "ncu --set full ... (what was captured)" [--traffic-sidecar profiles/stack_traffic.json]

Reads the report with `ncu -i <rep> --page raw --csv`; one column per captured launch. Nothing here runs on the GPU.
"""
import argparse
import csv
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

METRICS = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__time_duration.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
    "launch__block_size", "launch__grid_size", "launch__cluster_dim_x", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]

_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out_csv")
    ap.add_argument("--header", required=True)
    ap.add_argument("--traffic-sidecar")
    ap.add_argument("--kernel-label", default="")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    names, units, launches = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(names)}
    with open(args.out_csv, "w") as f:
        f.write(f"# {args.header}\n")
        f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(launches))) + "\n")
        for m in METRICS:
            if m in col:
                f.write(",".join([m, units[col[m]]] + [r[col[m]].replace(",", "") for r in launches]) + "\n")
    if args.traffic_sidecar:
        import bench                                         # the build id is bench.py's own definition
        def bytes_of(metric):
            return float(launches[0][col[metric]].replace(",", "")) * _SCALE[units[col[metric]]]
        side = {"build_id": bench.build_id(), "dram_bytes_read": bytes_of("dram__bytes_read.sum"),
                "dram_bytes_write": bytes_of("dram__bytes_write.sum"), "kernel": args.kernel_label,
                "source": f"{args.out_csv} (ncu --set full, launch 0)"}
        with open(args.traffic_sidecar, "w") as f:
            json.dump(side, f, indent=1, sort_keys=True)
    print(f"wrote {args.out_csv}: {len(launches)} launch(es)")


if __name__ == "__main__":
    main()
