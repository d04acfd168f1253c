#!/usr/bin/env python
"""Summarise .ncu-rep captures (gpurun_out/) into the small CSVs kept under profiles/:
    python tools/ncu_summary.py gpurun_out/ncu_r2_c3_assign_dmma_resident.ncu-rep profiles/ncu_r2_assign_dmma_summary.csv "comment"
Keeps the metrics the roofline discussion uses (durations, DRAM bytes, pipe utilisation, stall reasons, occupancy,
launch geometry); the first line is a comment naming the capture."""
import csv, io, subprocess, sys

KEEP = ("gpu__time_duration", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "sm__pipe_tensor", "sm__pipe_fp64",
        "sm__pipe_alu", "sm__pipe_fma", "sm__inst_executed_pipe", "sm__throughput", "sm__warps_active", "smsp__average_warp", "smsp__warp_issue_stalled",
        "smsp__average_warps_issue_stalled", "l1tex__data_bank_conflicts", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate", "lts__t_bytes",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem", "launch__occupancy", "launch__waves",
        "smsp__inst_executed.sum", "smsp__issue_active", "sm__cycles_elapsed.max", "smsp__pcsamp_warps_issue_stalled")


def main():
    rep, out, comment = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, vals = rows[0], rows[1], rows[2:]
    kname = vals[0][header.index("Kernel Name")] if "Kernel Name" in header else ""
    with open(out, "w") as f:
        f.write("# %s | kernel: %s | one launch per row set, ncu --set full --clock-control none\n" % (comment, kname))
        for li, v in enumerate(vals):
            for h, u, x in zip(header, units, v):
                if any(h.startswith(k) or ("." + k) in h for k in KEEP):
                    f.write("%s%s,%s,%s\n" % (("launch%d:" % li) if len(vals) > 1 else "", h, u, x))
    print("wrote", out, "kernel", kname)


if __name__ == "__main__":
    main()
