"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion uses.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_ncu.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {path}")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"\nkernel: {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:75s} {r[i]:>16s} {units[i]}")
    stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("_not_issued")]
    for r in rows[2:] if stall else []:
        vals = sorted(((float(r[hdr.index(h)] or 0), h) for h in stall), reverse=True)[:6]
        tot = sum(float(r[hdr.index(h)] or 0) for h in stall) or 1.0
        print(f"\ntop warp stall reasons (pc samples) of {r[hdr.index('Kernel Name')][:60]}:")
        for v, h in vals:
            print(f"  {h.replace('smsp__pcsamp_warps_issue_stalled_', ''):40s} {v:10.0f}  {100 * v / tot:5.1f} %")

if __name__ == "__main__":
    main(sys.argv[1])
