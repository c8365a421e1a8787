"""Key metrics of an .ncu-rep (one line per metric, one column per captured launch), read with `ncu -i`.
Used to turn the captures of tools/evidence.sh into the text summaries committed under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [metric-prefix ...] > profiles/x.txt
"""
import csv
import subprocess
import sys

DEFAULT = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
    "sm__inst_executed_pipe_tc", "sm__inst_executed_pipe_tma", "sm__inst_executed_pipe_tmem", "sm__inst_executed_pipe_xu",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_active.avg",
    "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit", "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled",
]


def main():
    rep = sys.argv[1]
    want = sys.argv[2:] or DEFAULT
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    idx = [i for i, h in enumerate(hdr) if any(h == w or h.startswith(w) for w in want)]
    print(f"# {rep}: ncu --set full --clock-control none (cold cache, serialised: compare shares, not absolutes)")
    for i in idx:
        if ".per_second" in hdr[i] or "Not Issued" in hdr[i] or ".max." in hdr[i] or ".min." in hdr[i]:
            continue
        vals = [r[i] for r in rows[2:]]
        if all(v in ("0", "") for v in vals):
            continue
        print(f"{hdr[i]} [{rows[1][i]}]: " + " | ".join(vals))


if __name__ == "__main__":
    main()
