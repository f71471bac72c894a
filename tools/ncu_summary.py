"""Summarise an ncu report (raw + source pages) for profiles/: key metrics and per-role stall samples.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_fp64.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("== kernel:", name[:100])
        for h, u, v in zip(hdr, units, row):
            if h in KEYS:
                print("  %-75s %-12s %s" % (h, u, v))
    src = page(rep, "source")
    if len(src) < 3:
        return
    hdr = src[1]
    data = src[2:]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = [r for r in data if len(r) > max(isrc, isamp, iex) and r[isamp].isdigit()]
    total = sum(int(r[isamp]) for r in data)
    print("== source page: %d SASS instructions, %d stall samples" % (len(data), total))
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top]
    for i in sorted(idx):
        r = data[i]
        print("  [%5d] samples=%8s (%4.1f%%) executed=%10s  %s" % (i, r[isamp], 100.0 * int(r[isamp]) / max(1, total),
                                                                  r[iex], r[isrc].strip()[:90]))


if __name__ == "__main__":
    main()
