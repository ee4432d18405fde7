"""Selected raw-page metrics of an `ncu --set full` report, one block per launch.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-name-substring] > profiles/rNN_ncu_x.txt

Reads the report with `ncu -i … --page raw --csv` (works without a GPU).
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "LTS.TriageCompute.lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "LTS.TriageCompute.lts__t_sector_throughput_srcunit_tex.avg.pct_of_peak_sustained_elapsed",
    "derived__lts__lts2xbar_bytes.sum.per_second", "derived__lts__lts2xbar_bytes.sum.peak_sustained",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_op_dmma.sum",
    "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__pcsamp_sample_count",
]


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    dmma = [h for h in hdr if "dmma" in h and h not in WANT]
    kn = hdr.index("Kernel Name")
    print(f"# {rep}: ncu --set full --clock-control none (cold-cache, serialised replays); selected raw-page metrics")
    for r in rows[2:]:
        if sub and sub not in r[kn]:
            continue
        print(f"== {r[kn]}  (launch id {r[0]})")
        for w in WANT + dmma:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:82s} {r[i]:>18s} {units[i]}")
        tot = []
        for s in stalls:
            try:
                tot.append((float(r[hdr.index(s)].replace(",", "")), s))
            except ValueError:
                pass
        tot.sort(reverse=True)
        n = sum(v for v, _ in tot) or 1.0
        for v, s in tot[:8]:
            print(f"   {s:82s} {v:18.0f} ({100 * v / n:.1f} % of stall samples)")


if __name__ == "__main__":
    main()
