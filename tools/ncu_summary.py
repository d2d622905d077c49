"""Text summary of an .ncu-rep (key metrics, stall reasons, top stalled SASS lines) for profiles/."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def raw(rep):
    """rows of `ncu --page raw --csv`: from an .ncu-rep, or from a .csv exported on the GPU box (the reports of a
    --set full capture are too large to bring back through gpurun_out/)"""
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    rows = raw(rep)
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print("  %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        print("  stall reasons (warps stalled per issue-active cycle):")
        st = []
        for i, k in enumerate(hdr):
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        for v, k in sorted(st, reverse=True)[:8]:
            print("    %-28s %.3f" % (k, v))
    if rep.endswith(".csv"):
        import gzip
        import os
        sp = rep.replace(".raw.csv", ".source.csv.gz")
        src = gzip.open(sp, "rt").read() if os.path.exists(sp) else ""
    else:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next((i for i, r in enumerate(rows) if "# Samples" in r), None)
    if hi is not None:
        h = rows[hi]
        iS, iSrc = h.index("# Samples"), h.index("Source")
        data, seen = [], set()
        for r in rows[hi + 1:]:
            try:
                n = int(r[iS])
            except (ValueError, IndexError):
                continue
            key = (r[0], r[iSrc])
            if key in seen:
                continue
            seen.add(key)
            data.append((n, r[iSrc]))
        tot = sum(n for n, _ in data) or 1
        print("  top sampled SASS instructions (share of warp-state samples):")
        for n, sline in sorted(data, reverse=True)[:12]:
            print("    %5.1f%%  %s" % (100.0 * n / tot, sline.strip()[:100]))


if __name__ == "__main__":
    main(sys.argv[1])
