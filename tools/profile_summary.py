"""Turn an ncu report (+ optional launch-list CSV) into the text summary committed under profiles/.

    python tools/profile_summary.py gpurun_out/x.ncu-rep profiles/x.txt ["command line that was profiled"]
    python tools/profile_summary.py --launches gpurun_out/launches.csv profiles/launches.txt
    python tools/profile_summary.py --traffic profiles/x_traffic.json workload=gpurun_out/x.ncu-rep:frames_per_launch[:profiles/x.txt] ...
"""
import csv, io, json, subprocess, sys

KEYS = ["Duration", "Elapsed Cycles", "SM Frequency", "SM Active Cycles", "Executed Ipc Active", "Issue Slots Busy", "Executed Instructions ",
        "Avg. Active Threads Per Warp", "Avg. Not Predicated Off Threads Per Warp", "Registers Per Thread", "Theoretical Occupancy", "Achieved Occupancy",
        "Grid Size", "Block Size", "Waves Per SM", "L1/TEX Hit Rate", "L2 Hit Rate", "DRAM Throughput", "Memory Throughput", "Mem Busy",
        "Compute (SM) Throughput", "Eligible Warps Per Scheduler", "Active Warps Per Scheduler", "No Eligible", "Warp Cycles Per Issued Instruction",
        "Branch Efficiency"]
RAW = ["dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
       "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "gpu__time_duration.sum"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = {}
    order = []
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        if name not in tot:
            tot[name] = [0, 0.0]
            order.append(name)
        tot[name][0] += 1
        tot[name][1] += v
    total = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write("ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache and serialised: compare SHARES)\n")
        f.write("source: %s\n\n%-70s %8s %14s %8s\n" % (src, "kernel", "launches", "total ns", "share"))
        for name in sorted(order, key=lambda n: -tot[n][1]):
            f.write("%-70s %8d %14.0f %7.2f%%\n" % (name[:70], tot[name][0], tot[name][1], 100 * tot[name][1] / total))
        f.write("\nper launch, in order:\n")
        for r in rows[1:]:
            f.write("%4s %-70s %12s ns\n" % (r[0], r[ki].split("(")[0].replace("void ", "")[:70], r[vi]))


def report(rep, dst, cmdline):
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = subprocess.run([sys.executable, __file__.replace("profile_summary.py", "ncu_lines.py"), rep, "0", "40"], capture_output=True, text=True).stdout
    with open(dst, "w") as f:
        f.write("ncu --set full --clock-control none --import-source on  (first captured launch)\nreport: %s\ncommand: %s\n\n" % (rep, cmdline))
        for line in det.splitlines():
            if "render_" in line and "Context" in line:
                f.write(line.strip() + "\n")
                break
        f.write("\n-- details page, selected --\n")
        seen = set()
        for line in det.splitlines():
            for k in KEYS:
                if line.strip().startswith(k) and k not in seen:
                    seen.add(k)
                    f.write(line.rstrip() + "\n")
        f.write("\n-- raw page, selected --\n")
        for k in RAW:
            if k in hdr:
                i = hdr.index(k)
                f.write("%-90s %-10s %s\n" % (k, units[i], vals[i]))
        f.write("\n-- hottest source lines (share of stall samples, share of executed warp instructions, active threads per instruction, top stall reasons) --\n")
        f.write(lines)


def traffic(dst, specs):
    """profiles/*_traffic.json: what bench.py's roofline reads -- DRAM bytes, executed warp instructions, issue-slot use and
    L2 / L1 hit rates of ONE render launch per workload, straight from the raw page of its `ncu --set full` capture."""
    out = {"_comment": "per workload: dram__bytes_read.sum + dram__bytes_write.sum, smsp__inst_executed.sum and friends of ONE "
                       "render_stream_kernel launch (a whole bench step), from the ncu --set full captures named in `report` "
                       "(tools/profile_summary.py --traffic). bench.py copies its workload's values into roofline.traffic / legs."}
    for spec in specs:
        workload, _, rest = spec.partition("=")
        parts = rest.split(":")
        rep, frames = parts[0], int(parts[1])
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]

        def get(name, scale_units=True):
            i = hdr.index(name)
            v = float(vals[i].replace(",", ""))
            u = units[i].lower()
            if scale_units:
                v *= {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "byte": 1.0}.get(u, 1.0)
            return v
        rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
        out[workload] = {"dram_read_mb": rd / 1e6, "dram_write_mb": wr / 1e6, "bytes": rd + wr, "frames_per_launch": frames,
                         "inst_executed": get("smsp__inst_executed.sum", False),
                         "issue_slots_busy_pct": get("sm__inst_issued.avg.pct_of_peak_sustained_active", False) if "sm__inst_issued.avg.pct_of_peak_sustained_active" in hdr else None,
                         "threads_per_inst": get("smsp__thread_inst_executed_per_inst_executed.ratio", False),
                         "l2_hit_pct": get("lts__t_sector_hit_rate.pct", False), "l1_hit_pct": get("l1tex__t_sector_hit_rate.pct", False),
                         "duration_under_ncu_ms": get("gpu__time_duration.sum", False) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")].lower(), 1.0),
                         "report": parts[2] if len(parts) > 2 else rep}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")


if __name__ == "__main__":
    if sys.argv[1] == "--traffic":
        traffic(sys.argv[2], sys.argv[3:])
    elif sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        report(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
