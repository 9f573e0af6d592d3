"""Summarise an `ncu --set full` report (one launch per kernel of a bench step) into
profiles/<tag>_ncu_top_kernels.txt and profiles/traffic.json.

    python tools/ncu_summary.py gpurun_out/r01_full.ncu-rep r01
"""
import csv, json, os, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out, traffic = [], {}
    out.append("ncu --set full --clock-control none, bench.py --steps 1 --warmup 3 (config C2, 5000 frames); one launch per kernel")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?")
        short = name.split("(")[0].replace("void ", "")
        out.append("== " + name[:100])
        for k in KEYS:
            if k in d:
                out.append("   %s = %s %s" % (k, d[k], u.get(k, "")))
        st = {h[len(STALLS):]: float(v) for h, v in d.items()
              if h.startswith(STALLS) and "not_issued" not in h and v not in ("", "n/a")}
        tot = sum(st.values()) or 1.0
        top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
        out.append("   stall samples: " + ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in top))

        def to_bytes(key):
            v = float(d.get(key, 0) or 0)
            unit = u.get(key, "")
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
        b = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        out.append("   dram read+write per launch = %.1f MB" % (b / 1e6))
        traffic[short] = max(traffic.get(short, 0.0), b)
    with open(os.path.join(root, "profiles", tag + "_ncu_top_kernels.txt"), "w") as fh:
        fh.write("\n".join(out) + "\n")
    with open(os.path.join(root, "profiles", "traffic.json"), "w") as fh:
        json.dump({"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, bench.py config C2 "
                           "(5000 frames), %s; the larger of the launches of a kernel" % tag,
                   "bytes_per_launch": traffic}, fh, indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main()
