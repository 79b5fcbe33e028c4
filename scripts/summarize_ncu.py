"""Summarise ncu artefacts from gpurun_out/ into small tracked files under profiles/.
  python scripts/summarize_ncu.py launches gpurun_out/X_launches.csv profiles/NAME.json
  python scripts/summarize_ncu.py full gpurun_out/X.ncu-rep profiles/NAME.csv
"""
import collections, csv, json, subprocess, sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__cluster_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    h = rows[0]
    iN, iV, iU, iG, iB = (h.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    agg, total, seq = collections.OrderedDict(), 0.0, []
    for r in rows[1:]:
        v = float(r[iV].replace(",", "")) / {"ns": 1e3, "us": 1.0, "ms": 1e-3}[r[iU]]
        name = r[iN].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, {"launches": 0, "us": 0.0})
        a["launches"] += 1
        a["us"] += v
        total += v
        seq.append([name, r[iG], round(v, 2)])
    for a in agg.values():
        a["share"] = round(a["us"] / total, 4)
        a["us"] = round(a["us"], 1)
    json.dump({"source": src, "metric": "gpu__time_duration.sum (cold-cache, serialised: compare shares)",
               "launches": len(seq), "total_us": round(total, 1),
               "by_kernel": dict(sorted(agg.items(), key=lambda kv: -kv[1]["us"])), "sequence": seq},
              open(dst, "w"), indent=0)
    print("wrote", dst)


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    idx = [i for i, k in enumerate(h) if k in KEEP or k in ("ID", "Kernel Name")]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([h[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i][:120] for i in idx])
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
