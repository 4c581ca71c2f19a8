"""Turn an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) and an ncu
launch list into the tracked summaries under profiles/:

    python scripts/summarise_profile.py gpurun_out/prof_TAG_raw.csv gpurun_out/launches_TAG.csv r1

writes profiles/<round>_kernels_ncu.txt (one block of key counters per kernel of the step),
profiles/<round>_launches_C3.csv (+ .txt, the per-kernel share of one step) and profiles/traffic.json
(dram bytes per launch of every kernel / stage, read by bench.py for `roofline.traffic`)."""
import csv, json, os, sys, collections

raw, launches, rnd = sys.argv[1], sys.argv[2], sys.argv[3]
extra_raw = sys.argv[4:]          # more raw exports (stand-alone stage kernels)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
out = ["# ncu --set full --clock-control none --import-source on, one launch of every kernel of a bench step",
       "#   python scripts/ncu_case.py (round 2; round 1: bench.py --steps 1 --warmup 3 --no-e2e --no-cpu)   (C3: global ORCA12 3059x4322, 633122 track points)",
       "# caches are flushed before every profiled launch and launches are serialised: compare SHARES with the live run", ""]
traffic = {}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for path in [raw] + extra_raw:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    ki = hdr.index("Kernel Name")
    if path != raw:
        out += ["# ---- stand-alone stage kernel (per-stage table of bench.py): " + os.path.basename(path), ""]
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "")
        out.append("== " + name)
        for k, i in idx:
            out.append("  %-82s %16s %s" % (k, r[i], units[i]))
        out.append("")
        f = lambda k: float(r[hdr.index(k)].replace(",", ""))
        u = lambda k: units[hdr.index(k)]
        b_ = f("dram__bytes_read.sum") * scale[u("dram__bytes_read.sum")] + f("dram__bytes_write.sum") * scale[u("dram__bytes_write.sum")]
        traffic.setdefault(name, []).append(b_)
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", rnd + "_kernels_ncu.txt"), "w").write("\n".join(out))

# launch list: last complete step
lr = list(csv.reader(open(launches)))
lh = next(r for r in lr if "Kernel Name" in r)
i0 = lr.index(lh)
a, b, c, d = lh.index("Kernel Name"), lh.index("Metric Name"), lh.index("Metric Value"), lh.index("ID")
rec = collections.OrderedDict()
for r in lr[i0 + 1:]:
    if len(r) > c:
        e = rec.setdefault(int(r[d]), {"name": r[a].split("(")[0].replace("void ", "")})
        e[r[b]] = float(r[c].replace(",", ""))
ids = list(rec)
starts = [k for k in ids if rec[k]["name"] == "k_near_search"]
steps = [[k for k in ids if a_ <= k < b_] for a_, b_ in zip(starts[:-1], starts[1:])]
step = [st for st in steps if any(rec[k]["name"].startswith("k_mesh_weights_interp") for k in st)][-1]     # a whole K1-K4 step
step = step[:[rec[k]["name"].startswith("k_path_fix") for k in step].index(True) + 1]
tot = sum(rec[k]["gpu__time_duration.sum"] for k in step)
txt = ["# one step of bench.py (C3) under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
       "--clock-control none`: serialised, cold caches", "# kernel, duration us, share of the step, dram read MB, dram write MB"]
for k in step:
    e = rec[k]
    txt.append("%-22s %8.2f  %5.1f%%  %8.2f  %8.2f" % (e["name"], e["gpu__time_duration.sum"] / 1e3,
               100 * e["gpu__time_duration.sum"] / tot, e.get("dram__bytes_read.sum", 0) / 1e6, e.get("dram__bytes_write.sum", 0) / 1e6))
txt.append("%-22s %8.2f" % ("sum", tot / 1e3))
open(os.path.join(ROOT, "profiles", rnd + "_launches_C3.txt"), "w").write("\n".join(txt) + "\n")
with open(os.path.join(ROOT, "profiles", rnd + "_launches_C3.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(lh)
    for r in lr[i0 + 1:]:
        if len(r) > d and r[d].isdigit() and int(r[d]) in step:
            w.writerow(r)
K1 = ["k_near_search", "k_near_slow", "k_chain", "k_tail", "k_compact", "k_walk<0>", "k_valid", "k_walk<1>"]


def pick(prefix):          # template arguments changed between rounds: match by prefix
    for k, v in per.items():
        if k.startswith(prefix):
            return v
    return 0.0
per = {k: sum(v) / len(v) for k, v in traffic.items()}
for k in step:      # kernels the full capture did not include: take the launch list's counters
    e = rec[k]
    per.setdefault(e["name"], e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0))
def pick_exact(base):      # "k_mesh_weights" or "k_mesh_weights<int>", not "k_mesh_weights_interp<...>"
    for k, v in per.items():
        if k == base or k.startswith(base + "<"):
            return v
    return 0.0


stage = {"K1_nearest": sum(per.get(k, 0.0) for k in K1), "K2K3_mesh_weights": pick_exact("k_mesh_weights"),
         "K4_interp_gather": pick("k_interp<float")}
stage["K2K3K4_bulk"] = pick("k_mesh_weights_interp<float")
stage["K1-K4_path"] = stage["K1_nearest"] + stage["K2K3K4_bulk"] + pick("k_path_fix")
stage["K1K2K3_mapping"] = stage["K1_nearest"] + stage["K2K3_mesh_weights"]
json.dump({"source": os.path.basename(raw), "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
           "kernels": per, **stage}, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("\n".join(txt))
