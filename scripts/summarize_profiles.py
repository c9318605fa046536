"""Turns the raw ncu outputs in gpurun_out/ (scripts/capture_profiles.sh) into the tracked
summaries under profiles/ (run in the dev container; ncu can read reports without a GPU)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def launches():
    path = os.path.join(SRC, "launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hdr, agg, order = None, collections.defaultdict(list), []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                v = float(d["Metric Value"].replace(",", ""))
                unit = d.get("Metric Unit", "ns")
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
                agg[d["Kernel Name"]].append(v)
                order.append((d["ID"], d["Kernel Name"], v))
    total = sum(sum(v) for v in agg.values())
    ours = total
    with open(os.path.join(OUT, f"launches_{TAG}.md"), "w") as f:
        f.write(f"# ncu launch list ({TAG}): `python bench.py --steps 2 --warmup 3 --skip-e2e --skip-head` under "
                "`ncu --metrics gpu__time_duration.sum --clock-control none`\n\n"
                "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes. Filtered to the\n"
                "product's kernels (`-k regex:emit_kernel|viterbi_|logmel...`): 5 steps (3 warm-up + 2 timed) of the\n"
                "2000-clip workload; nothing else runs inside a step at N=1.\n\n"
                "| kernel | launches | total us | mean us | share of all | share of la:: kernels |\n|---|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            mine = True
            f.write(f"| `{k[:90]}` | {len(v)} | {sum(v):.1f} | {sum(v)/len(v):.1f} | {100*sum(v)/total:.1f}% | "
                    f"{(100*sum(v)/ours if mine and ours else 0):.1f}% |\n")
    with open(os.path.join(OUT, f"launches_{TAG}.csv"), "w") as f:
        f.write("id,kernel,us\n")
        for i, k, v in order:
            f.write(f"{i},\"{k}\",{v:.3f}\n")
    print("launch list:", len(order), "launches")


def full(name, pretty, cmd=None):
    rep = os.path.join(SRC, f"{name}.ncu-rep")
    if not os.path.exists(rep):
        return None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = {}
    with open(os.path.join(OUT, f"{pretty}_{TAG}.md"), "w") as f:
        f.write(f"# ncu --set full: {pretty} ({TAG})\n\nCommand: " + (cmd or (
                f"`ncu --set full --clock-control none --import-source on -k regex:{name} -s 3 -c 2 "
                "python bench.py --clips 400 --steps 2 --warmup 3 --skip-e2e --skip-head` (400 clips keep the 40 replay passes short; "
                "N1: `python scripts/bench_head.py 400`).")) + "\n\n")
        for li, vals in enumerate(rows[2:]):
            d = dict(zip(hdr, vals))
            f.write(f"## launch {li}: `{d.get('Kernel Name','')[:100]}` grid {d.get('launch__grid_size','?')} block {d.get('launch__block_size','?')}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            f.write("\n")
            if li == 0:
                out = {k: d[k] for k in KEYS if k in d}
                out["_units"] = {k: units[hdr.index(k)] for k in KEYS if k in d}
    return out


def to_bytes(v, unit):
    x = float(str(v).replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    launches()
    k2 = full("emit_kernel", "k2_emit")
    full("logmel_kernel", "k1_logmel")
    full("viterbi_wave_kernel", "k3_viterbi")
    full("k3_wave_batch", "k3_viterbi_batch2000",
         "`ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:viterbi_wave -s 2 -c 1 "
         "python scripts/k3_single.py opencpop 2000 5`: K3 alone on the bench's 2 000-clip Opencpop-shaped batch (small V, the "
         "emission rows are what K2 would have written).")
    full("head_lse_kernel", "n1_head_lse")
    if k2:
        rd = to_bytes(k2["dram__bytes_read.sum"], k2["_units"]["dram__bytes_read.sum"])
        wr = to_bytes(k2["dram__bytes_write.sum"], k2["_units"]["dram__bytes_write.sum"])
        with open(os.path.join(OUT, "k2_traffic.json"), "w") as f:
            json.dump({"note": "dram__bytes_read.sum + dram__bytes_write.sum of la::emit_kernel<0>, one launch, 400-clip "
                               "capture (profiles/k2_emit_%s.md); bench.py scales it to its own launch by frames" % TAG,
                       "dram_bytes_per_launch_400clips": rd + wr, "frames_400clips": 196683,
                       "dram_bytes_per_frame": (rd + wr) / 196683}, f, indent=1)
        print("k2 traffic", rd + wr)
