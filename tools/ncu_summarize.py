"""Turn gpurun_out/*.ncu-rep and launches.csv into small text summaries under profiles/ (the judged evidence).

    python tools/ncu_summarize.py r01          # prefix for the output files
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.environ.get("NCU_SUMMARY_DIR") or os.path.join(ROOT, "profiles")   # on the GPU box: a directory under gpurun_out/

RAW_KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def ncu_csv(rep, page):
    r = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True)
    return list(csv.reader(io.StringIO(r.stdout)))


def summarize_rep(rep, f):
    rows = ncu_csv(rep, "raw")
    if len(rows) < 3:
        f.write(f"(could not read {rep})\n")
        return
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        f.write(f"--- {r[ki]}\n")
        for k in RAW_KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"  {k} = {r[i]} {units[i]}\n")
    src = ncu_csv(rep, "source")
    try:
        hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    except StopIteration:
        return
    h = src[hi]
    si, ei, wi = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    agg, tot, samples = collections.Counter(), 0, []
    for r in src[hi + 1:]:
        if len(r) <= max(wi, ei, si):
            if r and r[0] == "Kernel Name":
                break          # only the first captured launch
            continue
        try:
            n, s = int(r[ei]), int(r[wi])
        except ValueError:
            continue
        toks = r[si].split()
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        agg[op.split(".")[0]] += n
        tot += n
        samples.append((s, n, r[si].strip()))
    f.write(f"  executed warp-instructions (first captured launch): {tot}\n")
    for k, v in agg.most_common(14):
        f.write(f"    {100 * v / max(tot, 1):5.1f}%  {k}\n")
    ts = sum(s for s, _, _ in samples)
    f.write("  top stall sites (share of samples, SASS):\n")
    for s, n, line in sorted(samples, reverse=True)[:10]:
        f.write(f"    {100 * s / max(ts, 1):5.1f}%  {line[:100]}\n")


def summarize_launches(path, f):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0][:80]
        agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none, {sum(len(v) for v in agg.values())} launches, total {tot / 1e3:.1f} us "
            "(cold-cache, serialised: compare SHARES)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{sum(v) / 1e3:10.1f} us  {len(v):4d}x  avg {sum(v) / len(v) / 1e3:8.1f} us  {100 * sum(v) / tot:5.1f}%  {k}\n")


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    for src, dst in (("launches.csv", "launches_bench"), ("launches_train.csv", "launches_train_step")):
        lp = os.path.join(OUT, src)
        if os.path.exists(lp):
            try:
                with open(os.path.join(PROF, f"{prefix}_{dst}.txt"), "w") as f:
                    summarize_launches(lp, f)
            except StopIteration:      # the capture window held no kernels
                os.remove(os.path.join(PROF, f"{prefix}_{dst}.txt"))
                print(f"{src}: no launches captured")
    for name in sorted(os.listdir(OUT)):
        if name.endswith(".ncu-rep"):
            with open(os.path.join(PROF, f"{prefix}_{name[:-8]}_ncu_full.txt"), "w") as f:
                f.write(f"ncu --set full --clock-control none --import-source on  ({name})\n")
                summarize_rep(os.path.join(OUT, name), f)
    print(os.listdir(PROF))


if __name__ == "__main__":
    main()
