"""profiles/<prefix>_gemm_traffic.json from gpurun_out/prof_gemm.ncu-rep (ncu --set full capture of 8 consecutive encoder GEMM
launches): measured DRAM bytes per launch of the GEMM family; bench.py reports it as roofline.traffic.

    python tools/gemm_traffic.py r01b"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROLE = {"gemm2_tn_kernel<0, 0>": "gemm_qkv", "gemm2_tn_kernel<0, 1>": "gemm_fc1", "gemm_tn_kernel<0, 2, 0, 0>": "gemm_proj", "gemm_tn_kernel<0, 1, 0, 0>": "gemm_fc1",
        "gemm2_tn_kernel<0, 2>": "gemm_fc2", "gemm_tn_kernel<0, 0, 0, 0>": "gemm_qkv"}


def main(prefix):
    rep = os.path.join(ROOT, "gpurun_out", "prof_gemm.ncu-rep")
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")

    def val(row, key):
        i = hdr.index(key)
        v = float(row[i].replace(",", ""))
        u = units[i].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3}.get(u, 1.0)

    per = {}
    for row in rows[2:]:
        role = next((v for k, v in ROLE.items() if k in row[ki]), None)
        if role is None or role in per:
            continue
        per[role] = dict(kernel=row[ki], dram_read_bytes=val(row, "dram__bytes_read.sum"), dram_write_bytes=val(row, "dram__bytes_write.sum"),
                         duration_us=float(row[hdr.index("gpu__time_duration.sum")].replace(",", "")),
                         tensor_pipe_active_pct=float(row[hdr.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")]))
    assert set(per) == {"gemm_qkv", "gemm_proj", "gemm_fc1", "gemm_fc2"}, per.keys()
    tot = 12 * sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in per.values())
    out = dict(source=f"ncu --set full --clock-control none, profiles/{prefix}_prof_gemm_ncu_full.txt (config 3, B=64, fp16 operands)",
               per_launch=per, gemm_family_bytes_per_step=tot, launches_per_step=48)
    with open(os.path.join(os.environ.get("NCU_SUMMARY_DIR") or os.path.join(ROOT, "profiles"), f"{prefix}_gemm_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: (round(v["duration_us"], 1), round((v["dram_read_bytes"] + v["dram_write_bytes"]) / 1e6)) for k, v in per.items()}), tot)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01b")
