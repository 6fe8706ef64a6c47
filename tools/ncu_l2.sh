#!/bin/bash
# L2 / crossbar pressure of the K = 768 GEMMs (fc1, qkv): one full ncu capture each, only the memory-side percentages printed.
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:gemm -s 18 -c 4 -o gpurun_out/prof_l2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-breakdown > gpurun_out/ncu_l2.log 2>&1
ncu -i gpurun_out/prof_l2.ncu-rep --page raw --csv > gpurun_out/l2_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/l2_raw.csv")))
h, u = rows[0], rows[1]
keys = [i for i, k in enumerate(h) if any(s in k for s in ("lts__t_bytes.sum", "lts__throughput", "lts__t_sectors_srcunit_tex.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_l1tex2xbar", "lts__t_sector_hit_rate", "sm__pipe_tensor_cycles_active.avg.pct", "gpu__time_duration.sum", "dram__throughput.avg.pct", "lts__d_sectors_fill",
        "smsp__cycles_active.avg", "lts__t_sectors_op_read.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem", "lts__average_t_sector"))]
for r in rows[2:]:
    print("---", r[h.index("Kernel Name")][:60])
    for i in keys:
        print("   ", h[i], "=", r[i], u[i])
PY
rm -f gpurun_out/prof_l2.ncu-rep
