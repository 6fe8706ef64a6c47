#!/bin/bash
# One GPU-box session: diagnostics, parity tests, bench, ncu evidence.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
python tools/gpu_diag.py all ${DIAG:-attn perf bwdunits} 2>&1 | tee gpurun_out/diag_stdout.txt | grep -v '"ok": true' | tail -70
echo "=== pytest -m gpu"
python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "=== bench"
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_stderr.txt | tee gpurun_out/bench.json
python bench.py --mode train --steps 5 --warmup 3 2>>gpurun_out/bench_stderr.txt | tee gpurun_out/bench_train.json
tail -5 gpurun_out/bench_stderr.txt
if [ "${NCU:-0}" = "1" ]; then
echo "=== ncu launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-breakdown > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
echo "=== ncu full: gemm, attention"
ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 14 -c 4 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-breakdown > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 3 -c 1 -o gpurun_out/prof_attn -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-breakdown > gpurun_out/ncu_attn.log 2>&1
fi
ls gpurun_out/
