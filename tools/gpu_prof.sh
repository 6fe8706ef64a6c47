#!/bin/bash
# ncu evidence for the current build (1 GPU): launch list of a bench run + full captures of the top kernels.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-breakdown --train-steps 0 > gpurun_out/ncu_bench.log 2>&1
for k in gemm:17:8 attention_fwd:3:1 logmel:1:1 layernorm_to16:3:1; do
  IFS=: read name skip count <<< "$k"
  ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c $count -o gpurun_out/prof_$name -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-breakdown --train-steps 0 > gpurun_out/ncu_$name.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:mel_ingest -s 4 -c 1 -o gpurun_out/prof_mel_ingest -f \
    python bench.py --mode ingest --steps 3 > gpurun_out/ncu_mel_ingest.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_bwd -s 2 -c 1 -o gpurun_out/prof_attention_bwd -f \
    python bench.py --mode train --batch 16 --steps 1 --warmup 3 > gpurun_out/ncu_attention_bwd.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1020 -c 340 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1
# the .ncu-rep files are too large to travel back (64 MiB cap): summarise them here, keep only the text
export NCU_SUMMARY_DIR=gpurun_out/profiles_out
mkdir -p $NCU_SUMMARY_DIR
python tools/ncu_summarize.py ${PREFIX:-r02c} > gpurun_out/summarize.log 2>&1
python tools/gemm_traffic.py ${PREFIX:-r02c} >> gpurun_out/summarize.log 2>&1
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out gpurun_out/profiles_out
