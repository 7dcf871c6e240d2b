#!/bin/bash
# Profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one --set full capture per kernel family.
# Usage on the GPU box: bash tools/profile.sh <tag>   -> gpurun_out/<tag>_*.{csv,ncu-rep,log}
TAG=${1:-r01}
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --chunks ${CHUNKS:-32}"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
tail -c 600 gpurun_out/${TAG}_plain.log
# every launch of the LAST (timed + profiled) steps with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
# kernel regex : launches to skip (warm-up) : launches to capture
for SPEC in ${KERNELS:-attention_tm_kernel:20:1 gemm2_tn_kernel:40:4 mel_stft_kernel:2:1 layernorm_kernel:20:1}; do
  K=${SPEC%%:*}; REST=${SPEC#*:}; S=${REST%%:*}; C=${REST#*:}
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o gpurun_out/${TAG}_$K $CMD > gpurun_out/${TAG}_ncu_$K.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu_$K.log
done
ls -la gpurun_out/ | tail -12
