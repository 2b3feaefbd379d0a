#!/bin/sh
# Round-1 evidence run (B200, one GPU): bench lines for every BASELINE configuration and one
# `ncu --set full` capture per kernel family.  Usage: gpurun -- sh profiles/scripts/capture_r1.sh
set -x
mkdir -p gpurun_out
: > gpurun_out/r01_bench_all.jsonl
for w in p8_tet_o2 n2curl4_tet_o1 hct_o2 ps6_o2 ps12_o2 gll_q10_hex_o1 p3_tri_o1; do
  python bench.py --workload $w --steps 30 2>/dev/null | tail -1 >> gpurun_out/r01_bench_all.jsonl
done
python bench.py --workload p8_tet_o2 --flags 4 --steps 30 2>/dev/null | tail -1 >> gpurun_out/r01_bench_all.jsonl
cap() {  # name workload kernel-regex skip extra-flags
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu --workload $2 --batch 262144 --e2e-points 1024 --e2e-steps 1 $5"
  $CMD > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -o gpurun_out/r01_prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
}
cap mma_p8 p8_tet_o2 k_mma 7 "--flags 4"
cap mma_n2curl n2curl4_tet_o1 k_mma 7 ""
cap small_hct hct_o2 k_small 7 ""
cap tensor_hex gll_q10_hex_o1 k_tensor 7 ""
