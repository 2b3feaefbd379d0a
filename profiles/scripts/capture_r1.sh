#!/bin/sh
# Round-1 evidence run (B200, one GPU): bench lines for every BASELINE configuration, the ncu launch list
# of the default bench command and one `ncu --set full` capture per kernel family at the bench's own
# batch size.  Usage: gpurun -- sh profiles/scripts/capture_r1.sh ; then, here,
# python profiles/scripts/make_traffic.py (summaries + traffic.json from the pulled reports).
set -x
mkdir -p gpurun_out
: > gpurun_out/r01_bench_all.jsonl
for w in p8_tet_o2 n2curl4_tet_o1 hct_o2 ps6_o2 ps12_o2 gll_q10_hex_o1 p3_tri_o1; do
  python bench.py --workload $w --steps 30 2>/dev/null | tail -1 >> gpurun_out/r01_bench_all.jsonl
done
python bench.py --workload p8_tet_o2 --flags 4 --steps 30 2>/dev/null | tail -1 >> gpurun_out/r01_bench_all.jsonl
python bench.py --workload n2curl4_tet_o1 --flags 16 --steps 30 --no-cpu 2>/dev/null | tail -1 >> gpurun_out/r01_bench_all.jsonl
python bench.py --workload p8_tet_o2 --flags 20 --steps 30 --no-cpu 2>/dev/null | tail -1 >> gpurun_out/r01_bench_all.jsonl
# launch list of the default command (short): the tabulation kernel's share of the step
python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/plain_default.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_p8_lattice.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
cap() {  # name workload kernel-regex launches-per-step extra-flags launches-to-skip
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu --workload $2 --e2e-points 1024 --e2e-steps 1 $5"
  $CMD > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$3 -s ${6:-0} -c $4 -o gpurun_out/r01_prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  # gpurun brings back at most 64 MiB: keep the raw metric page, drop the report
  ncu -i gpurun_out/r01_prof_$1.ncu-rep --page raw --csv > gpurun_out/r01_raw_$1.csv 2>/dev/null
  rm -f gpurun_out/r01_prof_$1.ncu-rep
}
cap lattice_p8 p8_tet_o2 k_lattice 1 "" 1      # launch 0 is the 96-point self-check of the product form
cap mma_p8 p8_tet_o2 k_mma 1 "--flags 4"
cap mma_n2curl_merged n2curl4_tet_o1 k_mma 1 ""
cap vals_hct hct_o2 k_vals 1 ""
cap vals_ps6 ps6_o2 k_vals 1 ""
cap vals_ps12 ps12_o2 k_vals 1 ""
cap tensor_hex gll_q10_hex_o1 k_tensor 1 ""
