#!/bin/sh
# Round-2 measurement step (B200, one GPU): GPU tests, throughput of the general (DMMA tile) paths, ncu of k_mma.
# usage: gpurun -- sh profiles/scripts/r02_step.sh TAG [skip-tests]
TAG=${1:-a}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_${TAG}_tests.txt 2>&1
  tail -5 gpurun_out/r02_${TAG}_tests.txt
fi
# general paths at bench scale
sh profiles/scripts/quick_bench.sh "--flags 4" p8_tet_o2 > /dev/null; cp gpurun_out/quick.txt gpurun_out/r02_${TAG}_quick.txt
sh profiles/scripts/quick_bench.sh "" n2curl4_tet_o1 hct_o2 ps12_o2 gll_q10_hex_o1 p3_tri_o1 p8_tet_o2 > /dev/null; cat gpurun_out/quick.txt >> gpurun_out/r02_${TAG}_quick.txt
sh profiles/scripts/bench_cases.sh gpurun_out/r02_${TAG}_cases.txt 4 gn_tet_o2 walkington_tet_o2 hct4_tri_o2 p10_tri_o2 p6_tet_o1 p5_tet_o3 ned1_3_tet_o1 on4_none_3d_o2 enriched_p4s_bubble5_tet_o2 argyris_tri_o2 > /dev/null
cat gpurun_out/r02_${TAG}_quick.txt gpurun_out/r02_${TAG}_cases.txt
cap() {  # name workload kernel-regex extra-flags
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu --workload $2 --e2e-points 1024 --e2e-steps 1 $4"
  ncu --set full --clock-control none --import-source on -k regex:$3 -c 1 -o gpurun_out/r02_${TAG}_prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/r02_${TAG}_prof_$1.ncu-rep --page raw --csv > gpurun_out/r02_${TAG}_raw_$1.csv 2>/dev/null
  ncu -i gpurun_out/r02_${TAG}_prof_$1.ncu-rep --page source --csv > gpurun_out/r02_${TAG}_src_$1.csv 2>/dev/null
  rm -f gpurun_out/r02_${TAG}_prof_$1.ncu-rep
}
cap mma_p8 p8_tet_o2 k_mma "--flags 4"
cap mma_n2curl n2curl4_tet_o1 k_mma ""
