#!/bin/sh
# Round-2 evidence run, tile kernels only (after the row-block hand-out order changed; B200, one GPU): GPU tests, the default bench line and the reference arm, the ncu launch list of
# the default bench command and one `ncu --set full` capture per kernel family at the bench's own batch size.
# Usage: gpurun -- sh profiles/scripts/capture_r2.sh ; then, here, python profiles/scripts/make_traffic.py r02
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests.txt 2>&1
tail -3 gpurun_out/r02_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; tail -2 gpurun_out/r02_smoke.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2>&1
# launch list of the default command (short): the tabulation kernel's share of the step
python bench.py --steps 4 --warmup 3 --no-cpu --no-legs > gpurun_out/plain_default.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_p8_lattice.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-legs > gpurun_out/ncu_launches.log 2>&1
cap() {  # name workload kernel-regex extra-flags launches-to-skip
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-legs --workload $2 --e2e-points 1024 --e2e-steps 1 $4"
  $CMD > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$3 -s ${5:-0} -c 1 -o gpurun_out/r02_prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  # gpurun brings back at most 64 MiB: keep the raw metric page, drop the report
  ncu -i gpurun_out/r02_prof_$1.ncu-rep --page raw --csv > gpurun_out/r02_raw_$1.csv 2>/dev/null
  rm -f gpurun_out/r02_prof_$1.ncu-rep
}
cap mma_p8 p8_tet_o2 k_mma "--flags 4" 1      # launch 0 is the 96-point self-check of the derived path
cap mma_p8_spectral p8_spectral_tet_o2 k_mma "" 1
cap mma_n2curl n2curl4_tet_o1 k_mma "" 1
# split-cell tile kernel (not a bench workload): Walkington tet order 2 at 2^20 points
FIATB200_FLAGS=0 python profiles/scripts/bench_case.py walkington_tet_o2 > gpurun_out/plain_cells.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_mma_cells -s 1 -c 1 -o gpurun_out/r02_prof_cells_walkington \
    python profiles/scripts/bench_case.py walkington_tet_o2 > gpurun_out/ncu_cells.log 2>&1
ncu -i gpurun_out/r02_prof_cells_walkington.ncu-rep --page raw --csv > gpurun_out/r02_raw_cells_walkington.csv 2>/dev/null
rm -f gpurun_out/r02_prof_cells_walkington.ncu-rep
sh profiles/scripts/bench_cases.sh gpurun_out/r02_other_elements.txt 4 gn_tet_o2 walkington_tet_o2 hct4_tri_o2 hct5_tri_o2 hct6_tri_o2 p10_tri_o2 p12_tri_o2 p12_spectral_tri_o2 p6_tet_o1 p5_tet_o3 ned1_3_tet_o1 on4_none_3d_o2 regge2_tet_o1 argyris_tri_o2 p10_spectral_tet_o2 enriched_p4s_bubble5_tet_o2 alfeld_sorokina_tet_adv_o2 > /dev/null
