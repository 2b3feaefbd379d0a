#!/bin/sh
# Round-2 step j: per-path errors of the high-degree elements, full GPU tests, perf check after the packing criterion change.
mkdir -p gpurun_out
python profiles/scripts/diag_paths.py p12_tri_o2 p10_spectral_tet_o2 walkington_tet_o2 hct6_tri_o2 > gpurun_out/r02_j_diag.txt 2>&1; cat gpurun_out/r02_j_diag.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_j_tests.txt 2>&1
tail -5 gpurun_out/r02_j_tests.txt
sh profiles/scripts/quick_bench.sh "--flags 4 --no-legs" p8_tet_o2 > /dev/null; cp gpurun_out/quick.txt gpurun_out/r02_j_quick.txt
sh profiles/scripts/quick_bench.sh "--no-legs" p8_spectral_tet_o2 n2curl4_tet_o1 > /dev/null; cat gpurun_out/quick.txt >> gpurun_out/r02_j_quick.txt
cat gpurun_out/r02_j_quick.txt
sh profiles/scripts/bench_cases.sh gpurun_out/r02_j_cases.txt 4 gn_tet_o2 walkington_tet_o2 hct5_tri_o2 p12_tri_o2 p10_tri_o2 > /dev/null
cat gpurun_out/r02_j_cases.txt
