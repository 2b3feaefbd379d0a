#!/bin/sh
# Round-2 step r (B200, one GPU): register-operand split-cell kernel, 256- vs 512-thread CTAs, row blocks per step.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "split_cell or self_checked or no_subcell" > gpurun_out/r02_r_tests.txt 2>&1
tail -3 gpurun_out/r02_r_tests.txt
CASES="walkington_tet_o2 gn_tet_o2 alfeld_sorokina_tet_adv_o2 hct5_tri_o2 hct6_tri_o2"
sh profiles/scripts/bench_cases.sh gpurun_out/r02_r_cases_default.txt 0 $CASES > /dev/null; cat gpurun_out/r02_r_cases_default.txt
FIATB200_CELLS_THREADS=256 sh profiles/scripts/bench_cases.sh gpurun_out/r02_r_cases_t256.txt 0 walkington_tet_o2 hct6_tri_o2 > /dev/null; cat gpurun_out/r02_r_cases_t256.txt
FIATB200_CELLS_THREADS=512 sh profiles/scripts/bench_cases.sh gpurun_out/r02_r_cases_t512.txt 0 gn_tet_o2 alfeld_sorokina_tet_adv_o2 hct5_tri_o2 > /dev/null; cat gpurun_out/r02_r_cases_t512.txt
FIATB200_CELLS_RB=2 sh profiles/scripts/bench_cases.sh gpurun_out/r02_r_cases_rb2.txt 0 gn_tet_o2 alfeld_sorokina_tet_adv_o2 hct5_tri_o2 > /dev/null; cat gpurun_out/r02_r_cases_rb2.txt
FIATB200_CELLS_RB=4 sh profiles/scripts/bench_cases.sh gpurun_out/r02_r_cases_rb4.txt 0 walkington_tet_o2 hct6_tri_o2 > /dev/null; cat gpurun_out/r02_r_cases_rb4.txt
