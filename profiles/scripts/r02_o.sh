#!/bin/sh
# Round-2 step o (B200, one GPU): register-operand split-cell kernel with the jump on the k-block number -- parity,
# throughput (row blocks per step 2 / 4), ncu of Walkington's element.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "split_cell or self_checked or no_subcell" > gpurun_out/r02_o_tests.txt 2>&1
tail -4 gpurun_out/r02_o_tests.txt
CASES="walkington_tet_o2 gn_tet_o2 alfeld_sorokina_tet_adv_o2 hct5_tri_o2 hct6_tri_o2"
sh profiles/scripts/bench_cases.sh gpurun_out/r02_o_cases_reg.txt 0 $CASES > /dev/null; cat gpurun_out/r02_o_cases_reg.txt
FIATB200_CELLS_RB=4 sh profiles/scripts/bench_cases.sh gpurun_out/r02_o_cases_rb4.txt 0 walkington_tet_o2 hct6_tri_o2 > /dev/null; cat gpurun_out/r02_o_cases_rb4.txt
FIATB200_CELLS_RB=2 sh profiles/scripts/bench_cases.sh gpurun_out/r02_o_cases_rb2.txt 0 gn_tet_o2 alfeld_sorokina_tet_adv_o2 hct5_tri_o2 > /dev/null; cat gpurun_out/r02_o_cases_rb2.txt
CMD="python profiles/scripts/bench_case.py walkington_tet_o2"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_cells_reg -s 3 -c 1 -o gpurun_out/r02_o_prof_cells $CMD > gpurun_out/ncu_cells_reg.log 2>&1
ncu -i gpurun_out/r02_o_prof_cells.ncu-rep --page raw --csv > gpurun_out/r02_o_raw_cells_reg.csv 2>/dev/null
ncu -i gpurun_out/r02_o_prof_cells.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r02_o_src_cells_reg.csv 2>/dev/null
rm -f gpurun_out/r02_o_prof_cells.ncu-rep
tail -2 gpurun_out/ncu_cells_reg.log
