#!/bin/sh
# Round-2 step n (B200, one GPU): register-operand split-cell kernel (cells_reg.cuh) -- parity, throughput against
# the segment kernel (FIATB200_CELLS_REG=0), ncu of Walkington's element.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split_cell or self_checked or no_subcell or no_write or kernel_selection or golden" > gpurun_out/r02_n_tests.txt 2>&1
tail -8 gpurun_out/r02_n_tests.txt
CASES="walkington_tet_o2 gn_tet_o2 alfeld_sorokina_tet_adv_o2 hct5_tri_o2 hct6_tri_o2 p2_alfeld_tet_o2 hct4_tri_o2"
sh profiles/scripts/bench_cases.sh gpurun_out/r02_n_cases_reg.txt 0 $CASES > /dev/null; cat gpurun_out/r02_n_cases_reg.txt
FIATB200_CELLS_REG=0 sh profiles/scripts/bench_cases.sh gpurun_out/r02_n_cases_seg.txt 0 $CASES > /dev/null; cat gpurun_out/r02_n_cases_seg.txt
CMD="python profiles/scripts/bench_case.py walkington_tet_o2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cells_reg -s 3 -c 1 -o gpurun_out/r02_n_prof_cells $CMD > gpurun_out/ncu_cells_reg.log 2>&1
ncu -i gpurun_out/r02_n_prof_cells.ncu-rep --page raw --csv > gpurun_out/r02_n_raw_cells_reg.csv 2>/dev/null
ncu -i gpurun_out/r02_n_prof_cells.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r02_n_src_cells_reg.csv 2>/dev/null
rm -f gpurun_out/r02_n_prof_cells.ncu-rep
tail -3 gpurun_out/ncu_cells_reg.log
