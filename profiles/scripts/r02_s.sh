#!/bin/sh
# Round-2 final step (B200, one GPU): the whole GPU suite, smoke, the default bench line, ncu of the register-operand
# split-cell kernel on the Guzman-Neilan element.
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02_s_tests.txt 2>&1
tail -14 gpurun_out/r02_s_tests.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_s_bench_default.json 2> gpurun_out/r02_s_bench_default.err
tail -c 1200 gpurun_out/r02_s_bench_default.json
CMD="python profiles/scripts/bench_case.py gn_tet_o2"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_cells_reg -s 3 -c 1 -o gpurun_out/r02_s_prof_cells $CMD > gpurun_out/ncu_cells_reg.log 2>&1
ncu -i gpurun_out/r02_s_prof_cells.ncu-rep --page raw --csv > gpurun_out/r02_raw_cells_reg_gn.csv 2>/dev/null
rm -f gpurun_out/r02_s_prof_cells.ncu-rep
sh profiles/scripts/bench_cases.sh gpurun_out/r02_s_cases.txt 0 walkington_tet_o2 gn_tet_o2 alfeld_sorokina_tet_adv_o2 hct5_tri_o2 hct6_tri_o2 > /dev/null; cat gpurun_out/r02_s_cases.txt
