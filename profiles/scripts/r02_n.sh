#!/bin/sh
# Round-2 step n (B200, one GPU): register-operand split-cell kernel (cells_reg.cuh) -- parity, throughput against
# the segment kernel (FIATB200_CELLS_REG=0), ncu of Walkington's element; quick plan; the reference's unit tests on
# the device drop-in (default selection); to_riesz.
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_parity.py tests/test_gpu_callers.py -m gpu -q -k "split_cell or self_checked or quick_plan or trace_and_quadrature or single_point or to_riesz or no_subcell" > gpurun_out/r02_n_tests.txt 2>&1
tail -12 gpurun_out/r02_n_tests.txt
CASES="walkington_tet_o2 gn_tet_o2 alfeld_sorokina_tet_adv_o2 hct5_tri_o2 hct6_tri_o2"
sh profiles/scripts/bench_cases.sh gpurun_out/r02_n_cases_reg.txt 0 $CASES > /dev/null; cat gpurun_out/r02_n_cases_reg.txt
FIATB200_CELLS_REG=0 sh profiles/scripts/bench_cases.sh gpurun_out/r02_n_cases_seg.txt 0 walkington_tet_o2 gn_tet_o2 hct5_tri_o2 > /dev/null; cat gpurun_out/r02_n_cases_seg.txt
rm -f /tmp/dropin_stats.jsonl
( cd /tmp && PYTHONPATH=/root/repo/oracle/_ref:/root/repo FIATB200_DROPIN=device FIATB200_DROPIN_STATS=/tmp/dropin_stats.jsonl \
  timeout 330 python -m pytest -p oracle.dropin_plugin -q -p no:cacheprovider -c /dev/null --rootdir /tmp -n 3 --tb=line \
  -k "not macro_gem and not macro_sympy" --durations=12 \
  /root/repo/oracle/_ref/ref_tests/test_fiat.py /root/repo/oracle/_ref/ref_tests/test_tensor_product.py \
  /root/repo/oracle/_ref/ref_tests/test_regge_hhj.py /root/repo/oracle/_ref/ref_tests/test_macro.py \
  /root/repo/oracle/_ref/ref_tests/test_hdivtrace.py /root/repo/oracle/_ref/ref_tests/test_discontinuous_taylor.py \
  /root/repo/oracle/_ref/ref_tests/test_serendipity.py /root/repo/oracle/_ref/ref_tests/test_quadrature_element.py \
  ) > gpurun_out/r02_n_reference_suite_device.txt 2>&1
tail -40 gpurun_out/r02_n_reference_suite_device.txt | cut -c1-260
CMD="python profiles/scripts/bench_case.py walkington_tet_o2"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_cells_reg -s 3 -c 1 -o gpurun_out/r02_n_prof_cells $CMD > gpurun_out/ncu_cells_reg.log 2>&1
ncu -i gpurun_out/r02_n_prof_cells.ncu-rep --page raw --csv > gpurun_out/r02_n_raw_cells_reg.csv 2>/dev/null
ncu -i gpurun_out/r02_n_prof_cells.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r02_n_src_cells_reg.csv 2>/dev/null
rm -f gpurun_out/r02_n_prof_cells.ncu-rep
tail -3 gpurun_out/ncu_cells_reg.log
