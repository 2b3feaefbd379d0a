#!/bin/sh
# Round-2 step c: caller tests, tile-kernel probes (B loads / stores removed), evaluate legs.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_callers.py tests/test_gpu_bench_parity.py -m gpu -x -q > gpurun_out/r02_c_tests.txt 2>&1
tail -5 gpurun_out/r02_c_tests.txt
q() {  # label env workload flags
  env $2 python bench.py --steps 20 --no-cpu --no-legs --e2e-points 1024 --e2e-steps 1 --workload $3 --flags $4 2>gpurun_out/r02_c_err_$1.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); ev=d.get('e2e_evaluate') or {}
print('$1', '$2', '$3', d['config']['kernel'], round(d['value']/1e9,1), 'Gval/s frac', round(d['roofline']['frac'],3), 'ms/launch', round(d['roofline']['kernel_ms'],4), 'e2e_eval Mpt/s', round(ev.get('value',0)/1e6,1))" >> gpurun_out/r02_c_exp.txt
}
: > gpurun_out/r02_c_exp.txt
q p8 A=1 p8_tet_o2 4
q p8_noB FIATB200_MMA_SKIP=4 p8_tet_o2 4
q p8_noST FIATB200_MMA_SKIP=8 p8_tet_o2 4
q p8_noB_noST FIATB200_MMA_SKIP=12 p8_tet_o2 4
q p8_norec_noB_noST FIATB200_MMA_SKIP=13 p8_tet_o2 4
q n2 A=1 n2curl4_tet_o1 0
q n2_t512 FIATB200_MMA_THREADS=512 n2curl4_tet_o1 0
q n2_nocontr FIATB200_MMA_SKIP=2 n2curl4_tet_o1 0
q n2_noB FIATB200_MMA_SKIP=4 n2curl4_tet_o1 0
q n2_noST FIATB200_MMA_SKIP=8 n2curl4_tet_o1 0
q n2_noB_noST FIATB200_MMA_SKIP=12 n2curl4_tet_o1 0
q hct A=1 hct_o2 0
q hex A=1 gll_q10_hex_o1 0
q p8_lattice A=1 p8_tet_o2 0
cat gpurun_out/r02_c_exp.txt
