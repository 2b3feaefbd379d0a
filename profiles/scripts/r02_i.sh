#!/bin/sh
# Round-2 step i: full GPU tests; split-cell tile kernel with pipelined segments; annealed packing.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_i_tests.txt 2>&1
tail -5 gpurun_out/r02_i_tests.txt
q() {  # label env workload flags
  env $2 timeout 300 python bench.py --steps 20 --no-cpu --no-legs --e2e-points 1024 --e2e-steps 1 --workload $3 --flags $4 2>gpurun_out/r02_i_err_$1.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); ev=d.get('e2e_evaluate') or {}
print('$1', '$2', '$3', d['config']['kernel'], round(d['value']/1e9,1), 'Gval/s frac', round(d['roofline']['frac'],3), 'ms/launch', round(d['roofline']['kernel_ms'],4), 'e2e_eval Mpt/s', round(ev.get('value',0)/1e6,1), d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> gpurun_out/r02_i_exp.txt
}
: > gpurun_out/r02_i_exp.txt
q p8 A=1 p8_tet_o2 4
q p8s A=1 p8_spectral_tet_o2 0
q n2 A=1 n2curl4_tet_o1 0
cat gpurun_out/r02_i_exp.txt
sh profiles/scripts/bench_cases.sh gpurun_out/r02_i_cases.txt 4 gn_tet_o2 walkington_tet_o2 hct4_tri_o2 hct5_tri_o2 hct6_tri_o2 p10_tri_o2 p6_tet_o1 p12_tri_o2 alfeld_sorokina_tet_adv_o2 > /dev/null
cat gpurun_out/r02_i_cases.txt
