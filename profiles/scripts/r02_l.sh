#!/bin/sh
# Round-2 step k: row-block hand-out order of the tile kernels (plan-side only).
mkdir -p gpurun_out
q() {  # label env workload flags
  env $2 timeout 300 python bench.py --steps 20 --no-cpu --no-legs --e2e-points 1024 --e2e-steps 1 --workload $3 --flags $4 2>gpurun_out/r02_l_err_$1.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', '$2', '$3', d['config']['kernel'], round(d['value']/1e9,1), 'Gval/s frac', round(d['roofline']['frac'],3), 'ms/launch', round(d['roofline']['kernel_ms'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> gpurun_out/r02_l_exp.txt
}
: > gpurun_out/r02_l_exp.txt
for pol in mixed strata16 strata8 strata4; do
  q p8_$pol FIATB200_RB_ORDER=$pol p8_tet_o2 4
  q p8s_$pol FIATB200_RB_ORDER=$pol p8_spectral_tet_o2 0
  q n2_$pol FIATB200_RB_ORDER=$pol n2curl4_tet_o1 0
done
cat gpurun_out/r02_l_exp.txt
for pol in strata16 strata8; do
  FIATB200_RB_ORDER=$pol sh profiles/scripts/bench_cases.sh gpurun_out/r02_l_cases_$pol.txt 4 gn_tet_o2 walkington_tet_o2 p10_tri_o2 p6_tet_o1 > /dev/null; cat gpurun_out/r02_l_cases_$pol.txt
done
