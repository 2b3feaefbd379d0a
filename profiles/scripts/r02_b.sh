#!/bin/sh
# Round-2 step b: GPU tests, tile-kernel geometry experiments (phase timing via FIATB200_MMA_SKIP), full bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_b_tests.txt 2>&1
tail -5 gpurun_out/r02_b_tests.txt
q() {  # label env workload flags
  env $2 python bench.py --steps 20 --no-cpu --no-legs --e2e-points 1024 --e2e-steps 1 --workload $3 --flags $4 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', '$2', '$3', d['config']['kernel'], round(d['value']/1e9,1), 'Gval/s frac', round(d['roofline']['frac'],3), 'ms/launch', round(d['roofline']['kernel_ms'],4))" >> gpurun_out/r02_b_exp.txt
}
: > gpurun_out/r02_b_exp.txt
q p8_default A=1 p8_tet_o2 4
q p8_pt128 FIATB200_MMA_PT=128 p8_tet_o2 4
q p8_pt64_norec FIATB200_MMA_SKIP=1 p8_tet_o2 4
q p8_pt64_nocontr FIATB200_MMA_SKIP=2 p8_tet_o2 4
q p8_pt64_nothing FIATB200_MMA_SKIP=3 p8_tet_o2 4
q p8_pt128_nocontr "FIATB200_MMA_PT=128 FIATB200_MMA_SKIP=2" p8_tet_o2 4
q p8_pt128_nothing "FIATB200_MMA_PT=128 FIATB200_MMA_SKIP=3" p8_tet_o2 4
q p8s_default A=1 p8_spectral_tet_o2 0
q n2_default A=1 n2curl4_tet_o1 0
q n2_t512 FIATB200_MMA_THREADS=512 n2curl4_tet_o1 0
q n2_pt64 FIATB200_MMA_PT=64 n2curl4_tet_o1 0
q n2_nocontr FIATB200_MMA_SKIP=2 n2curl4_tet_o1 0
q n2_nothing FIATB200_MMA_SKIP=3 n2curl4_tet_o1 0
cat gpurun_out/r02_b_exp.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_b_bench.json 2> gpurun_out/r02_b_bench.err
tail -c 3000 gpurun_out/r02_b_bench.json; tail -5 gpurun_out/r02_b_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_b_ref.json 2>&1; tail -c 600 gpurun_out/r02_b_ref.json
