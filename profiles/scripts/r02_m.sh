#!/bin/sh
# Round-2 step m (B200, one GPU): GPU tests, the reference's whole unit-test suite judged on the device drop-in,
# then the default bench (roofline.traffic from profiles/traffic.json).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_m_tests.txt 2>&1
tail -5 gpurun_out/r02_m_tests.txt
rm -f /tmp/dropin_stats.jsonl
( cd /tmp && PYTHONPATH=/root/repo/oracle/_ref:/root/repo FIATB200_DROPIN=device FIATB200_DROPIN_STATS=/tmp/dropin_stats.jsonl \
  timeout 2400 python -m pytest -p oracle.dropin_plugin -q -p no:cacheprovider -c /dev/null --rootdir /tmp -n 3 \
  -k "not macro_gem and not macro_sympy" --durations=25 /root/repo/oracle/_ref/ref_tests \
  --ignore=/root/repo/oracle/_ref/ref_tests/test_precision.py ) > gpurun_out/r02_m_reference_suite_device.txt 2>&1
tail -45 gpurun_out/r02_m_reference_suite_device.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_m_bench_default.json 2> gpurun_out/r02_m_bench_default.err
tail -c 1500 gpurun_out/r02_m_bench_default.json
