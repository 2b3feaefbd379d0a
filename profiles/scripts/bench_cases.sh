#!/bin/sh
# usage: bench_cases.sh out.txt flags case [case...]  -> one throughput line per golden case
OUT="$1"; FLAGS="$2"; shift 2
: > "$OUT"
for c in "$@"; do
  FIATB200_FLAGS=$FLAGS python profiles/scripts/bench_case.py $c 2>&1 | tail -1 >> "$OUT"
done
cat "$OUT"
