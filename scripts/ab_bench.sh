#!/bin/bash
# usage: scripts/ab_bench.sh <variantA.so|-> <variantB.so|-> ...   ("-" = the in-tree library); alternates A,B,A,B on one box
for rep in 1 2; do
  for v in "$@"; do
    if [ "$v" = "-" ]; then unset B200POA_LIB; else export B200POA_LIB="$v"; fi
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
