#!/usr/bin/env bash
# Dev helper (run under gpurun): A/B of runtime knobs on one box lease.  usage: tools/ab_env.sh "VAR=0" "VAR=1" ...  (each twice)
cd "$(dirname "$0")/.."
for rep in 1 2; do
for kv in "$@"; do
  env $kv timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$kv', 'rep$rep', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], 'MHz', {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}, 'parity', round(d['parity']['median'],6), round(d['parity']['max'],6), d['parity']['top5_lists_identical'])"
done
done
