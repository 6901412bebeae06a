#!/usr/bin/env bash
# Dev helper (run under gpurun): A/B library variants built into vit.cpp_b200/variants/lib_*.so -- each is copied over
# libvitb200.so and benched (device-timed step + per-kernel CUDA-event times) on the same box lease.
cd "$(dirname "$0")/.."
cp vit.cpp_b200/libvitb200.so /tmp/lib_keep.so
for rep in 1 2; do
for f in vit.cpp_b200/variants/lib_*.so; do
  cp "$f" vit.cpp_b200/libvitb200.so
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$f', 'rep$rep', round(d['value']), 'img/s', d['ms_per_step'].__round__(3), 'ms', d['clocks']['sm_mhz'], 'MHz', {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}, 'parity', round(d['parity']['median'],6))"
done
done
cp /tmp/lib_keep.so vit.cpp_b200/libvitb200.so
