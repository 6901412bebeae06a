for m in 0 4 8; do echo "LN_DBG=$m"; VITB200_LN_DBG=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --sustained-seconds 0.2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(round(d['value']), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}, d['parity']['median'])"; done
