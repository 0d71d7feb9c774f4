# end-to-end (asynchronous host-buffer products, two in flight) under the pipeline's knobs: bash tools/e2e_async_sweep.sh
run() { tag=$1; shift; env "$@" python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline --no-side-workloads --e2e-steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('$tag', 'async', round(e['value']/1e6,2), 'sync', round(e['sync_value']/1e6,2), e['parity_ok'], e['wire'])"; }
run default X=1
run c32 NTTB200_WIRE_C32=1
run slots8 NTTB200_WIRE_SLOTS=8
run kw512 NTTB200_WIRE_KWORDS=512
run kw2048 NTTB200_WIRE_KWORDS=2048
run w32 NTTB200_WIRE=32
run default2 X=1
