run() { tag=$1; shift; env "$@" python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline --no-side-workloads --e2e-steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('$tag', 'async', round(e['value']/1e6,2), 'sync', round(e['sync_value']/1e6,2), e['parity_ok'], e['wire'])"; }
run ahead2 NTTB200_WIRE_AHEAD=2
run ahead3 NTTB200_WIRE_AHEAD=3
run ahead4 NTTB200_WIRE_AHEAD=4
run ahead5 NTTB200_WIRE_AHEAD=5
run s8_ahead5 NTTB200_WIRE_SLOTS=8 NTTB200_WIRE_AHEAD=5
run s8_ahead6 NTTB200_WIRE_SLOTS=8 NTTB200_WIRE_AHEAD=6
run kw512_s10_ahead6 NTTB200_WIRE_KWORDS=512 NTTB200_WIRE_SLOTS=10 NTTB200_WIRE_AHEAD=6
run kw512_s10_ahead8 NTTB200_WIRE_KWORDS=512 NTTB200_WIRE_SLOTS=10 NTTB200_WIRE_AHEAD=8
