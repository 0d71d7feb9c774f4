# usage: bash tools/small_sweep.sh <tag> [ENV=VAL ...]  -- c2, c3, c4 device-resident values of one build/knob set
tag=$1; shift
for w in c2 c3 c4; do
  env "$@" timeout 300 python bench.py --workload $w --steps 200 --warmup 10 --no-cpu-baseline --no-side-workloads --e2e-steps 3 > gpurun_out/ss_${tag}_$w.json 2> gpurun_out/ss_${tag}_$w.err
  python -c "
import json
d=json.load(open('gpurun_out/ss_${tag}_$w.json')); print('$tag $w', round(d['value']/1e6,1), 'M/s  sustained', round(d['sustained']['value']/1e6,1), ' two-streams', round(d['two_streams']['value']/1e6,1) if d['two_streams'] else None, ' e2e', round(d['e2e']['value']/1e6,2), d['parity_ok'], d['e2e']['parity_ok'])"
done
