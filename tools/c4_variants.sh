# c4 (n=1024, q=12289, batch 2^18) with variant builds of the library: bash tools/c4_variants.sh "" _x3 _x4 ...
L=$PWD/ntt-based-polynomial-multiplier-fpga_b200
for t in "$@"; do
  NTTB200_LIB=$L/libnttb200$t.so timeout 300 python bench.py --workload c4 --steps 100 --warmup 10 --no-cpu-baseline --no-side-workloads --e2e-steps 2 > gpurun_out/c4v$t.json 2> gpurun_out/c4v$t.err
  python -c "
import json
d=json.load(open('gpurun_out/c4v$t.json')); print('c4 lib$t', round(d['value']/1e6,1), 'sustained', round(d['sustained']['value']/1e6,1), d['sustained']['clocks']['sm_mhz'], d['sustained']['clocks']['reasons'], d['parity_ok'])"
done
