# c4 (n=1024) with the signed kernel and variant builds: bash tools/c4_signed_variants.sh "" _cs0 _cs2 ...
L=$PWD/ntt-based-polynomial-multiplier-fpga_b200
for t in "$@"; do
  NTTB200_PLANT_SIGNED=1 NTTB200_LIB=$L/libnttb200$t.so timeout 300 python bench.py --workload c4 --steps 100 --warmup 10 --no-cpu-baseline --no-side-workloads --e2e-steps 2 > gpurun_out/c4s$t.json 2> gpurun_out/c4s$t.err
  python -c "
import json
d=json.load(open('gpurun_out/c4s$t.json')); print('c4 signed lib$t', round(d['value']/1e6,1), 'sustained', round(d['sustained']['value']/1e6,1), d['sustained']['clocks']['sm_mhz'], d['sustained']['clocks']['reasons'], d['parity_ok'])"
done
