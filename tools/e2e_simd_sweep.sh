# host-side narrowing / widening variants (NTTB200_WIRE_SIMD) on the GPU box: the host ceiling alone,
# then the end-to-end call of bench.py
gcc -O2 -pthread -I ntt-based-polynomial-multiplier-fpga_b200/csrc -o /tmp/hostwire_bench tools/hostwire_bench.c ntt-based-polynomial-multiplier-fpga_b200/csrc/hostwire.c
nproc; grep -m1 "model name" /proc/cpuinfo
for v in avx2 avx2nt avx512 avx512nt; do
  echo "== $v"; NTTB200_WIRE_SIMD=$v /tmp/hostwire_bench | sed -n 3,6p
  for w in c2 c3; do
  NTTB200_WIRE_SIMD=$v timeout 300 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline --no-side-workloads --e2e-steps 20 > gpurun_out/e2e_$v_$w.json 2> gpurun_out/e2e_$v_$w.err
  python -c "
import json
d=json.load(open('gpurun_out/e2e_$v_$w.json')); print('$v $w e2e', round(d['e2e']['value']/1e6,2), 'M/s', d['e2e']['wire'], d['e2e']['parity_ok'])"
  done
done
