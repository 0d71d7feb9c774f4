run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --workload c5 --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/c5s_$name.json 2> gpurun_out/c5s_$name.err
  python -c "
import json
d=json.load(open('gpurun_out/c5s_$name.json')); print('$name', round(d['value']/1e6,3), 'M/s sustained', round(d['sustained']['value']/1e6,3), d['parity_ok'], d['sustained']['clocks']['reasons'])"
}
L=$PWD/ntt-based-polynomial-multiplier-fpga_b200
run base A=1
run row4 NTTB200_LIB=$L/libnttb200_row4.so
run row4_l4 NTTB200_LIB=$L/libnttb200_row4.so NTTB200_LARGE_LANES=4
run l2 NTTB200_LARGE_LANES=2
run l4 NTTB200_LARGE_LANES=4
run l6 NTTB200_LARGE_LANES=6
run s16 NTTB200_LARGE_SCRATCH_MB=16
run s64 NTTB200_LARGE_SCRATCH_MB=64
run s16l6 NTTB200_LARGE_SCRATCH_MB=16 NTTB200_LARGE_LANES=6
run row4_s16l6 NTTB200_LIB=$L/libnttb200_row4.so NTTB200_LARGE_SCRATCH_MB=16 NTTB200_LARGE_LANES=6
