#!/bin/bash
mkdir -p gpurun_out/wire
W=${W:-c2}
run() {  # name, env...
  name=$1; shift
  env "$@" python bench.py --workload $W --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/wire/${W}_$name.json 2> gpurun_out/wire/${W}_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/wire/${W}_$name.json"))
print("$W $name", round(d["e2e"]["value"]/1e6,2), "M polymul/s", d["e2e"]["wire"], d["e2e"]["parity_ok"])
PY
}
run default A=1
run noramp NTTB200_WIRE_RAMP=0
run nta NTTB200_WIRE_NTA=1
run kw512 NTTB200_WIRE_KWORDS=512
run kw512_nta NTTB200_WIRE_KWORDS=512 NTTB200_WIRE_NTA=1
run kw512_s4 NTTB200_WIRE_KWORDS=512 NTTB200_WIRE_SLOTS=4
run kw2048 NTTB200_WIRE_KWORDS=2048
run s4 NTTB200_WIRE_SLOTS=4
run s3 NTTB200_WIRE_SLOTS=3
run t15 NTTB200_HOST_THREADS=15
run t14 NTTB200_HOST_THREADS=14
W=c3 run default A=1
W=c3 run w32 NTTB200_WIRE=32
