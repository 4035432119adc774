#!/bin/bash
# A/B of engine switches on one box (every run bounded by `timeout`): usage  bash tools/ab_bench.sh "NAME ENV=.." ...
run() { # name, env...
  name=$1; shift
  timeout 240 env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err || tail -3 gpurun_out/ab_$name.err
}
rm -f gpurun_out/ab_*.json
for spec in "$@"; do run $spec; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"]), round(d["ms_per_step"],3), "cycles(M)", round(d["ms_per_step"]*d["clocks"]["sm_mhz"]/1e3,2), d["gpu_launches"]/d["steps"], "b1", round(d["latency_b1_ms"],3), d["clocks"]["sm_mhz"], "e2e", round(d["e2e"]["value"]), {k:round(v["ms_serialised"],3) for k,v in d["kernel_classes"].items()})
    except Exception as ex: print(f, ex)
PY
