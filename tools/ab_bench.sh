#!/bin/bash
# tools/ab_bench.sh lib1.so lib2.so[:ENV=VAL[,ENV=VAL]] ... -- A/B the headline bench over several builds of libpsb (run under gpurun).
# Prints value / e2e / per-phase ms / clocks per build.  Experiments only; never a reported bench number.
mkdir -p gpurun_out
for spec in "$@"; do
  lib=${spec%%:*}
  envs=""
  tag=$lib
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); tag="$lib.$(echo "${spec#*:}" | tr -c 'A-Za-z0-9_\n' '_')"; fi
  env $envs PSB_LIB=$PWD/ps-signature-and-el-passo_b200/$lib python bench.py --steps 2 --warmup 2 --no-cpu-baseline ${AB_ARGS} > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err || tail -3 gpurun_out/ab_$tag.err
  python - "$tag" <<'PY'
import json, sys
lib = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{lib}.json"))
    print(lib, "value=%.0f e2e=%.0f" % (d["value"], d["e2e"]["value"]), {k: round(v, 1) for k, v in d["roofline"]["phase_ms"].items()},
          d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(lib, "FAILED", e)
PY
done
