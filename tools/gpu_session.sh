#!/bin/bash
# tools/gpu_session.sh TAG [lib specs...] -- one gpurun call of the round: GPU parity tests, micro-benchmarks, an A/B of
# the given library builds (tools/ab_bench.sh syntax), a launch list and one ncu --set full capture of a full wave.
# Everything lands in gpurun_out/TAG_*.  Sections can be skipped with SKIP="tests micro ab ncu".
TAG=$1; shift
mkdir -p gpurun_out
skip() { [[ " $SKIP " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
if ! skip tests; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
fi
if ! skip micro; then
  timeout 300 python tools/microbench.py > /dev/null 2> gpurun_out/${TAG}_microbench.err && mv gpurun_out/microbench.json gpurun_out/${TAG}_microbench.json
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_microbench.json"))
best = {}
for r in d["mad"]:
    best[r["probe"][:60]] = max(best.get(r["probe"][:60], 0), r["ops_per_clk_per_sm_at_1965MHz"])
for k, v in best.items(): print("  %-62s %.1f /clk/SM" % (k, v))
PY
fi
if ! skip ab; then
  bash tools/ab_bench.sh "$@" | tee gpurun_out/${TAG}_ab.txt
fi
if ! skip ncu; then
  NCU_LIB=${NCU_LIB:-libpsb.so}
  M=sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fmalite.sum
  PSB_LIB=$PWD/ps-signature-and-el-passo_b200/$NCU_LIB timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:k_verify -s 3 -c 3 -f \
      -o gpurun_out/${TAG}_prof python tools/prof_verify.py ${NCU_LANES:-75776} 20 2 > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
  tail -2 gpurun_out/${TAG}_ncu.log
fi
