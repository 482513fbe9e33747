#!/bin/bash
# Refresh every number DESIGN.md quotes in ONE gpurun call (one B200, about 4 GPU-minutes):
#
#   gpurun --timeout 420 -- 'bash profiles/refresh.sh'
#
# Outputs land in gpurun_out/ (scratch); copy what should be judged into profiles/rNN_*.  Every ncu pass runs only after
# the same command has exited 0 without ncu; numbers printed under ncu are never bench values.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
out=gpurun_out
mkdir -p "$out"
py=python
slim="--no-cpu-baseline --no-e2e --no-aux"

timeout 300 $py -m pytest tests -q -x -m gpu > "$out/pytest_gpu.log" 2>&1
echo "pytest -m gpu: rc=$? $(tail -1 "$out/pytest_gpu.log")"

timeout 240 $py bench.py > "$out/bench_wn.json" 2> "$out/bench_wn.err" || { echo "bench failed"; tail -5 "$out/bench_wn.err"; exit 1; }
timeout 240 $py bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_ref.json" 2> "$out/bench_ref.err"
timeout 120 $py bench.py --workload fb15k237 --no-aux --no-cpu-baseline > "$out/bench_fb.json" 2> "$out/bench_fb.err"
cut -c1-400 "$out/bench_wn.json"

# launch list of one eager layer step (cold-cache, serialised: compare SHARES with the graph-captured step)
if timeout 120 $py bench.py --steps 2 --warmup 1 --no-graph $slim > /dev/null 2>&1; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file "$out/launches.csv" \
      $py bench.py --steps 2 --warmup 1 --no-graph $slim > "$out/ncu_launch.log" 2>&1
  $py profiles/summarize_launches.py "$out/launches.csv" > "$out/launches.md" 2>/dev/null || true
  # full capture of the aggregation / tail kernels of one step (source-level stalls, DRAM traffic)
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"agg_lean|tail_fwd|tail_bwd_apply|rows_reduce|gemm_tf32x3|gemm_tn_tc" \
      -s 12 -c 12 -o "$out/layer_kernels" -f $py bench.py --steps 2 --warmup 1 --no-graph $slim > "$out/ncu_full.log" 2>&1
fi
ls -la "$out" | tail -20
