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
slim="--no-cpu-baseline --no-e2e --no-aux --no-extras --no-parity"

timeout 300 $py -m pytest tests -q -x -m gpu > "$out/pytest_gpu.log" 2>&1
echo "pytest -m gpu: rc=$? $(tail -1 "$out/pytest_gpu.log")"

# N = 1: the Wikidata5M-shape headline with the WN18RR / FB15k-237 entries, e2e, aux and cpu_baseline in the same line
timeout 600 $py bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err" || { echo "bench failed"; tail -5 "$out/bench_n1.err"; exit 1; }
timeout 400 $py bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_ref.json" 2> "$out/bench_ref.err"
cut -c1-400 "$out/bench_n1.json"
# N > 1 (gpurun --gpus N): python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N

# launch list of one eager layer step (cold-cache, serialised: compare SHARES with the graph-captured step)
if timeout 120 $py bench.py --steps 2 --warmup 1 --no-graph $slim > /dev/null 2>&1; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file "$out/launches.csv" \
      $py bench.py --steps 2 --warmup 1 --no-graph $slim > "$out/ncu_launch.log" 2>&1
  $py profiles/summarize_launches.py "$out/launches.csv" > "$out/launches.md" 2>/dev/null || true
  # full capture of the aggregation / tail kernels of one step (source-level stalls, DRAM traffic)
  # (about 12 GPU-minutes at the Wikidata5M shape: ncu saves / restores the 16 GB tables for every replay pass)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"agg_lean|tail_|rows_reduce|gemm_tf32x3|gemm_tn_tc_kernel" \
      -s 54 -c 18 -o "$out/layer_kernels" -f $py bench.py --steps 1 --warmup 3 --no-graph $slim > "$out/ncu_full.log" 2>&1
  # then here: ncu -i gpurun_out/layer_kernels.ncu-rep --page raw --csv > /tmp/raw.csv
  #            python profiles/summarize_traffic.py /tmp/raw.csv wikidata5m --json profiles/r02_ncu_traffic.json
fi
ls -la "$out" | tail -20
