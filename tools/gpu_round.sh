#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (headline + tables), reference arm, ncu launch list + one full capture
# of the headline kernel.   usage (under gpurun): bash tools/gpu_round.sh [top-kernel-regex]
set -u
mkdir -p gpurun_out
REGEX=${1:-vt_z4_kernel}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 2000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cut -c1-400 gpurun_out/bench_ref.json
if [ "${NCU:-0}" = "1" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
  $CMD > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "ncu list rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:$REGEX -s 6 -c 2 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
ls -la gpurun_out | tail -20
