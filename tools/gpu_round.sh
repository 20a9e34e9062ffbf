#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list + one full capture of the top kernel.
# usage (under gpurun): bash tools/gpu_round.sh [top-kernel-regex]
set -u
mkdir -p gpurun_out
REGEX=${1:-vt_slice_kernel}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
python bench.py --workload modes --steps 3 --warmup 2 > gpurun_out/modes.json 2> gpurun_out/modes.err; echo "modes rc=$?"
cat gpurun_out/modes.json
python bench.py --workload project --steps 5 --warmup 2 > gpurun_out/project.json 2> gpurun_out/project.err; echo "project rc=$?"
python tools/prefilter_probe.py 250 256 512 > gpurun_out/prefilter_probe.log 2>&1; echo "prefilter probe rc=$?"
python tools/sweep_probe.py 256 > gpurun_out/sweep_probe.log 2>&1; echo "sweep probe rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --batch 2 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$REGEX -s 4 -c 2 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
