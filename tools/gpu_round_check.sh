# One GPU call: parity tests, bench (all workloads), launch list, timeline.
# usage: bash tools/gpu_round_check.sh <tag> [tests|bench|full]
set -x
TAG=${1:-run}
WHAT=${2:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
if [ "$WHAT" != "bench" ]; then
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/${TAG}_pytest.log
fi
if [ "$WHAT" != "tests" ]; then
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
for W in vidvrd_single vidvrd_test vidor_val stress; do
timeout 900 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${W}.json 2> gpurun_out/${TAG}_bench_${W}.err; echo "bench $W rc=$?"
cat gpurun_out/${TAG}_bench_${W}.json; tail -5 gpurun_out/${TAG}_bench_${W}.err
done
timeout 300 python tools/trace_step.py --steps 2 > gpurun_out/${TAG}_timeline.txt 2> gpurun_out/${TAG}_timeline.err
timeout 300 python tools/trace_step.py --steps 2 --span-proposals 0 > gpurun_out/${TAG}_timeline_nonms.txt 2>> gpurun_out/${TAG}_timeline.err
fi
if [ "$WHAT" = "full" ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cat gpurun_out/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eager > gpurun_out/${TAG}_b_ncu.log 2>&1
fi
ls -la gpurun_out | head -40
