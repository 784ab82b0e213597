# One GPU call: parity tests, bench (both arms), launch list, full ncu capture of the hot kernels.
# usage: bash tools/gpu_round_check.sh <tag> [full]
set -x
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
if [ "$2" = "full" ]; then
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>&1; cat gpurun_out/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eager > gpurun_out/${TAG}_b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_geo_kernel|survivor_rows|predicate_tc|topk_kernel|video_top_triplets|tracklet_rows' -s 12 -c 10 -o gpurun_out/${TAG}_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eager > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/trace_step.py --steps 2 > gpurun_out/${TAG}_timeline.txt 2> gpurun_out/${TAG}_timeline.err
python tools/ncu_summary.py gpurun_out/${TAG}_prof.ncu-rep gpurun_out/${TAG}_ncu_summary.md > /dev/null 2>&1
fi
ls -la gpurun_out
