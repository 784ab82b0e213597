set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1s4_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r1s4_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r1s4_bench.json 2> gpurun_out/r1s4_bench.err; echo "bench rc=$?"
cat gpurun_out/r1s4_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1s4_bench_ref.json 2>&1; cat gpurun_out/r1s4_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1s4_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eager > gpurun_out/r1s4_b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_geo_kernel|assemble_kernel|span_proposals|predicate_tc|topk_kernel' -s 10 -c 8 -o gpurun_out/r1s4_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eager > gpurun_out/r1s4_ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out
