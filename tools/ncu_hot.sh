# ncu --set full of the side-branch kernels on the bench workload (eager launches so that every kernel is a launch)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'span_select|survivor_rows|scores_topk|ppn_embed|pair_top_predicates' -s 20 -c 6 -o gpurun_out/${1:-r2}_side python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eager > gpurun_out/${1:-r2}_side_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/${1:-r2}_side.ncu-rep gpurun_out/${1:-r2}_side_summary.md > /dev/null 2>&1
