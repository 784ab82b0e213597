set -x
O=gpurun_out
timeout 120 python -m pytest tests/test_gpu_tensor.py -m gpu -q -x -k span_head_tensor > $O/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r2t_pytest.log
timeout 120 python tools/bench_span_head.py 256 1024 300 > $O/r2t_span_head.jsonl 2> $O/r2t_span_head.err; echo rc=$?
timeout 120 python tools/bench_span_head.py 1024 1024 2000 >> $O/r2t_span_head.jsonl 2>> $O/r2t_span_head.err; echo rc=$?
TSPN_SPAN_HEAD_ONE_CTA=1 timeout 120 python tools/bench_span_head.py 256 1024 300 >> $O/r2t_span_head.jsonl 2>> $O/r2t_span_head.err; echo rc=$?
TSPN_SPAN_HEAD_ONE_CTA=1 timeout 120 python tools/bench_span_head.py 1024 1024 2000 >> $O/r2t_span_head.jsonl 2>> $O/r2t_span_head.err; echo rc=$?
cat $O/r2t_span_head.jsonl; tail -5 $O/r2t_span_head.err
nvidia-smi --query-gpu=name,clocks.sm --format=csv
