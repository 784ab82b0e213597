O=gpurun_out
timeout 120 python -m pytest tests/test_gpu_tensor.py -m gpu -q -x -k span_head_tensor > $O/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2v_pytest.log
rm -f $O/r2v_span_head.jsonl
for ONE in 0 1; do for D in 0 1 2 3; do
  if [ $ONE = 1 ]; then export TSPN_SPAN_HEAD_ONE_CTA=1; else unset TSPN_SPAN_HEAD_ONE_CTA; fi
  echo "one_cta=$ONE dbg=$D" >> $O/r2v_span_head.jsonl
  TSPN_SPAN_HEAD_DEBUG=$D timeout 120 python tools/bench_span_head.py 256 1024 300 >> $O/r2v_span_head.jsonl 2>> $O/r2v_span_head.err
done; done
for ONE in 0 1; do
  if [ $ONE = 1 ]; then export TSPN_SPAN_HEAD_ONE_CTA=1; else unset TSPN_SPAN_HEAD_ONE_CTA; fi
  echo "one_cta=$ONE big" >> $O/r2v_span_head.jsonl
  timeout 120 python tools/bench_span_head.py 1024 1024 2000 >> $O/r2v_span_head.jsonl 2>> $O/r2v_span_head.err
done
unset TSPN_SPAN_HEAD_ONE_CTA
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2v_launches_pair.csv python tools/bench_span_head.py 256 1024 300 > /dev/null 2>&1
TSPN_SPAN_HEAD_ONE_CTA=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2v_launches_one.csv python tools/bench_span_head.py 256 1024 300 > /dev/null 2>&1
cut -c1-130 $O/r2v_span_head.jsonl
