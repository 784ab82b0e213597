O=gpurun_out
rm -f $O/r2u_span_head.jsonl
for ONE in 0 1; do for D in 0 1 2 3; do
  if [ $ONE = 1 ]; then export TSPN_SPAN_HEAD_ONE_CTA=1; else unset TSPN_SPAN_HEAD_ONE_CTA; fi
  echo "one_cta=$ONE dbg=$D" >> $O/r2u_span_head.jsonl
  TSPN_SPAN_HEAD_DEBUG=$D timeout 120 python tools/bench_span_head.py 256 1024 300 >> $O/r2u_span_head.jsonl 2>> $O/r2u_span_head.err
done; done
unset TSPN_SPAN_HEAD_ONE_CTA
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2u_launches_pair.csv python tools/bench_span_head.py 256 1024 300 > /dev/null 2>&1
TSPN_SPAN_HEAD_ONE_CTA=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2u_launches_one.csv python tools/bench_span_head.py 256 1024 300 > /dev/null 2>&1
cat $O/r2u_span_head.jsonl
tail -12 $O/r2u_launches_pair.csv | cut -c1-200
tail -6 $O/r2u_launches_one.csv | cut -c1-200
