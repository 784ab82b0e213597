set -x
mkdir -p gpurun_out
for cfg in "8 0" "8 -1" "0 -1" "6 -1"; do
set -- $cfg
TSPN_GEO_RESERVE_SMS=$1 TSPN_SIDE1_PRIORITY=$2 timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline 2> gpurun_out/t_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('reserve $1 side1 prio $2: value %.2fM e2e %.2fM step %.4f geo %.4f alone %.4f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], r['avg_launch_ms'], r['alone']['avg_launch_ms']))
"
done
TSPN_GEO_RESERVE_SMS=8 TSPN_SIDE1_PRIORITY=-1 timeout 300 python tools/trace_step.py --steps 1 2>/dev/null | tail -18
