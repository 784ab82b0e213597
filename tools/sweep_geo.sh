#!/bin/bash
# Rebuild the library with different pair-geometry kernel shapes and time the bench's geometry stage.
# Usage (GPU box): bash tools/sweep_geo.sh "MINCTAS RING CHUNK OG" ...
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  export TSPN_NVCC_EXTRA="-DTSPN_GEO_MIN_CTAS=$1 -DTSPN_GEO_RING=$2 -DTSPN_GEO_CHUNK=$3 -DTSPN_GEO_OBJ_GROUP=$4"
  python -m tspn_b200.build --force > /dev/null 2> gpurun_out/sweep_build.err || { echo "cfg $cfg: build failed"; tail -5 gpurun_out/sweep_build.err; continue; }
  timeout 200 python bench.py --no-cpu-baseline --steps 20 2> gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('cfg $cfg: geo %.4f ms  %.0f GB/s  frac %.3f | step %.4f ms value %.2fM e2e %.2fM' % (r['avg_launch_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))
" || tail -3 gpurun_out/sweep.err
done
unset TSPN_NVCC_EXTRA
