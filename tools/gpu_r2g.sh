# round 2, pass g: N GPUs - NCCL / peer-exchange parity tests, e2e with either exchange
# usage: bash tools/gpu_r2g.sh <N>
set -x
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -30; fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
for C in nccl peer; do
  timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --collective $C > gpurun_out/r2g_single_n${N}_$C.json 2> gpurun_out/r2g_single_n${N}_$C.err; echo "rc=$?"; tail -3 gpurun_out/r2g_single_n${N}_$C.err
  timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 5 --collective $C > gpurun_out/r2g_single200_n${N}_$C.json 2> gpurun_out/r2g_single200_n${N}_$C.err; echo "rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2g_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("%-46s n=%d value %.1fM e2e %.1fM ms %.4f coll %s" % (f, d["n_gpus"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d["config"].get("collective")))
    except Exception as e: print(f, e)
PY
