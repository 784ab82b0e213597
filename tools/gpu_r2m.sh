# round 2, pass m: 8 GPUs - e2e with the NCCL all-gather vs the peer-memory exchange, windowed layout, sharded config
set -x
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --collective nccl > gpurun_out/r2m_single_n${N}_nccl.json 2> gpurun_out/r2m_single_n${N}_nccl.err; echo "rc=$?"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --collective peer > gpurun_out/r2m_single_n${N}_peer.json 2> gpurun_out/r2m_single_n${N}_peer.err; echo "rc=$?"; tail -3 gpurun_out/r2m_single_n${N}_peer.err
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 5 --collective nccl > gpurun_out/r2m_single200_n${N}_nccl.json 2> /dev/null; echo "rc=$?"
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 5 --collective peer > gpurun_out/r2m_single200_n${N}_peer.json 2> /dev/null; echo "rc=$?"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --collective peer --geo-layout windowed > gpurun_out/r2m_single_n${N}_peer_windowed.json 2> gpurun_out/r2m_single_n${N}_peer_windowed.err; echo "rc=$?"
timeout 300 $TR bench.py --gpus $N --workload vidor_val --steps 5 --warmup 3 > gpurun_out/r2m_val_n${N}.json 2> gpurun_out/r2m_val_n${N}.err; echo "rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2m_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("%-50s n=%d value %.1fM e2e %.1fM ms %.4f coll %s layout %s" % (f, d["n_gpus"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d["config"].get("collective"), d["config"].get("geo_layout")))
    except Exception as e: print(f, e)
PY
