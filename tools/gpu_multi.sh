# N-GPU checks: NCCL parity test, sharded and replica benches, what limits the host-facing loop.
# usage: bash tools/gpu_multi.sh <tag> <N>
set -x
TAG=${1:-m}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1; nproc >> gpurun_out/${TAG}_topo.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/${TAG}_topo.txt
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5; fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_single_n$N.json 2> gpurun_out/${TAG}_bench_single_n$N.err; echo "rc=$?"; tail -2 gpurun_out/${TAG}_bench_single_n$N.err
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --collective peer > gpurun_out/${TAG}_bench_single_n${N}_peer.json 2> gpurun_out/${TAG}_bench_single_n${N}_peer.err; echo "rc=$?"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --collective peer --geo-layout windowed > gpurun_out/${TAG}_bench_single_n${N}_peer_windowed.json 2> /dev/null; echo "rc=$?"
timeout 900 $TR bench.py --gpus $N --workload vidor_val --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_val_n$N.json 2> gpurun_out/${TAG}_bench_val_n$N.err; echo "rc=$?"; tail -2 gpurun_out/${TAG}_bench_val_n$N.err
timeout 900 $TR tools/e2e_limiter.py --steps 200 > gpurun_out/${TAG}_limiter_n$N.jsonl 2> gpurun_out/${TAG}_limiter_n$N.err; echo "rc=$?"; tail -2 gpurun_out/${TAG}_limiter_n$N.err

cat gpurun_out/${TAG}_bench_single_n$N.json gpurun_out/${TAG}_bench_single_n${N}_peer.json gpurun_out/${TAG}_bench_single_n${N}_peer_windowed.json gpurun_out/${TAG}_bench_val_n$N.json gpurun_out/${TAG}_limiter_n$N.jsonl | cut -c1-700
