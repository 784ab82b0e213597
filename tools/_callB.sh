set -x
N=8; O=gpurun_out; TAG=r2s
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_single_n$N.json 2> $O/${TAG}_single_n$N.err; echo "rc=$?"; tail -2 $O/${TAG}_single_n$N.err
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --geo-layout windowed > $O/${TAG}_single_n${N}_windowed.json 2> /dev/null; echo "rc=$?"
timeout 400 $TR bench.py --gpus $N --workload vidor_val --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_val_n$N.json 2> $O/${TAG}_val_n$N.err; echo "rc=$?"; tail -2 $O/${TAG}_val_n$N.err
cat $O/${TAG}_single_n$N.json $O/${TAG}_single_n${N}_windowed.json $O/${TAG}_val_n$N.json | cut -c1-900
