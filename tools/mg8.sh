# 8-GPU measurements (gpurun --gpus 8): BASELINE configs[3] (C4), C2 weak scaling, C3 mesh config
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --config c4 --steps 3 --warmup 2 > gpurun_out/bench_g${N}_c4.log 2>&1; tail -1 gpurun_out/bench_g${N}_c4.log | cut -c1-2200
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_g${N}_c2.log 2>&1; tail -1 gpurun_out/bench_g${N}_c2.log | cut -c1-1200
timeout 600 $TR bench.py --gpus $N --config mesh --steps 3 --warmup 2 > gpurun_out/bench_g${N}_mesh.log 2>&1; tail -1 gpurun_out/bench_g${N}_mesh.log | cut -c1-1200
