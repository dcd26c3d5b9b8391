set -x
mkdir -p gpurun_out/s5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/s5/bench_8gpu_b4096.json 2> gpurun_out/s5/bench_8gpu_b4096.err || tail -5 gpurun_out/s5/bench_8gpu_b4096.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --batch 8192 --steps 50 --warmup 10 > gpurun_out/s5/bench_8gpu_b8192.json 2> gpurun_out/s5/bench_8gpu_b8192.err || tail -5 gpurun_out/s5/bench_8gpu_b8192.err
cut -c 1-260 gpurun_out/s5/bench_8gpu_b4096.json gpurun_out/s5/bench_8gpu_b8192.json
