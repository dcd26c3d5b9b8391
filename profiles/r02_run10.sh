set -x
mkdir -p gpurun_out/s5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/s5/bench_2gpu.json 2> gpurun_out/s5/bench_2gpu.err || tail -5 gpurun_out/s5/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 5 --warmup 3 > gpurun_out/s5/bench_2gpu_ref.json 2> gpurun_out/s5/bench_2gpu_ref.err || tail -5 gpurun_out/s5/bench_2gpu_ref.err
cut -c 1-300 gpurun_out/s5/bench_2gpu.json gpurun_out/s5/bench_2gpu_ref.json
