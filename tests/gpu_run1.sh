set -x
mkdir -p gpurun_out/s3
timeout 600 python -m pytest tests/test_gpu_encoders.py tests/test_dropin_gpu.py -m gpu -q 2>&1 | tail -40 > gpurun_out/s3/pytest_enc.txt
cat gpurun_out/s3/pytest_enc.txt
timeout 300 python bench.py --encoder GraphSAGE --steps 30 > gpurun_out/s3/bench_sage.json 2> gpurun_out/s3/bench_sage.err || tail -5 gpurun_out/s3/bench_sage.err
timeout 300 python bench.py --encoder GCN --steps 30 > gpurun_out/s3/bench_gcn.json 2> gpurun_out/s3/bench_gcn.err || tail -5 gpurun_out/s3/bench_gcn.err
cat gpurun_out/s3/bench_sage.json gpurun_out/s3/bench_gcn.json | cut -c 1-400
