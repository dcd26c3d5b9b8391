set -x
mkdir -p gpurun_out/s3
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s3/pytest.txt
timeout 300 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/s3/bench_b.json 2> gpurun_out/s3/bench_b.err
cat gpurun_out/s3/pytest.txt
