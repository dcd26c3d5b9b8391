set -x
mkdir -p gpurun_out/s3
timeout 600 python -m pytest tests/test_gpu_encoders.py tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -12 > gpurun_out/s3/pytest_enc.txt
cat gpurun_out/s3/pytest_enc.txt
for c in 3 0 6; do SCGIB_EGO_CTAS=$c timeout 300 python bench.py --steps 60 --no-cpu-baseline > gpurun_out/s3/bench_egoc$c.json 2> gpurun_out/s3/bench_egoc$c.err; done
