set -x
mkdir -p gpurun_out/s3
timeout 300 python bench.py --steps 60 --no-cpu-baseline > gpurun_out/s3/bench_pf1.json 2> gpurun_out/s3/bench_pf1.err
SCGIB_PRE_PF=0 timeout 300 python bench.py --steps 60 --no-cpu-baseline > gpurun_out/s3/bench_pf0.json 2> gpurun_out/s3/bench_pf0.err
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/s3/pytest_pf.txt
cat gpurun_out/s3/pytest_pf.txt
