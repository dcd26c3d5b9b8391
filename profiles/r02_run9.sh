set -x
mkdir -p gpurun_out/s4
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_hidden128.py tests/test_gpu_encoders.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/s4/pytest.txt
cat gpurun_out/s4/pytest.txt
timeout 300 python bench.py --steps 100 > gpurun_out/s4/bench_fp32.json 2> gpurun_out/s4/bench_fp32.err
timeout 300 python bench.py --steps 100 --dtype bf16 --no-cpu-baseline > gpurun_out/s4/bench_bf16.json 2> gpurun_out/s4/bench_bf16.err
bash profiles/r02_evidence_ncu.sh
