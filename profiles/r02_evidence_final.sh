# Round-2 evidence of the FINAL build (gpurun): GPU tests, smoke, every kept bench line, ncu launch list.
#   gpurun --timeout 1500 -- 'bash profiles/r02_evidence_final.sh'   -> gpurun_out/ev2/*
set -x
mkdir -p gpurun_out/ev2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/ev2/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ev2/smoke.txt 2>&1
b() { tag=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/ev2/bench_$tag.json 2> gpurun_out/ev2/bench_$tag.err || tail -3 gpurun_out/ev2/bench_$tag.err; }
b fp32 --steps 100
b bf16 --dtype bf16 --steps 100 --no-cpu-baseline
b reference --impl reference --steps 20 --warmup 3
b k2 --k 2 --no-cpu-baseline --steps 50
b k3 --k 3 --no-cpu-baseline --steps 50
b k2_bf16 --k 2 --dtype bf16 --no-cpu-baseline --steps 50
b k3_bf16 --k 3 --dtype bf16 --no-cpu-baseline --steps 50
b b8192 --batch 8192 --no-cpu-baseline --steps 50
b b8192_bf16 --batch 8192 --dtype bf16 --no-cpu-baseline --steps 50
b b128 --batch 128 --no-cpu-baseline --steps 200
b b128_bf16 --batch 128 --dtype bf16 --no-cpu-baseline --steps 200
b dims128 --dims 128 --no-cpu-baseline --steps 30
b dims128_bf16 --dims 128 --dtype bf16 --no-cpu-baseline --steps 30
b finetune_pep_dims128 --workload finetune --shape peptides --batch 1024 --dims 128 --steps 30 --no-cpu-baseline
b finetune_pep_dims128_bf16 --workload finetune --shape peptides --batch 1024 --dims 128 --dtype bf16 --steps 30 --no-cpu-baseline
b finetune_pep_dims64 --workload finetune --shape peptides --batch 1024 --no-cpu-baseline --steps 30
b finetune_pcqm --workload finetune --shape pcqm --batch 4096 --no-cpu-baseline --steps 50
b logm_k1 --recons_type logM --no-cpu-baseline --steps 50
b encoder_graphsage --encoder GraphSAGE --steps 30
b encoder_gcn --encoder GCN --steps 30
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ev2/plain_fp32.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 200 --csv --log-file gpurun_out/ev2/r02_ncu_launches_fp32_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ev2/ncu_launches.log 2>&1
cat gpurun_out/ev2/pytest_gpu.txt gpurun_out/ev2/smoke.txt
