# Round-2 ncu full captures (gpurun): GIN kernels fp32 / bf16 and the other kernels of the step.  The reports (30-40 MB each)
# exceed gpurun's 64 MiB copy-back cap, so they are exported to raw CSV on the box and deleted; profiles/make_traffic.py reads the CSVs.
set -x
mkdir -p gpurun_out/ev
cap() { tag=$1; filt=$2; skip=$3; cnt=$4; shift 4; python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ev/plain_$tag.log 2>&1 && ncu --set full --clock-control none -k regex:"$filt" -s $skip -c $cnt -f -o /tmp/r02_prof_$tag python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ev/ncu_full_$tag.log 2>&1; ncu -i /tmp/r02_prof_$tag.ncu-rep --page raw --csv > gpurun_out/r02_prof_$tag.csv 2> /dev/null; rm -f /tmp/r02_prof_$tag.ncu-rep; }
cap fp32_gin gin_ 24 12
cap bf16_gin gin_ 24 12 --dtype bf16
cap fp32_other "contrastive|head_fwd_tc|recon|graph_gate|gate_lin|input_proj|ego_pool|fwd_prep|reduce_partials" 20 16
ls -la gpurun_out/*.csv
