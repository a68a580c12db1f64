#!/bin/bash
# Second evidence pass of round 2 (after the head rewrite + programmatic dependent launch): bench line, layer tables,
# latency, ncu launch list, full captures of the kernels that changed (head) or were not captured before in the
# two-term mode (pw1 / pw2).   gpurun -- 'bash tools/evidence_final.sh'
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/evidence_r2b; mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/bench_1gpu_fp32.json 2> $O/bench_fp32.err; tail -c 300 $O/bench_1gpu_fp32.json
python tools/profile_layers.py 64 500 fp32 > $O/layers_fp32.txt 2>&1
python tools/profile_layers.py 64 500 bf16 > $O/layers_bf16.txt 2>&1
python tools/latency_single.py > $O/latency_single.txt 2>&1
SPARKCODEC_PDL=0 python tools/latency_single.py > $O/latency_single_no_pdl.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_step.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sub none > $O/ncu_launches.log 2>&1
cap() {  # name, kernel regex, skip, count
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name "regex:$2" --launch-skip $3 --launch-count $4 -f -o /tmp/$1 \
    python tools/profile_layers.py 64 500 fp32 > /tmp/$1.log 2>&1
  python tools/ncu_table.py /tmp/$1.ncu-rep > $O/ncu_full_$1.csv 2>> $O/ncu.err
  python tools/ncu_summary.py /tmp/$1.ncu-rep > $O/ncu_full_$1_metrics.txt 2>> $O/ncu.err
}
cap head head_kernel 0 1
cap dwconv_ln dwconv_ln_kernel 4 2
# conv kernel launches of a pass, in order: k7 embed conv, then (pw1, pw2) of the first ConvNeXt block
cap pw1_pw2 conv_gemm_tc_kernel 1 2
ls -la $O
