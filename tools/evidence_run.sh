#!/bin/bash
# Round evidence on the GPU box: tests, bench lines, per-layer tables, ncu launch list and full captures
# (summarised here; the .ncu-rep files stay on the box).   gpurun -- 'bash tools/evidence_run.sh [quick]'
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/evidence_r2; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/bench_1gpu_fp32.json 2> $O/bench_fp32.err; tail -c 300 $O/bench_1gpu_fp32.json
python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference.json 2> $O/bench_ref.err
python tools/profile_layers.py 64 500 fp32 > $O/layers_fp32.txt 2>&1
python tools/profile_layers.py 64 500 bf16 > $O/layers_bf16.txt 2>&1
SPARKCODEC_FP32_TERMS=3 python tools/profile_layers.py 64 500 fp32 > $O/layers_fp32_three_term.txt 2>&1
SPARKCODEC_FP32_TERMS=3 python bench.py --steps 20 --warmup 5 --sub none --no-cpu-baseline > $O/bench_1gpu_fp32_three_term.json 2> $O/bench_t3.err
[ -x tools/micro/umma_kinds ] && timeout 120 tools/micro/umma_kinds 2>&1 | grep -v illegal > $O/umma_kinds.txt
python tools/bench_tokenize.py > $O/tokenize.txt 2>&1
python tools/latency_single.py > $O/latency_single.txt 2>&1
[ "$1" = quick ] && exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_step.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sub none > $O/ncu_launches.log 2>&1
cap() {  # name, kernel regex, skip, count
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name "regex:$2" --launch-skip $3 --launch-count $4 -f -o /tmp/$1 \
    python tools/profile_layers.py 64 500 fp32 > /tmp/$1.log 2>&1
  python tools/ncu_table.py /tmp/$1.ncu-rep > $O/ncu_full_$1.csv 2>> $O/ncu.err
  python tools/ncu_summary.py /tmp/$1.ncu-rep > $O/ncu_full_$1_metrics.txt 2>> $O/ncu.err
}
# conv kernel launches of a pass, in order: 35 prenet convs, linear, conv-in, up-sampler 0, then the C=768 k7 conv + its 1x1
cap k7_768_and_1x1 conv_gemm_tc_kernel 38 2
cap resunit_fused resunit_fused_kernel 0 9
cap upsamplers conv_gemm_tc_kernel 44 4
cap stream "head_kernel|dwconv_ln_kernel" 3 2
ls -la $O
