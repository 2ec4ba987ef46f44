#!/bin/bash
# One gpurun call that produces the round's ncu evidence (copied from gpurun_out/ to profiles/ by the builder):
#   launch list of exactly one eager forward, and `--set full` captures of the kernels the docs cite.
# Usage (on the GPU box, from the repo root):  bash tools/gpu_profile.sh
set -u
mkdir -p gpurun_out
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum'
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_full_step_launches.csv \
    python bench.py --profile-step > gpurun_out/r02_profile_step.log 2>&1
echo "launch list rc=$?"
# one non-local block (7 GEMMs per image): pre-pass, scores + exp, P v^T
ncu --set full --clock-control none -k regex:'gemm_pair_kernel|gemm_kernel' --launch-skip 7 --launch-count 7 -o gpurun_out/r02_attn \
    python tools/microbench.py attn > gpurun_out/r02_attn.log 2>&1
echo "attn rc=$?"
ncu --set full --clock-control none -k regex:'gemm_pair_kernel|gemm_kernel' --launch-skip 2 --launch-count 1 -o gpurun_out/r02_conv512 \
    python tools/microbench.py conv rb512_80 > gpurun_out/r02_conv512.log 2>&1
ncu --set full --clock-control none -k regex:gemm_tapfuse --launch-skip 2 --launch-count 1 -o gpurun_out/r02_conv64 \
    python tools/microbench.py conv rb64_640 > gpurun_out/r02_conv64.log 2>&1
ncu --set full --clock-control none -k regex:"tap_gather|gemm_tapfuse" --launch-skip 4 --launch-count 2 -o gpurun_out/r02_conv_last \
    python tools/microbench.py last > gpurun_out/r02_conv_last.log 2>&1
ncu --set full --clock-control none -k regex:flow_warp --launch-skip 3 --launch-count 1 -o gpurun_out/r02_flow \
    python tools/microbench.py flow1 > gpurun_out/r02_flow.log 2>&1
ncu --set full --clock-control none -k regex:'gemm_pair_kernel' --launch-skip 2 --launch-count 1 -o gpurun_out/r02_conv128 \
    python tools/microbench.py conv rb128_320 > gpurun_out/r02_conv128.log 2>&1
ncu --set full --clock-control none -k regex:gemm_tapfuse --launch-skip 2 --launch-count 1 -o gpurun_out/r02_vgg_conv1_2 \
    python tools/microbench.py conv vgg64_1280 > gpurun_out/r02_vgg.log 2>&1
# gpurun brings back at most 64 MiB: keep text summaries (tools/ncu_summary.py: the metrics the docs cite), drop the reports
for r in attn conv512 conv128 conv64 vgg_conv1_2 conv_last flow; do
  python tools/ncu_summary.py gpurun_out/r02_$r.ncu-rep > gpurun_out/r02_ncu_$r.txt 2>&1
  rm -f gpurun_out/r02_$r.ncu-rep
done
echo "set-full captures done"
python tools/microbench.py attn last flow1 small conv > gpurun_out/r02_microbench.jsonl 2>gpurun_out/r02_microbench.err
