mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:flow_warp --launch-skip 3 --launch-count 1 -o /tmp/fw python tools/microbench.py flow1 > /dev/null 2>&1
ncu -i /tmp/fw.ncu-rep --page source --csv > gpurun_out/r02_flow_source.csv 2>&1
ncu -i /tmp/fw.ncu-rep --page details --csv > gpurun_out/r02_flow_details.csv 2>&1
ncu --set full --clock-control none --import-source on -k regex:vq_finish --launch-skip 2 --launch-count 1 -o /tmp/vq python tools/microbench.py vq1 > /dev/null 2>&1
ncu -i /tmp/vq.ncu-rep --page source --csv > gpurun_out/r02_vqfinish_source.csv 2>&1
python tools/ncu_summary.py /tmp/vq.ncu-rep > gpurun_out/r02_ncu_vqfinish.txt 2>&1
ls -la gpurun_out | grep -E "flow_|vqfin"
