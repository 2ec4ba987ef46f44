mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k 'regex:tapfuse_kernel<\(int\)64, \(int\)1' -c 2 -o /tmp/vgg python bench.py --profile-step > gpurun_out/r02_vgg_prof.log 2>&1
ls -la /tmp/vgg.ncu-rep
ncu -i /tmp/vgg.ncu-rep --page source --csv > gpurun_out/r02_vgg_source.csv 2>&1
python tools/ncu_summary.py /tmp/vgg.ncu-rep > gpurun_out/r02_ncu_vgg.txt 2>&1
ls -la gpurun_out | grep -E "vgg"
