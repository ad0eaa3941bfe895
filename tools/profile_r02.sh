#!/bin/bash
# r02 profiling pass (run under gpurun on ONE GPU): launch list of the bench command + one `--set full` capture of the
# dominant kernels; raw pages exported to CSV next to the reports.  Numbers printed under ncu are never bench values.
set -u
OUT=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $OUT/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 > $OUT/r02_bench_under_ncu.log 2>&1
$NCU --set full --import-source on -k regex:"conv_pair" -s 8 -c 6 -f -o $OUT/r02_conv_pair python tools/devbench.py blur 8192 > $OUT/r02_ncu_blur.log 2>&1
$NCU --set full --import-source on -k regex:"resize_._stream" -s 6 -c 4 -f -o $OUT/r02_resize python tools/devbench.py resize 4096 > $OUT/r02_ncu_resize.log 2>&1
$NCU --set full --import-source on -k regex:"conv2d_dense|colorspace_kernel|histogram_kernel" -s 4 -c 6 -f -o $OUT/r02_misc python tools/devbench.py conv2d 4096 > $OUT/r02_ncu_misc.log 2>&1
for r in r02_conv_pair r02_resize r02_misc; do
  ncu -i $OUT/$r.ncu-rep --page raw --csv > $OUT/${r}_raw.csv 2>/dev/null
done
ls -la $OUT/*.ncu-rep
