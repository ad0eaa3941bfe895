#!/bin/bash
# r02 final profiling pass (ONE GPU): the bench line, the launch list of the same command under ncu, one `--set full`
# capture of the dominant kernels (the two conv_mma passes of BlurImage).  Numbers printed under ncu are never bench values.
set -u
O=gpurun_out
timeout 900 python bench.py > $O/r02b_bench_line.json 2> $O/r02b_bench.err
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/r02b_launches_bench.csv python bench.py --steps 2 --warmup 3 > $O/r02b_bench_under_ncu.log 2>&1
$NCU --set full --import-source on -k regex:"conv_mma" -s 4 -c 2 -f -o $O/r02b_conv_mma python tools/devbench.py blur 8192 > $O/r02b_ncu_blur.log 2>&1
ncu -i $O/r02b_conv_mma.ncu-rep --page raw --csv > $O/r02b_conv_mma_raw.csv 2>/dev/null
python tools/ncu_pick.py $O/r02b_conv_mma_raw.csv > $O/r02b_conv_mma_pick.txt 2>&1
head -c 1200 $O/r02b_bench_line.json; echo; cat $O/r02b_conv_mma_pick.txt | head -12; wc -l $O/r02b_launches_bench.csv
