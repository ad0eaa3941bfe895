set -u
O=gpurun_out
NCU="ncu --clock-control none"
MB200_RESIZE_TMA=1 $NCU --set full --import-source on -k regex:"resize_h" -s 3 -c 2 -f -o $O/r02b_resize_h_tma python tools/devbench.py resize 4096 > $O/s13_a.log 2>&1
MB200_RESIZE_TMA=0 $NCU --set full --import-source on -k regex:"resize_h" -s 3 -c 2 -f -o $O/r02b_resize_h_cpasync python tools/devbench.py resize 4096 > $O/s13_b.log 2>&1
for r in r02b_resize_h_tma r02b_resize_h_cpasync; do ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null; python tools/ncu_pick.py $O/${r}_raw.csv; done
