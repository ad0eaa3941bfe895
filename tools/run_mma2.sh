set -u
O=gpurun_out
timeout 300 python tools/diag_unsharp.py > $O/mma_diag.log 2>&1
NCU="ncu --clock-control none"
MB200_MMA=1 $NCU --set full --import-source on -k regex:"conv_mma" -s 4 -c 2 -f -o $O/mma_v1 python tools/devbench.py blur 8192 > $O/mma_ncu.log 2>&1
ncu -i $O/mma_v1.ncu-rep --page raw --csv > $O/mma_v1_raw.csv 2>/dev/null
ncu -i $O/mma_v1.ncu-rep --page source --csv > $O/mma_v1_src.csv 2>/dev/null
cat $O/mma_diag.log; python tools/ncu_pick.py $O/mma_v1_raw.csv
