set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mma.py -x -q 2>&1 | tail -4 > $O/mma_tests.log
echo "--- dfma" > $O/mma_dev.log
timeout 300 python tools/devbench.py blur 8192 >> $O/mma_dev.log 2>&1
for strip in 512; do for minb in 3 4; do
echo "--- mma strip=$strip minb=$minb" >> $O/mma_dev.log
MB200_MMA=1 MB200_MMA_STRIP=$strip MB200_MMA_MINB=$minb timeout 300 python tools/devbench.py blur 8192 >> $O/mma_dev.log 2>&1
done; done
NCU="ncu --clock-control none"
MB200_MMA=1 MB200_MMA_MINB=4 $NCU --set full --import-source on -k regex:"conv_mma" -s 4 -c 2 -f -o $O/mma_v3 python tools/devbench.py blur 8192 > $O/mma_ncu.log 2>&1
ncu -i $O/mma_v3.ncu-rep --page raw --csv > $O/mma_v3_raw.csv 2>/dev/null
ncu -i $O/mma_v3.ncu-rep --page source --csv > $O/mma_v3_src.csv 2>/dev/null
cat $O/mma_tests.log $O/mma_dev.log; python tools/ncu_pick.py $O/mma_v3_raw.csv
