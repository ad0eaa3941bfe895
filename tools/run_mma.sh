set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mma.py -x -q 2>&1 | tail -15 > $O/mma_tests.log
echo "--- dfma" > $O/mma_dev.log
timeout 300 python tools/devbench.py blur 8192 >> $O/mma_dev.log 2>&1
for strip in 512 1024; do for minb in 3 4; do
echo "--- mma strip=$strip minb=$minb" >> $O/mma_dev.log
MB200_MMA=1 MB200_MMA_STRIP=$strip MB200_MMA_MINB=$minb timeout 300 python tools/devbench.py blur 8192 >> $O/mma_dev.log 2>&1
done; done
cat $O/mma_tests.log $O/mma_dev.log
