set -u
O=gpurun_out
MB200_RESIZE_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r02.py tests/test_gpu_fullsize.py -x -q -m gpu -k "resize or config3" 2>&1 | tail -8 > $O/s12_tests.log
for t in 0 1 2; do
echo "--- tma=$t" >> $O/s12_dev.log
MB200_RESIZE_TMA=$t timeout 300 python tools/devbench.py resize 4096 >> $O/s12_dev.log 2>&1
MB200_RESIZE_TMA=$t timeout 300 python tools/devbench.py resize 8192 >> $O/s12_dev.log 2>&1
done
cat $O/s12_tests.log $O/s12_dev.log
