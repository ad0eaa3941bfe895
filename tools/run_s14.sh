set -u
O=gpurun_out
for t in 1 3 4; do
echo "--- tma=$t" >> $O/s14_dev.log
MB200_RESIZE_TMA=$t timeout 300 python tools/devbench.py resize 4096 >> $O/s14_dev.log 2>&1
MB200_RESIZE_TMA=$t timeout 300 python tools/devbench.py resize 8192 >> $O/s14_dev.log 2>&1
done
MB200_RESIZE_TMA=3 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "resize" 2>&1 | tail -3 >> $O/s14_dev.log
cat $O/s14_dev.log
