set -u
O=gpurun_out
timeout 600 python tools/devbench.py taps 8192 > $O/mma_taps.log 2>&1
cat $O/mma_taps.log
