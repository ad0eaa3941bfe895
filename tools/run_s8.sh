set -u
O=gpurun_out
for pf in 0 1; do
echo "--- l2pf=$pf" >> $O/s8_taps.log
MB200_MMA_L2PF=$pf timeout 300 python - >> $O/s8_taps.log 2>&1 <<'PY'
import sys; sys.path.insert(0,'.')
import torch, imagemagick_b200 as im
x = im.Image(torch.rand(8192, 8192, 4, device="cuda") * 65535)
def t(fn, iters=7, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(iters):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2]
for s in (1.0,2.0,3.0,4.0):
    print(f"sigma={s} blur {t(lambda: im.BlurImage(x,0.0,s)):.3f} ms")
print(f"unsharp(0,2) {t(lambda: im.UnsharpMaskImage(x,0.0,2.0,1.5,0.02)):.3f} ms")
PY
done
cat $O/s8_taps.log
