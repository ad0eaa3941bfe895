import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, util
import imagemagick_b200 as im
from util import P
src = util.make_image(128, 96, 3, seed=41)
src[0, :8, :3] = [0, 1, 2]
src[1, :8, :3] = [65535, 2650, 2651]
src[2, :7, :3] = [[0, 0, 0], [65535, 65535, 65535], [32768, 32768, 32768], [65535, 0, 0], [0, 65535, 0], [0, 0, 65535], [257, 257, 257]]
for frm,to in ((23,11),(11,23),(11,26),(23,26),(26,23)):
    s2=src.copy()
    if frm==11: s2[3, :3, :3] = [[0, 32767.5, 32767.5], [65535, 32767.5, 32767.5], [0, 0, 0]]
    want=s2.copy(); assert util.oracle().orc_colorspace(P(want),128,96,3,frm,to)==0
    b=im.Image(torch.from_numpy(s2.copy()).cuda()); b.colorspace=frm
    im.TransformImageColorspace(b,to)
    got=b.pixels.cpu().numpy()
    d=util.ulp_distance(got,want)
    idx=np.argwhere(d>1)
    print(frm,to,"bad",len(idx))
    for (y,x,c) in idx[:12]:
        print("  px",s2[y,x],"c",c,"got",repr(float(got[y,x,c])),"want",repr(float(want[y,x,c])),"ulp",int(d[y,x,c]))
