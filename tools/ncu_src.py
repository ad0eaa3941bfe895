"""Summarise an `ncu --page source --csv` dump: top stalled SASS lines + samples by opcode.
usage: ncu_src.py src.csv [top] [kernel_index]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
iex = hdr.index("Instructions Executed")
want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
body, k = [], -1
for r in rows:
    if not r or r[0] in ("Kernel Name", "Address"):
        if r and r[0] == "Kernel Name":
            k += 1
            print("kernel", k, r[1][:100]) if k == want else None
        continue
    if k == want: body.append(r)
istall = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
tot = sum(int(r[isamp] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
byop = collections.Counter(); cnt = collections.Counter()
for r in body:
    op = r[isrc].split()[0] if not r[isrc].strip().startswith('@') else r[isrc].split()[1]
    byop[op] += int(r[isamp] or 0); cnt[op] += 1
for op, s in byop.most_common(14):
    print(f"  {op:22s} {s:7d} {100*s/tot:5.1f}%  ({cnt[op]} instrs)")
idx = sorted(range(len(body)), key=lambda i: -int(body[i][isamp] or 0))[:top]
for i in sorted(idx):
    r = body[i]
    why = " ".join(f"{h[6:]}={r[j]}" for h, j in istall.items() if r[j] not in ("", "0") and int(r[j]) * 5 >= int(r[isamp]))
    print(f"{i:5d} {int(r[isamp]):6d} {100*int(r[isamp])/tot:5.1f}%  {r[isrc].strip()[:70]:70s} {why}")
