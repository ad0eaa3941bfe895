"""Decode the scheduling control fields of `cuobjdump -sass` output (sm_90/sm_100 128-bit encoding):
stall count, yield, write/read scoreboard index, wait mask.  usage: sass_ctrl.py file.sass [grep-regex]"""
import re, sys
lines = open(sys.argv[1]).read().split("\n")
pat = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/")
pat2 = re.compile(r"^\s*/\* (0x[0-9a-f]{16}) \*/")
out = []
i = 0
while i < len(lines):
    m = pat.match(lines[i])
    if m and i + 1 < len(lines):
        m2 = pat2.match(lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ctrl = (hi >> 41) & 0x7fffff
            stall = ctrl & 0xf; yld = (ctrl >> 4) & 1; wr = (ctrl >> 5) & 7; rd = (ctrl >> 8) & 7; wait = (ctrl >> 11) & 0x3f
            out.append((m.group(1), m.group(2).strip(), stall, yld, wr, rd, wait))
            i += 2
            continue
    i += 1
rx = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
for n, (addr, ins, stall, yld, wr, rd, wait) in enumerate(out):
    if rx and not rx.search(ins): continue
    w = "".join(str(b) if wait >> b & 1 else "-" for b in range(6))
    print(f"{n:5d} {addr} st={stall:2d} y={yld} W={'-' if wr == 7 else wr} R={'-' if rd == 7 else rd} wait={w}  {ins[:90]}")
