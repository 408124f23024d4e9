"""Aggregate ncu `--page source --print-source cuda,sass --csv` output per CUDA source line.
usage: ncu -i rep --page source --csv --print-source cuda,sass --kernel-name K ... | python tools/ncu_lines.py [top]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
fpath = None; hdr = None; items = []; tot = 0; totinst = 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed'); ti = hdr.index('Thread Instructions Executed'); continue
    if hdr is None or r[0] == '': continue
    try:
        v = int(r[si]); ins = int(r[ii] or 0); th = int(r[ti] or 0)
    except ValueError:
        continue
    tot += v; totinst += ins
    if v > 0 or ins > 0: items.append((v, ins, th, fpath, r[0], r[1].strip()[:100]))
items.sort(reverse=True)
print('total samples', tot, 'total warp instructions', totinst)
for v, ins, th, f, ln, src in items[:top]:
    print(f"{v:6d} {100*v/max(tot,1):5.1f}% inst={ins:9d} lanes={th/max(ins,1):5.1f} {f}:{ln} {src}")
