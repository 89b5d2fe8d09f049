"""Executed-instruction / stall-sample profile per code region of one kernel from `ncu --page source --csv` output.
usage: src_regions.py file.csv [kernel-substring] [block]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
B = int(sys.argv[3]) if len(sys.argv) > 3 else 40
kernels = []
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels.append([r[1], None, []])
    elif r and r[0] == "Address":
        kernels[-1][1] = r
    elif kernels and kernels[-1][1] is not None and len(r) == len(kernels[-1][1]):
        kernels[-1][2].append(r)
for name, hdr, data in kernels:
    if want not in name:
        continue
    ix = {h: i for i, h in enumerate(hdr)}
    ex = [int(r[ix['Instructions Executed']] or 0) for r in data]
    sm = [int(r[ix['# Samples']] or 0) for r in data]
    tw = ex[0]
    tot = sum(ex)
    print(name, "warps", tw, "sass", len(data), "instr/warp %.1f" % (tot / tw), "samples", sum(sm))
    for b in range(0, len(data), B):
        e = sum(ex[b:b + B]); s = sum(sm[b:b + B])
        if e > 0.003 * tot:
            print("  %5d-%5d exec/warp %.3f share %5.1f%% samples %5.1f%%  %s" % (b, b + B - 1, e / B / tw, 100 * e / tot, 100 * s / max(sum(sm), 1), data[b][ix['Source']].strip()[:50]))
    break
