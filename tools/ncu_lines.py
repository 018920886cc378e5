"""Per-CUDA-source-line executed instructions and stall samples from an ncu report (needs -lineinfo, --import-source on).
    ncu -i x.ncu-rep --page source --print-source cuda,sass --csv > cs.csv ; python tools/ncu_lines.py cs.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Line No' in r)
hdr = rows[hi]
iln, isrc = hdr.index('Line No'), hdr.index('Source')
iex = hdr.index('Instructions Executed'); ist = hdr.index('Warp Stall Sampling (All Samples)')
lines = []
for r in rows[hi + 1:]:
    if r[iln].strip().isdigit():
        try: lines.append((int(r[iex]), int(r[ist]), int(r[iln]), r[isrc].strip()))
        except ValueError: pass
tot = sum(l[0] for l in lines); ts = sum(l[1] for l in lines)
print('total instrs', tot, 'stall samples', ts)
for ex, st, ln, src in sorted(lines, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f'{100*ex/tot:5.1f}% instr {100*st/max(ts,1):5.1f}% stall  L{ln:<4d} {src[:110]}')
