"""Aggregates an ncu source page (SASS) by opcode / by the hottest instruction ranges.
    ncu -i x.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Source' in r)
hdr = rows[hi]
isrc, iex = hdr.index('Source'), hdr.index('Instructions Executed')
ist = hdr.index('Warp Stall Sampling (All Samples)')
agg, stall = collections.Counter(), collections.Counter(); tot = 0; tst = 0
for r in rows[hi + 1:]:
    try: n = int(r[iex]); s = int(r[ist])
    except Exception: continue
    op = re.sub(r'^@!?U?P\w+\s+', '', r[isrc].strip()).split()
    op = op[0].split('.')[0] if op else '?'
    agg[op] += n; tot += n; stall[op] += s; tst += s
print('total executed warp-instrs', tot, ' stall samples', tst)
for k, v in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f'{k:10s} {v:>14d} {100*v/tot:5.1f}%   stall samples {100*stall[k]/max(tst,1):5.1f}%')
