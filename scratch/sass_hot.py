"""Hot region of an `ncu --page source --csv --print-source sass` export: python scratch/sass_hot.py file.sass.csv.gz [top]
Prints the contiguous instruction window holding the most samples (the hot loop) with per-instruction samples and main stall."""
import csv, gzip, sys
rows = list(csv.reader(gzip.open(sys.argv[1], 'rt')))
hi = next(i for i, r in enumerate(rows) if 'Address' in r)
hdr = rows[hi]; body = rows[hi + 1:]
S = hdr.index('# Samples'); SRC = hdr.index('Source'); EX = hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
def num(x):
    try: return float(x)
    except ValueError: return 0.0
samp = [num(r[S]) for r in body]
tot = sum(samp)
# find loop: window between a backward branch target... simply take the instructions with the most executions
exe = [num(r[EX]) for r in body]
mx = max(exe)
hot = [i for i, e in enumerate(exe) if e > 0.5 * mx]
lo, hi2 = min(hot), max(hot)
print("total samples %d; hot window %d..%d (%d instr) holds %.1f%% of samples, executed %.3g x" % (tot, lo, hi2, hi2 - lo + 1, 100 * sum(samp[lo:hi2 + 1]) / tot, mx))
agg = {}
for i in range(lo, hi2 + 1):
    r = body[i]
    st = sorted(((num(r[k]), hdr[k][6:]) for k in stalls), reverse=True)[:2]
    print("%5d %6.0f %5.2f%%  %-60s %s" % (i, samp[i], 100 * samp[i] / tot, r[SRC].strip()[:60], ' '.join('%s:%d' % (n, v) for v, n in st if v > 0)))
    for k in stalls: agg[hdr[k][6:]] = agg.get(hdr[k][6:], 0) + num(r[k])
t = sum(agg.values()) or 1
print("stalls in window:", ' '.join('%s:%.1f%%' % (k, 100 * v / t) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
