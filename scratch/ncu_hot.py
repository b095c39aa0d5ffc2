import csv, subprocess, sys, collections, re
f = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu','-i',f,'--page','source','--csv'],capture_output=True,text=True).stdout
lines = out.splitlines()
# find header line
hi = [i for i,l in enumerate(lines) if l.startswith('"Address"')][0]
rows = list(csv.reader(lines[hi:]))
hdr = rows[0]; rows = rows[1:]
iS = hdr.index('# Samples'); iSrc = hdr.index('Source'); iEx = hdr.index('Instructions Executed')
tot = sum(int(r[iS] or 0) for r in rows)
print('total samples', tot, 'instructions', len(rows))
byop = collections.Counter(); exop = collections.Counter()
for r in rows:
    op = re.sub(r'^@!?U?P\d+\s+', '', r[iSrc].strip()).split()[0]
    byop[op.split('.')[0]] += int(r[iS] or 0); exop[op.split('.')[0]] += int(r[iEx] or 0)
print('samples by opcode:', ' '.join('%s:%.1f%%' % (k, 100*v/tot) for k,v in byop.most_common(18)))
texe = sum(exop.values())
print('executed by opcode:', ' '.join('%s:%.1f%%' % (k, 100*v/texe) for k,v in exop.most_common(18)))
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i][iS] or 0))[:topn]
for i in sorted(idx):
    r = rows[i]
    print('%6d %5.2f%% ex=%-9s %s' % (i, 100*int(r[iS] or 0)/tot, r[iEx], r[iSrc].strip()[:90]))
