"""samples inside innermost FP64 loops vs elsewhere: python ncu_regions.py rep"""
import csv, subprocess, sys, re, collections
f = sys.argv[1]
out = subprocess.run(['ncu','-i',f,'--page','source','--csv'],capture_output=True,text=True).stdout
lines = out.splitlines(); hi = [i for i,l in enumerate(lines) if l.startswith('"Address"')][0]
rows = list(csv.reader(lines[hi:])); hdr = rows[0]; rows = rows[1:]
iA=hdr.index('Address'); iS=hdr.index('# Samples'); iSrc=hdr.index('Source'); iEx=hdr.index('Instructions Executed')
addr = [int(r[iA],16) for r in rows]; base = addr[0]
idx = {a:i for i,a in enumerate(addr)}
loops=[]
for i,r in enumerate(rows):
    m = re.search(r'BRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)', r[iSrc])
    if m:
        t = int(m.group(1),16)
        if t < addr[i] and t in idx: loops.append((idx[t], i))
# innermost loops containing DFMA
inner=[]
for (s,e) in loops:
    if not any('DFMA' in rows[k][iSrc] for k in range(s,e+1)): continue
    if any((s2>s or e2<e) and s2>=s and e2<=e and (s2,e2)!=(s,e) and any('DFMA' in rows[k][iSrc] for k in range(s2,e2+1)) for (s2,e2) in loops): continue
    inner.append((s,e))
tot = sum(int(r[iS] or 0) for r in rows); tex = sum(int(r[iEx] or 0) for r in rows)
inl = set()
for s,e in inner:
    samples = sum(int(rows[k][iS] or 0) for k in range(s,e+1)); ex = sum(int(rows[k][iEx] or 0) for k in range(s,e+1))
    if samples > 0.005*tot: print('loop rows %d-%d (%d instr): samples %.1f%%  executed %.1f%%' % (s,e,e-s+1,100*samples/tot,100*ex/tex))
    inl.update(range(s,e+1))
ins = sum(int(rows[k][iS] or 0) for k in inl); iex = sum(int(rows[k][iEx] or 0) for k in inl)
print('inner FP64 loops: samples %.1f%% executed %.1f%%; outside: samples %.1f%% executed %.1f%%' % (100*ins/tot, 100*iex/tex, 100-100*ins/tot, 100-100*iex/tex))
# top outside instructions
outside = sorted(((int(rows[k][iS] or 0), k) for k in range(len(rows)) if k not in inl), reverse=True)[:25]
for s,k in outside: print('   %5.2f%% row %d  %s' % (100*s/tot, k, rows[k][iSrc].strip()[:80]))
