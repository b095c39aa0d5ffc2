"""dump normalised SASS per kernel (addresses/encodings stripped) for before/after comparisons:
   python scratch/sass_dump.py <lib.so> <outdir>"""
import subprocess, sys, re, os, hashlib
so, outdir = sys.argv[1], sys.argv[2]
os.makedirs(outdir, exist_ok=True)
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
cur = None; funcs = {}
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', line)
    if m and cur is not None:
        funcs[cur].append(m.group(2).strip())
for name, ins in funcs.items():
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    short = re.sub(r'qmc::\(anonymous namespace\)::', '', dem)
    short = re.sub(r'\(.*', '', short).replace('void ', '').replace(' ', '')
    body = "\n".join(ins)
    h = hashlib.md5(body.encode()).hexdigest()[:12]
    open(os.path.join(outdir, re.sub(r'[^A-Za-z0-9_<>,-]', '_', short) + '.sass'), 'w').write(body + "\n")
    print('%-60s %6d instr  %s' % (short, len(ins), h))
