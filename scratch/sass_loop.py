"""Count instructions of the hottest loops (delimited by backward branches) of a kernel: python sass_loop.py <so> <mangled-substring>"""
import subprocess, sys, re, collections
so, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
cur = None; funcs = {}
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', line)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name: continue
    print('==', name[:100], len(ins), 'instructions')
    addr = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r'BRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)', t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr: loops.append((addr[tgt], i))
    for (s, e) in loops:
        body = ins[s:e + 1]
        if not any('DFMA' in t for _, t in body): continue
        ops = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0] for _, t in body)
        fp64 = sum(v for k, v in ops.items() if k in ('DFMA', 'DADD', 'DMUL', 'DSETP'))
        print('  loop @%#x..%#x: %d instr, fp64 %d, other %d :: %s' % (ins[s][0], ins[e][0], len(body), fp64, len(body) - fp64,
              ' '.join('%s:%d' % kv for kv in ops.most_common())))
