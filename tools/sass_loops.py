"""Instruction mix of the FFMA2-heavy loops of one kernel in a built .so (exploration helper).
usage: python tools/sass_loops.py lib.so <substring of the mangled kernel name>"""
import re, subprocess, sys
from collections import Counter
so, fun = sys.argv[1], sys.argv[2]
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
for blk in out.split('Function : ')[1:]:
    name = blk.split('\n', 1)[0]
    if fun not in name:
        continue
    ins = []
    for l in blk.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    print(name[:110], len(ins), 'instructions')
    for i, (a, t) in enumerate(ins):
        m = re.search(r'BRA\S*\s+(?:.*\s)?0x([0-9a-f]+)', t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                body = ins[addr[tgt]:i + 1]
                n2 = sum(1 for _, x in body if 'FFMA2' in x or 'FMUL2' in x)
                if n2 >= 16 and len(body) < 600:
                    c = Counter((x.split()[1] if x.startswith('@') else x.split()[0]) for _, x in body)
                    print('  loop %#x..%#x: %d instr, FFMA2+FMUL2 %d -> %s' % (tgt, a, len(body), n2, dict(c.most_common(12))))
