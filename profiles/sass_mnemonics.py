"""SASS mnemonic counts per kernel of libsfx.so (cuobjdump -sass of the sm_100a cubin): the table of
profiles/r02_sass_mnemonics.txt.

    python profiles/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'smplify-x-partial_b200', 'csrc', 'libsfx.so')
COLS = ['UTCHMMA', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDTM', 'UTCBAR', 'MAPA', 'UCGABAR', 'HMMA', 'LDL', 'STL']

def strip_params(name):
    """'void f<(bool)1>(A, B)' -> 'void f<(bool)1>': drops the parameter list (last balanced group)."""
    if not name.endswith(')'):
        return name
    depth = 0
    for i in range(len(name) - 1, -1, -1):
        depth += name[i] == ')'
        depth -= name[i] == '('
        if depth == 0:
            return name[:i]
    return name


sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True,
                       text=True).stdout.split('\n')
counts, order, cur, k = {}, [], None, 0
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = strip_params(names[k]) if k < len(names) and names[k] else m.group(1)
        k += 1
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m and cur:
        op = m.group(1)
        counts[cur]['instr'] += 1
        for c in COLS:
            if op.startswith(c):
                counts[cur][c] += 1
print('SASS mnemonic counts per kernel section of libsfx.so (cuobjdump -sass of the sm_100a cubin; device functions are')
print('linked into the section of the kernel that calls them).  UTC*MMA = tcgen05.mma, UTMALDG = TMA tensor load,')
print('UBLKCP = TMA bulk copy, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, MAPA / UCGABAR = cluster (DSMEM) addressing / barrier.')
print()
print('%-50s %6s' % ('kernel', 'instr') + ''.join(' %8s' % c for c in COLS))
for n in sorted(order, key=lambda n: -counts[n]['instr']):
    print('%-50s %6d' % (n[:50], counts[n]['instr']) + ''.join(' %8d' % counts[n][c] for c in COLS))
