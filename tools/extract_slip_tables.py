#!/usr/bin/env python3
"""Development-time helper: read the slip-system geometry the reference defines in
src/mod_crystals.f (fcc :438-515, bcc :516-593, single :594-603, roters :604-704, bcc12 :705-811,
bcc48 :812-1205) and print it as integer Miller
indices (direction b, plane normal n) in the reference's system ORDER, which the
per-system slip history depends on.  Output was pasted (as data, integer form)
into oracle/slip_tables.inc and cpfft_b200/csrc/slip_tables.cuh.

Usage: python tools/extract_slip_tables.py /root/reference/src/mod_crystals.f
"""
import re, sys, math

def main(path):
    src = open(path).read().split('\n')
    blocks = {1: (438, 515), 2: (516, 593), 3: (594, 603), 6: (604, 704), 7: (705, 811), 8: (812, 1205)}
    consts = {'z0': 0, 'f': None, 'f2': 1, 'f3': 1, 'f112': 1, 'f211': 2, 'f123': 1, 'f213': 2, 'f312': 3}
    for st, (a, b) in blocks.items():
        bi, ni = {}, {}
        for line in src[a - 1:b]:
            m = re.search(r'%(bi|ni)\(\s*(\d+)\s*,\s*(\d)\s*\)\s*=\s*([-+]?)\s*([\w.]+)', line)
            if not m:
                continue
            which, s, k, sign, val = m.groups()
            s, k = int(s), int(k)
            if val in ('0', '0.0', '0.d0'):
                v = 0
            elif val == '1.0':
                v = 1
            elif val == 'f':
                v = 1
            else:
                v = consts[val]
            if sign == '-':
                v = -v
            (bi if which == 'bi' else ni).setdefault(s, [0, 0, 0])[k - 1] = v
        n = max(bi)
        print(f'/* slip_type {st}: {n} systems */')
        for s in range(1, n + 1):
            print('  {%2d,%2d,%2d, %2d,%2d,%2d},' % (*bi[s], *ni[s]))

if __name__ == '__main__':
    main(sys.argv[1])
