"""Dev-time helper: read the three recursive curve generators of the reference
(src/share/spacecurve_mod.F90: Cinco :39-503, PeanoM :505-681, hilbert :683-769)
and print their per-sub-cell (main axis, main dir, joiner axis, joiner dir)
rules as compact integer codes.  The codes are pasted into tse_mesh.cpp.

code = [lma, lmd, lja, ljd] with
  lma: 0 -> ma, 1 -> (ma+1) mod 2
  lmd: +1 -> md, -1 -> -md
  lja: 0 -> ma, 1 -> (ma+1) mod 2, 2 -> ja (inherit)
  ljd: +1 -> md, -1 -> -md, 0 -> jd (inherit)
"""
import re, sys
src = open('/root/reference/src/share/spacecurve_mod.F90').read()
def table(name):
    m = re.search(r'recursive function %s\(l,type,ma,md,ja,jd\)(.*?)end function %s' % (name, name), src, re.S)
    body = m.group(1)
    out = []
    cur = {}
    for line in body.splitlines():
        line = line.strip()
        mm = re.match(r'(lma|lmd|lja|ljd)\s*=\s*(.*)$', line)
        if not mm: continue
        var, rhs = mm.group(1), mm.group(2).replace(' ', '')
        if var == 'lma':
            cur = {}
            cur['lma'] = 0 if rhs == 'ma' else 1
            assert rhs in ('ma', 'MOD(ma+1,maxdim)'), rhs
        elif var == 'lmd':
            cur['lmd'] = {'md': 1, '-md': -1}[rhs]
        elif var == 'lja':
            if rhs == 'lma': cur['lja'] = cur['lma']
            elif rhs == 'ma': cur['lja'] = 0
            elif rhs == 'MOD(ma+1,maxdim)': cur['lja'] = 1
            elif rhs == 'ja': cur['lja'] = 2
            elif rhs == 'MOD(lma+1,maxdim)': cur['lja'] = (cur['lma'] + 1) % 2
            else: raise ValueError(rhs)
        elif var == 'ljd':
            if rhs == 'lmd': cur['ljd'] = cur['lmd']
            elif rhs == 'md': cur['ljd'] = 1
            elif rhs == '-md': cur['ljd'] = -1
            elif rhs == 'jd': cur['ljd'] = 0
            elif rhs == '-lmd': cur['ljd'] = -cur['lmd']
            else: raise ValueError(rhs)
            out.append([cur['lma'], cur['lmd'], cur['lja'], cur['ljd']])
    return out
for n in ('hilbert', 'PeanoM', 'Cinco'):
    t = table(n)
    print(n, len(t))
    print('{' + ','.join('{%d,%d,%d,%d}' % tuple(r) for r in t) + '}')
