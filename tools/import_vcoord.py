"""Dev-time helper: convert the reference's ACME 72-level hybrid-coordinate
tables (test/vcoord/acme-72i.ascii, acme-72m.ascii; read the way
hybvcoord_mod.F90:36-153 reads them) into one JSON fixture that travels with
the repo (the GPU box has no /root/reference)."""
import json, sys
def read(path):
    toks = open(path).read().split('\n')
    vals = []
    out = []
    i = 0
    lines = [l.strip() for l in toks if l.strip()]
    pos = 0
    while pos < len(lines):
        n = int(lines[pos].split()[0]); pos += 1
        arr = []
        while len(arr) < n:
            arr += [float(t.replace('D', 'E').replace('d', 'e')) for t in lines[pos].replace(',', ' ').split()]
            pos += 1
        out.append(arr)
    return out
hyai, hybi = read('/root/reference/test/vcoord/acme-72i.ascii')
hyam, hybm = read('/root/reference/test/vcoord/acme-72m.ascii')
assert len(hyai) == 73 and len(hybi) == 73 and len(hyam) == 72 and len(hybm) == 72
json.dump({'source': 'reference test/vcoord/acme-72{i,m}.ascii', 'hyai': hyai, 'hybi': hybi, 'hyam': hyam, 'hybm': hybm},
          open('transport_se_b200/data/acme72_vcoord.json', 'w'), indent=0)
print('ok', hyai[0], hyai[-1], hybi[0], hybi[-1])
