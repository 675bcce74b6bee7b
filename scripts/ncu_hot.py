"""Top stall locations of a kernel from `ncu --page source --csv` output.  usage: ncu_hot.py file.csv [topn]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
iS, iN, iE = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == 'Kernel Name':
        break
    try:
        data.append((int(r[iN]), r))
    except ValueError:
        pass
tot = sum(n for n, _ in data)
print('total samples', tot, 'instructions', len(data))
top = sorted(range(len(data)), key=lambda k: -data[k][0])[:topn]
for k in sorted(top):
    n, r = data[k]
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stalls if r[i] not in ('', '0')), reverse=True)[:2]
    print(k, n, f"{100 * n / tot:.1f}%", r[iS].strip()[:64], r[iE], st)
