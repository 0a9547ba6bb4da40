"""Hottest instructions of a kernel from `ncu --page source --csv` (needs --import-source on / -lineinfo): address, samples,
top stall reasons, SASS.  usage: python tools/ncu_hot_instructions.py <rep> [top-n]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 60
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]; idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot = sum(int(r[idx['# Samples']]) for r in body)
print('total samples', tot, 'instructions', len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][idx['# Samples']]))[:topn]
for i in sorted(order):
    r = body[i]; n = int(r[idx['# Samples']])
    st = sorted(((int(r[idx[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    print('%5d %5.2f%% exec %8s | %-40s | %s' % (i, 100.0 * n / tot, r[idx['Instructions Executed']], ' '.join('%s %d' % (b, a) for a, b in st if a), r[idx['Source']].strip()[:90]))
