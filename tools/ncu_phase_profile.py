"""Coarse phase profile of a kernel from `ncu --page source --csv`: samples and stall reasons between barriers.
usage: python tools/ncu_phase_profile.py <rep> [kernel-index]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot = sum(int(r[idx['# Samples']]) for r in body)
print('total samples', tot, 'instructions', len(body))
seg, segs = [], []
for r in body:
    seg.append(r)
    op = r[idx['Source']].split()
    name = op[1] if op and op[0].startswith('@') and len(op) > 1 else (op[0] if op else '')
    if name.startswith('BAR') or name.startswith('WARPSYNC') or name.startswith('BRA') and len(seg) > 150:
        segs.append(seg); seg = []
if seg: segs.append(seg)
for s in segs:
    n = sum(int(r[idx['# Samples']]) for r in s)
    if n < tot * 0.004: continue
    st = collections.Counter()
    for r in s:
        for h in stalls: st[h] += int(r[idx[h]] or 0)
    ops = collections.Counter((r[idx['Source']].split()[1] if r[idx['Source']].split()[0].startswith('@') else r[idx['Source']].split()[0]).split('.')[0] for r in s if r[idx['Source']].split())
    ex = sum(int(r[idx['Instructions Executed']]) for r in s)
    top = ', '.join('%s %.0f%%' % (k[6:], 100.0 * v / max(n, 1)) for k, v in st.most_common(5))
    print('%5.1f%% samples  %4d instrs  exec %9d  end=%-28s | %s | %s' % (100.0 * n / tot, len(s), ex, s[-1][idx['Source']].strip()[:28], top,
          ' '.join('%s:%d' % kv for kv in ops.most_common(6))))
