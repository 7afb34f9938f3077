#!/usr/bin/env python
"""Print the SASS context (with stall reasons) around the hottest stalled instructions of a kernel in an .ncu-rep.
usage: python tools_ncu_context.py REP KERNEL_REGEX [min_pct]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) > idx['# Samples'] and r[idx['# Samples']].isdigit()]
tot = sum(int(r[idx['# Samples']]) for r in body)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for i, r in enumerate(body):
    if 100.0 * int(r[idx['# Samples']]) / tot >= minp:
        for q in body[max(0, i - 8):i + 4]:
            top = sorted(((int(q[idx[s]]), s) for s in stalls if q[idx[s]].isdigit()), reverse=True)[:2]
            print('%6.2f%% %9s thr=%3s %-70s %s' % (100.0 * int(q[idx['# Samples']]) / tot, q[idx['Instructions Executed']],
                  q[idx['Avg. Threads Executed']], q[idx['Source']][:70], ' '.join('%s=%d' % (s[6:], v) for v, s in top if v)))
        print('---')
