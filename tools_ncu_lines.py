#!/usr/bin/env python
"""Per-source-line totals (warp instructions, stall samples) of a kernel from an .ncu-rep, via the cuda,sass view.
usage: python tools_ncu_lines.py REP KERNEL_REGEX [top]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.OrderedDict()
fname, hdr = None, None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; i_s = hdr.index('# Samples'); i_e = hdr.index('Instructions Executed'); i_w = hdr.index('L1 Wavefronts Shared'); continue
    if hdr is None or len(r) <= i_s: continue
    if r[0].strip():
        cur = (fname, r[0], r[1].strip()[:90]); agg.setdefault(cur, [0, 0, 0])
    if r[i_s].isdigit():
        a = agg[cur]; a[0] += int(r[i_s]); a[1] += int(r[i_e]) if r[i_e].isdigit() else 0; a[2] += int(r[i_w]) if r[i_w].isdigit() else 0
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values()); tw = sum(a[2] for a in agg.values())
print("total samples %d, warp inst %d, smem wavefronts %d" % (ts, ti, tw))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% samp %5.1f%% inst %5.1f%% smem  %s:%s  %s" % (100*a[0]/max(ts,1), 100*a[1]/max(ti,1), 100*a[2]/max(tw,1), k[0], k[1], k[2]))
