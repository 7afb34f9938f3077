#!/usr/bin/env python
"""Sample share per code region of one kernel in an .ncu-rep (regions = runs of SASS with the same execution count).
usage: python tools_ncu_regions.py x.ncu-rep KERNEL_REGEX [launch_index] [min_pct]"""
import csv
import io
import subprocess
import sys


def main():
    rep, regex = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    heads = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r]
    h = heads[which]
    end = heads[which + 1] - 1 if which + 1 < len(heads) else len(rows)
    idx = {k: i for i, k in enumerate(rows[h])}
    S, IE, SRC, TH = idx["# Samples"], idx["Instructions Executed"], idx["Source"], idx["Avg. Threads Executed"]
    body = [r for r in rows[h + 1:end] if len(r) > IE and r[IE].isdigit()]
    tot = sum(int(r[S]) for r in body)
    tot_i = sum(int(r[IE]) for r in body)
    print("samples %d, warp instructions %d, SASS lines %d" % (tot, tot_i, len(body)))
    seg, cur = [], None
    for i, r in enumerate(body):
        n, s = int(r[IE]), int(r[S])
        if cur and abs(n - cur["n"]) <= 0.02 * max(n, cur["n"]) + 2:
            cur["s"] += s
            cur["cnt"] += 1
            cur["end"] = i
            cur["inst"] += n
        else:
            if cur:
                seg.append(cur)
            cur = {"n": n, "s": s, "cnt": 1, "start": i, "end": i, "thr": r[TH], "inst": n}
    seg.append(cur)
    for g in seg:
        if g["s"] > min_pct / 100 * tot:
            print("%5d-%5d  exec=%10d thr=%4s lines=%4d inst=%5.2f%% samples=%5.2f%%  %s" % (
                g["start"], g["end"], g["n"], g["thr"], g["cnt"], 100 * g["inst"] / tot_i, 100 * g["s"] / tot,
                body[g["start"]][SRC].strip()[:50]))


if __name__ == "__main__":
    main()
