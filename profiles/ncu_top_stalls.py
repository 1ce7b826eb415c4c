"""Summarise an `ncu --page source --csv` dump: per profiled launch, the SASS instructions holding the most
warp-stall samples.  usage: ncu_top_stalls.py dump.csv [launch_index|-1 for all] [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
for si, sec in enumerate(sections):
    if which >= 0 and si != which:
        continue
    hdr, data = sec["hdr"], sec["data"]
    iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[iS]) for r in data)
    print(f"== launch {si}: {sec['name']} | total samples {tot}")
    top = sorted(range(len(data)), key=lambda k: -int(data[k][iS]))[:n]
    for k in sorted(top):
        r = data[k]
        st = {hdr[i][6:]: int(r[i]) for i in stall_cols if int(r[i]) > 0}
        st = dict(sorted(st.items(), key=lambda x: -x[1])[:3])
        print(f"{k:5d} {int(r[iS]):6d} {100 * int(r[iS]) / tot:5.1f}% ex={r[iEx]:>8s} {r[iSrc].strip()[:72]:72s} {st}")
