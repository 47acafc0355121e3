"""Condensed view of an `ncu --page source --csv` dump: SASS lines with executed-instruction counts and stall samples.
  python tools/ncu_src.py file.csv [min_fraction]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
hdr = rows[1]
iS, iN, iE, iT = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
body = []
for r in rows[2:]:
    if len(r) < len(hdr): break
    body.append(r)
tot_e = sum(int(r[iE]) for r in body); tot_s = sum(int(r[iN]) for r in body)
print("total warp-inst", tot_e, "samples", tot_s)
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for k, r in enumerate(body):
    e, s = int(r[iE]), int(r[iN])
    if e >= thr * tot_e or s >= thr * tot_s:
        st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols if int(r[i]) > 0), reverse=True)[:3]
        print(f"{k:5d} {100*e/tot_e:5.2f}%e {100*s/max(1,tot_s):5.2f}%s thr={int(r[iT])/max(1,e):4.1f} {r[iS].strip()[:70]:70s} {' '.join(f'{n}:{c}' for c,n in st)}")
