"""SASS view of an .ncu-rep source page: every instruction with its sample count, executions and top stall.
usage: ncu_lines.py rep [min_samples]   (prints the whole kernel; rows below min_samples are abbreviated)"""
import csv, subprocess, sys, io
rep = sys.argv[1]; mins = int(sys.argv[2]) if len(sys.argv) > 2 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ix['# Samples']]) for r in data)
print("total samples", tot, "instructions", len(data))
for n, r in enumerate(data):
    s = int(r[ix['# Samples']])
    if s < mins:
        continue
    st = sorted(((int(r[ix[k]]), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{n:5d} {s:6d} {100*s/tot:5.1f}% ex={int(r[ix['Instructions Executed']]):9d} " + " ".join(f"{k}={v}" for v, k in st if v).ljust(30) + " " + r[ix['Source']].strip()[:100])
