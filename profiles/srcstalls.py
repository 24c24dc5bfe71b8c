"""Summarise an `ncu --page source --csv` dump: top stalled SASS lines and every memory/branch/barrier instruction."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si, ci, ei = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break  # only the first kernel instance of the dump
    if len(r) > max(ci, ei) and r[0] != "Address":
        body.append(r)
data = [(float(r[ci] or 0), k, r[si].strip(), r[ei]) for k, r in enumerate(body)]
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for v, k, s, e in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print("%7.0f %5.1f%% line %5d exec %9s  %s" % (v, 100 * v / tot, k, e, s[:90]))
print("--- memory / control instructions in program order")
for v, k, s, e in data:
    if any(t in s for t in ("LDG", "STG", "BAR", "BRA", "MUFU", "FRND", "LDS", "STS", "SHFL", "CALL", "RET")):
        print("%7.0f line %5d exec %9s  %s" % (v, k, e, s[:90]))
