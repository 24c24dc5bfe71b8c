"""Opcode histogram (per atom) of the hot loop from an `ncu --page source --csv` dump.
usage: ophist.py dump.csv n_atoms_total [min_exec]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
natoms = float(sys.argv[2])
hdr = rows[1]
si, ei, ci = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) > ei and r[0] != "Address":
        body.append(r)
tot = sum(int(r[ei]) for r in body)
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0
ops, stalls = collections.Counter(), collections.Counter()
for r in body:
    e = int(r[ei])
    if e >= thr:
        t = r[si].strip().split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += e
        stalls[op.split(".")[0]] += float(r[ci] or 0)
print("total warp instr %d = %.1f per atom" % (tot, tot * 32 / natoms))
for op, c in ops.most_common(40):
    print("%-12s %7.2f per atom   stall samples %6.0f" % (op, c * 32 / natoms, stalls[op]))
