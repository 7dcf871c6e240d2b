"""Per-opcode and hottest-instruction summary of an `ncu --page source --csv --print-source sass` export (first kernel only).
  ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv ; python tools/ncu_sass.py src.csv [top_n]"""
import csv, sys, collections, io
lines = open(sys.argv[1]).read().split("\n")
blocks = []
cur = []
for l in lines:
    if l.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = []
    else:
        cur.append(l)
blocks.append(cur)
rd = list(csv.DictReader(io.StringIO("\n".join(blocks[0]))))
ops = collections.Counter(); samp = collections.Counter()
tot_inst = 0; tot_samp = 0
for r in rd:
    src = r["Source"].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("MUFU") and "." in op else "")
    n = int(r["Instructions Executed"]); s = int(r["# Samples"])
    ops[op] += n; samp[op] += s; tot_inst += n; tot_samp += s
print(f"total warp-instructions {tot_inst}, samples {tot_samp}")
for op, n in ops.most_common(28):
    print(f"  {op:14s} {n:12d} {100*n/tot_inst:5.1f}%   samples {100*samp[op]/max(tot_samp,1):5.1f}%")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
stall_cols = [c for c in rd[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
print("hottest instructions by samples:")
for i, r in sorted(enumerate(rd), key=lambda ir: -int(ir[1]["# Samples"]))[:top]:
    st = sorted(((int(r[c]), c) for c in stall_cols), reverse=True)[:3]
    print(f"  #{i:5d} {int(r['# Samples']):7d} {100*int(r['# Samples'])/tot_samp:5.1f}%  {r['Source'].strip()[:70]:70s} " + " ".join(f"{c[6:]}={v}" for v, c in st if v))
