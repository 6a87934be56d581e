"""opcode histogram (weighted by executed warp instructions) + stall-sample share from
`ncu -i X.ncu-rep --page source --csv --print-source sass --kernel-name regex:K > sass.csv`"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Instructions Executed" in r)
ia = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); isrc = hdr.index("Source")
ops = collections.Counter(); samp = collections.Counter(); tot = 0; ts = 0; nstatic = 0
for r in rows:
    if len(r) <= ia or r is hdr or not r[ia].isdigit():
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    if not m:
        continue
    op = m.group(2).split('.')[0]
    n = int(r[ia]); ops[op] += n; tot += n; s = int(r[isamp]); samp[op] += s; ts += s; nstatic += 1
print("static instr", nstatic, "executed warp instr", tot, "samples", ts)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:12s} {n / tot * 100:6.2f}%  samples {samp[op] / max(ts, 1) * 100:6.2f}%")
