"""executed-instruction / stall-sample share per SOURCE line from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`
(CPU box): python scripts/src_lines.py dump.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == "Line No")
ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
agg = collections.OrderedDict(); tot = ts = 0
for r in rows:
    if len(r) <= ii or not r[0].isdigit() or r[2] != "-":        # source lines have "-" as address; sass rows carry addresses
        continue
    try:
        n, s = int(r[ii]), int(r[isamp])
    except ValueError:
        continue
    k = (int(r[0]), r[1].strip()[:110])
    a = agg.setdefault(k, [0, 0]); a[0] += n; a[1] += s; tot += n; ts += s
print("warp instr", tot, "samples", ts)
for (ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n / tot * 100:6.2f}%  smp {s / max(ts, 1) * 100:6.2f}%  L{ln:<4d} {src}")
