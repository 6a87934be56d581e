"""executed-instruction share per source line: joins `nvdisasm -g -c` output of one kernel (k.sass)
with the ncu SASS csv of the same kernel (same instruction order).
usage: sass_lines.py k.sass sass.csv [top]"""
import re, csv, collections, sys
cur = None; seq = []
for l in open(sys.argv[1]):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,6}\*/', l): seq.append((cur, l))
rows = list(csv.reader(open(sys.argv[2])))
hdr = next(r for r in rows if "Instructions Executed" in r)
ia = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); isrc = hdr.index("Source")
data = [r for r in rows if len(r) > ia and r[ia].isdigit()]
print("static instructions:", len(seq), len(data))
bad = 0
for (c, l), r in zip(seq, data):
    op1 = re.search(r'\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', l).group(2)
    op2 = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc]).group(2)
    bad += op1 != op2
print("opcode mismatches:", bad)
agg = collections.Counter(); sm = collections.Counter(); st = collections.Counter(); tot = 0; ts = 0
for (c, l), r in zip(seq, data):
    agg[c] += int(r[ia]); sm[c] += int(r[isamp]); st[c] += 1; tot += int(r[ia]); ts += int(r[isamp])
for k, n in agg.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 40):
    print(k, f"exec {n / tot * 100:5.2f}%  samples {sm[k] / ts * 100:5.2f}%  static {st[k]}")
