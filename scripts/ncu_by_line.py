"""Join an ncu source-page CSV (per-instruction executed counts / stall samples) with nvdisasm line info.
usage: ncu_by_line.py <src.csv> <nvdisasm dump> <kernel substring> [top]"""
import collections, csv, re, sys
src, dump, key = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
linemap = {}; cur = None; on = False
for line in open(dump):
    if line.startswith(".text."):
        on = key in line; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", line)
    if m: linemap[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src)))
hdr = rows[1]; iA = hdr.index("Address"); iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples")
base = int(rows[2][iA], 16)
ex = collections.Counter(); sm = collections.Counter(); tot = 0; tots = 0
for r in rows[2:]:
    off = int(r[iA], 16) - base
    ln = linemap.get(off, (None, ""))[0]
    e = int(r[iE]); s = int(r[iS]); ex[ln] += e; sm[ln] += s; tot += e; tots += s
print("total warp-instructions", tot, "samples", tots)
byf = collections.Counter()
for k, v in ex.items(): byf[k[0] if k else None] += v
print([(k, round(100 * v / tot, 1)) for k, v in byf.most_common()])
for k, v in ex.most_common(top):
    print(f"{100*v/tot:5.1f}% exec {100*sm[k]/max(tots,1):5.1f}% samples  {k}")
