"""Attribute SASS instruction counts of one kernel to source lines (needs -lineinfo).
usage: sass_by_line.py <nvdisasm --print-line-info dump> <kernel-name-substring> [top]"""
import collections, re, sys
dump, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cnt = collections.Counter(); cur = None; on = False
for line in open(dump):
    if line.startswith(".text."):
        on = key in line
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        cnt[cur] += 1
tot = sum(cnt.values())
print("total instructions", tot, "=", tot * 16 // 1024, "KB")
byfile = collections.Counter()
for k, v in cnt.items():
    byfile[k[0] if k else None] += v
print(byfile.most_common())
for k, v in cnt.most_common(top):
    print(v, k)
