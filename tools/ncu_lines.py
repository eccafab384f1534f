"""Aggregate ncu SASS-level warp-stall samples by source line.
usage: ncu_lines.py <nvdisasm -g section txt> <ncu --page source --csv> [topN]
(profiling helper; not part of the product)"""
import csv, re, sys, collections
csv.field_size_limit(10**9)
sec, src = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
addr2line = {}
cur = ("?", 0)
for ln in open(sec):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S+)', ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci = {k: hdr.index(k) for k in hdr}
base = None
tot = collections.Counter(); by = collections.defaultdict(collections.Counter); inst = collections.Counter()
stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for r in rows[h + 1:]:
    if len(r) != len(hdr) or r[0] == "Address": continue
    a = int(r[ci["Address"]], 16)
    if base is None: base = a
    key, op = addr2line.get(a - base, (("?", 0), "?"))
    n = int(r[ci["# Samples"]] or 0)
    tot[key] += n
    inst[key] += int(r[ci["Instructions Executed"]] or 0)
    for k in stall_cols:
        v = int(r[ci[k]] or 0)
        if v: by[key][k[6:]] += v
S = sum(tot.values()); I = sum(inst.values())
print("total samples", S, "warp instructions", I)
byfile = collections.Counter(); ifile = collections.Counter()
for k, v in tot.items(): byfile[k[0]] += v; ifile[k[0]] += inst[k]
for f, v in byfile.most_common(): print("  file %-16s samples %5.1f%%  inst %5.1f%%" % (f, 100 * v / S, 100 * ifile[f] / I))
for k, v in tot.most_common(topn):
    top = ", ".join("%s %d" % kv for kv in by[k].most_common(3))
    print("%-14s:%4d  samples %5.2f%%  inst %5.2f%%  [%s]" % (k[0], k[1], 100 * v / S, 100 * inst[k] / I, top))
