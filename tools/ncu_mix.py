"""Dynamic SASS opcode mix per kernel from `ncu --page source --csv --print-source sass`
(profiling helper). usage: ncu_mix.py <csv> [kernel-substring]"""
import csv, sys, collections
csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
sub = sys.argv[2] if len(sys.argv) > 2 else ""
i = 0; seen = set()
while i < len(rows):
    r = rows[i]
    if r and r[0] == "Kernel Name":
        name = r[1]; hdr = rows[i + 1]; ci = {k: j for j, k in enumerate(hdr)}
        j = i + 2; mix = collections.Counter(); smp = collections.Counter()
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            rr = rows[j]
            if len(rr) == len(hdr):
                op = rr[ci["Source"]].split()
                op = [o for o in op if not o.startswith("@")]
                o = op[0].split(".")[0] if op else "?"
                mix[o] += int(rr[ci["Instructions Executed"]] or 0)
                smp[o] += int(rr[ci["# Samples"]] or 0)
            j += 1
        if sub in name and name not in seen:
            seen.add(name)
            T = sum(mix.values()); S = sum(smp.values())
            print(name.split("(")[0], "warp-inst", T, "samples", S)
            for o, v in mix.most_common(22):
                print("  %-8s %10d %5.1f%%   samples %5.1f%%" % (o, v, 100 * v / T, 100 * smp[o] / max(S, 1)))
        i = j
    else:
        i += 1
