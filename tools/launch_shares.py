"""Per-kernel shares of an ncu launch list (--metrics gpu__time_duration.sum --csv).
usage: launch_shares.py launches.csv [skip_first_n]   (profiling helper; not part of the product)"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("pb::", "").replace("(anonymous namespace)::", "")
    us = float(r[14]) / (1e3 if r[13] == "ns" else 1.0)
    tot[name] += us; cnt[name] += 1
T = sum(tot.values())
print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
for k, v in tot.most_common():
    print(f"| {k} | {cnt[k]} | {v:.1f} | {v / cnt[k]:.1f} | {100 * v / T:.1f}% |")
rhs = sum(v for k, v in tot.items() if k.startswith("k_pre") or k.startswith("k_main"))
nrhs = sum(c for k, c in cnt.items() if k.startswith("k_main"))
print(f"\ntotal {T:.1f} us over {len(rows)} launches; RHS kernels {100 * rhs / T:.1f}% ({rhs / max(nrhs, 1):.1f} us per evaluation, {nrhs} evaluations); "
      f"other kernels {(T - rhs) / max(nrhs, 1):.1f} us per evaluation")
