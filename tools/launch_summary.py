"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (per-launch times are
cold-cache and serialised: compare SHARES)."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
ik, iv, iu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
    tot[r[ik]] += v; cnt[r[ik]] += 1
T = sum(tot.values())
for k, v in tot.most_common():
    print(f"{k[:90]:90s} n={cnt[k]:4d} total_us={v:12.1f} avg_us={v / cnt[k]:10.2f} share={v / T:.4f}")
