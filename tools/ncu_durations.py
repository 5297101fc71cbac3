"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/ncu_durations.py file.csv [skip_first_n_of_marker marker]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
seq = []
for r in rows[1:]:
    try:
        seq.append((r[ki].split("(")[0].replace("corb::", "").replace("void ", ""), float(r[vi].replace(",", "")), r[gi]))
    except ValueError:
        pass
if len(sys.argv) > 3:
    idx = [i for i, s in enumerate(seq) if sys.argv[3] in s[0]]
    seq = seq[idx[int(sys.argv[2])]:]
tot, cnt = collections.defaultdict(float), collections.Counter()
for n, t, g in seq:
    tot[n] += t
    cnt[n] += 1
T = sum(tot.values())
print("serialised device time: %.2f ms over %d launches" % (T / 1e6, len(seq)))
for n, t in sorted(tot.items(), key=lambda kv: -kv[1])[:24]:
    print("%-44s %8.1f us x %4d = %7.2f ms (%4.1f%%)" % (n[:44], t / cnt[n] / 1e3, cnt[n], t / 1e6, 100 * t / T))
