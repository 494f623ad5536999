"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
   python tools/launch_summary.py launches.csv [top N]"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    n = r[ki]
    n = re.sub(r"\(anonymous namespace\)::|<unnamed>::|unnamed>::", "", n)   # (so that the next two lines cut at the argument list)
    n = re.sub(r"\(.*", "", n)
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"<.*", "", n)[:80]
    if "spin_kernel" in n:   # torch.cuda._sleep in front of bench.py's eager timing steps: not part of the step
        continue
    agg[n][0] += 1
    agg[n][1] += v
    tot += v
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print("total %.1f us over %d launches (cold-cache, serialised: compare shares)" % (tot / 1e3, len(rows) - 1))
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print("%9.1f us %5d  %5.1f%%  %s" % (t / 1e3, c, 100 * t / tot, n))
