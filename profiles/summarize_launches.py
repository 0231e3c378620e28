"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, average ns, share of device time.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_xxx.txt"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or r[mi] != "gpu__time_duration.sum":
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0][:76]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES)")
print("# kernel | launches | avg ns | share of captured device time")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-78s %5d %11.1f %6.1f%%" % (name, c, t / c, 100 * t / tot))
own = {k: v for k, v in agg.items() if "peak_kernel" not in k and not k.startswith("void at::") and "aos_to_soa" not in k}
t2 = sum(a[1] for a in own.values())
print("# the step's own kernels only (peak probes, torch's L2-flush fill and the one-off upload excluded): share of one EM step")
for name, (c, t) in sorted(own.items(), key=lambda kv: -kv[1][1]):
    print("%-78s %5d %11.1f %6.1f%%" % (name, c, t / c, 100 * t / t2))
