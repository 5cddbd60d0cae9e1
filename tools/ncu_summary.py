"""Compact summary of `ncu --page raw --csv` exports and of a gpu__time_duration launch list."""
import collections, csv, re, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(path):
    rows = list(csv.reader(open(path)))
    h, v = rows[0], rows[2]
    print("==", path, "|", v[h.index("Kernel Name")][:70])
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print("   %-70s %-12s %s" % (k, rows[1][i], v[i]))


def launches(path, per, step):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    kn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    data = [(r[kn], float(r[mv].replace(",", "")) / 1e3) for r in rows[h + 1:] if len(r) > mv and r[idc].isdigit()]
    sel = data[step * per:(step + 1) * per]
    agg = collections.OrderedDict()
    for n, t in sel:
        m = re.search(r"(\w+_kernel(<[^>]*>)?|vectorized_elementwise_kernel)", n)
        a = agg.setdefault(m.group(1) if m else n[:30], [0, 0.0]); a[0] += 1; a[1] += t
    tot = sum(v[1] for v in agg.values())
    print("launches in file %d; step %d: %d launches, %.1f us" % (len(data), step, len(sel), tot))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("   %-45s %3d %8.1f us %5.1f %%" % (k, c, t, 100 * t / tot))


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        for p in sys.argv[2:]:
            raw(p)
    else:
        launches(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
