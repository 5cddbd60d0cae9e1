"""Opcode histogram (executed warp instructions, stall samples) of `ncu --page source --csv --print-source sass` exports."""
import collections, csv, sys


def main(path, top=22):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and len(r) > 5)
    h = rows[hi]
    ie, isrc, ismp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    tot = 0; agg = collections.Counter(); smp = collections.Counter(); segs = []
    for r in rows[hi + 1:]:
        if len(r) <= ie:
            continue
        try:
            n = int(r[ie]); s = int(r[ismp])
        except ValueError:
            continue
        w = r[isrc].split()
        op = (w[1] if w and w[0].startswith("@") else (w[0] if w else "?")).split(".")[0]
        agg[op] += n; smp[op] += s; tot += n
        segs.append((n, s, r[isrc][:60]))
    print("%s: %d warp instructions executed" % (path, tot))
    for k, v in agg.most_common(top):
        print("   %-10s %12d %5.1f %%   samples %d" % (k, v, 100.0 * v / tot, smp[k]))
    return segs


if __name__ == "__main__":
    main(sys.argv[1])
