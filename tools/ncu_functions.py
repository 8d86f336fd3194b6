"""Per-function breakdown of an `ncu --page source --csv` export: device functions are delimited by CALL targets."""
import bisect
import collections
import csv
import re
import sys


def main(path, top=28):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data, base = [], None
    for r in rows[2:]:
        if len(r) <= iex:
            continue
        a = int(r[ia], 16)
        if base is None:
            base = a
        data.append((a - base, r[isrc].strip(), int(r[isamp]), int(r[iex])))
    targets = {0}
    for off, src, s, e in data:
        m = re.search(r"CALL\.REL\.NOINC\s+(0x[0-9a-f]+)", src)
        if m:
            t = int(m.group(1), 16)
            targets.add(t - base if t >= base else t)
    targets = sorted(targets)
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, collections.Counter()])
    for off, src, s, e in data:
        f = targets[bisect.bisect_right(targets, off) - 1]
        g = agg[f]
        g[0] += 1
        g[1] += s
        g[2] += e
        parts = src.split()
        op = parts[0] if parts and not parts[0].startswith("@") else (parts[1] if len(parts) > 1 else "")
        if op.startswith("IMAD.WIDE"):
            g[3] += 1
        if op.startswith("CALL"):
            t = int(re.search(r"(0x[0-9a-f]+)", src).group(1), 16)
            g[4][t - base if t >= base else t] += 1
    tot_s = sum(g[1] for g in agg.values())
    tot_e = sum(g[2] for g in agg.values())
    print("functions", len(agg), "total samples", tot_s)
    for f, g in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"off {f:#8x} static {g[0]:5d} wideMAC {g[3]:4d} samples {100 * g[1] / tot_s:5.1f}% executed "
              f"{100 * g[2] / tot_e:5.1f}% calls->{dict((hex(k), v) for k, v in g[4].most_common(5))}")


if __name__ == "__main__":
    main(sys.argv[1])
