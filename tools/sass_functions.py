"""Static per-function opcode histogram of one kernel in a cuobjdump -sass listing: device functions are delimited by
CALL.REL.NOINC targets.  Usage: python tools/sass_functions.py <sass file> <kernel name substring> [--dump <start addr hex>]"""
import bisect
import collections
import re
import sys


def load(path, kernel):
    txt = open(path).read()
    for fn in re.split(r"Function : ", txt)[1:]:
        name = fn.split()[0]
        if kernel in name:
            return name, re.findall(r"/\*([0-9a-f]{4,6})\*/\s+((?:@!?U?P\d+\s+)?)([A-Z0-9_.]+)([^;]*);", fn)
    raise SystemExit("kernel not found")


def main():
    path, kernel = sys.argv[1], sys.argv[2]
    name, ins = load(path, kernel)
    targets = {0}
    for addr, pred, op, rest in ins:
        if op.startswith("CALL"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m:
                targets.add(int(m.group(1), 16))
    targets = sorted(targets)
    if "--dump" in sys.argv:
        start = int(sys.argv[sys.argv.index("--dump") + 1], 16)
        end = targets[bisect.bisect_right(targets, start)] if bisect.bisect_right(targets, start) < len(targets) else 1 << 30
        for addr, pred, op, rest in ins:
            if start <= int(addr, 16) < end:
                print(addr, pred, op, rest)
        return
    agg = collections.defaultdict(collections.Counter)
    calls = collections.defaultdict(collections.Counter)
    for addr, pred, op, rest in ins:
        a = int(addr, 16)
        f = targets[bisect.bisect_right(targets, a) - 1]
        key = op.split(".")[0] + (".WIDE" if "WIDE" in op else "") + (".MOV" if op.startswith("IMAD.MOV") else "")
        agg[f][key] += 1
        if op.startswith("CALL"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m:
                calls[f][int(m.group(1), 16)] += 1
    print(name, "instructions", len(ins))
    for f in targets:
        c = agg[f]
        tot = sum(c.values())
        print("%#8x total %5d wide %4d | %s | calls %s" % (f, tot, c["IMAD.WIDE"], dict(c.most_common(9)),
                                                          {hex(k): v for k, v in calls[f].most_common(6)}))


if __name__ == "__main__":
    main()
