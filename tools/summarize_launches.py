"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list into per-kernel totals of ONE step."""
import collections
import csv
import re
import sys


def main(path, marker="split_points_kernel", which=-3):
    """One step = the launches from one `marker` kernel (the first kernel of a step) to the next."""
    rows = []
    lines = [l for l in open(path) if l.startswith('"')]
    rd = csv.reader(lines)
    next(rd)
    for r in rd:
        try:
            rows.append((int(r[0]), r[4], float(r[-1])))
        except ValueError:
            pass
    starts = [i for i, (_, n, _) in enumerate(rows) if marker in n]
    if len(starts) >= 2:
        seg = rows[starts[which]:starts[which + 1]] if which + 1 != 0 else rows[starts[which]:]
    else:
        seg = rows[starts[-1]:] if starts else rows
    agg = collections.OrderedDict()
    for _, n, t in seg:
        k = re.sub(r"\(.*", "", n).replace("void ", "")[:90]
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += t
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%10.1f us %5.1f%%  x%3d  %s" % (v[1] / 1e3, 100 * v[1] / tot, v[0], k))
    print("total %.3f ms over %d launches" % (tot / 1e6, len(seg)))


if __name__ == "__main__":
    main(sys.argv[1], *(sys.argv[2:3] or ["split_points_kernel"]), *([int(sys.argv[3])] if len(sys.argv) > 3 else []))
