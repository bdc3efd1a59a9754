"""Share of device time per kernel from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0][-70:]
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        tot[name] += v * scale
        cnt[name] += 1
    T = sum(tot.values())
    print(f"launches={sum(cnt.values())} total={T:.1f} us")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{v / T * 100:6.2f}%  {v:10.1f} us  n={cnt[k]:4d}  avg={v / cnt[k]:8.2f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
