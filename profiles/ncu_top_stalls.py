"""Top stall sites from `ncu -i X.ncu-rep --page source --csv` (per-SASS-instruction warp-state samples)."""
import collections
import csv
import subprocess
import sys


def main(rep, top=30):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            n = int(r[ix["# Samples"]])
        except ValueError:
            continue
        data.append((n, r))
        for c in stall_cols:
            try:
                tot[c] += int(r[ix[c]])
            except ValueError:
                pass
    N = sum(n for n, _ in data)
    print(rows[0][1])
    print("total samples", N, "instructions", len(data))
    print({k: v for k, v in tot.most_common(12)})
    data.sort(key=lambda x: -x[0])
    for n, r in data[:top]:
        st = {c: int(r[ix[c]]) for c in stall_cols if r[ix[c]] not in ("", "0")}
        t3 = sorted(st.items(), key=lambda x: -x[1])[:3]
        print(f"{n:6d} {100 * n / N:5.1f}%  {r[ix['Source']][:80]:80s} {t3}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
