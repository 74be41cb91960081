"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total time, share, count."""
import collections
import csv
import sys


def main(path, top=30):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows:
        n = r["Kernel Name"].split("(")[0][-80:]
        tot[n] += float(r["Metric Value"])
        cnt[n] += 1
    T = sum(tot.values())
    print(f"# {path}: {len(rows)} launches, {T / 1e6:.3f} ms total (cold-cache, serialised: compare shares)")
    for n, t in sorted(tot.items(), key=lambda x: -x[1])[:top]:
        print(f"{t / 1e6:10.3f} ms {100 * t / T:5.1f}%  x{cnt[n]:<4d} avg {t / cnt[n] / 1e3:9.1f} us  {n}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
