"""Summarise an `ncu --page source --csv` dump: executed-instruction mix and top stall sites."""
import collections
import csv
import sys


def main(path, top=22):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    si, ei, wi = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    agg, tot, samples = collections.Counter(), 0, []
    for r in rows[hi + 1:]:
        if len(r) <= wi:
            continue
        try:
            n, s = int(r[ei]), int(r[wi])
        except ValueError:
            continue
        toks = r[si].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        agg[op.split(".")[0]] += n
        tot += n
        samples.append((s, n, r[si].strip()))
    print("warp-instructions executed:", tot)
    for k, v in agg.most_common(top):
        print(f"{v:12d} {100 * v / tot:5.1f}% {k}")
    ts = sum(s for s, _, _ in samples)
    print("top stall sites (samples, % of all, executed, SASS):")
    for s, n, src in sorted(samples, reverse=True)[:top]:
        print(f"{s:8d} {100 * s / max(ts, 1):5.1f}% {n:10d}  {src[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 22)
