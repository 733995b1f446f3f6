"""Per-kernel device times of the last full pass in an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys


def main(path, which=-1):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for row in r:
        v = float(row[vi].replace(",", ""))
        u = row[ui]
        v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
        seq.append((row[ki].split("(")[0], v))
    # a pass starts at its first k_generate; with concurrent path ranges (RTX_OPT_PASS_PARTS) the parts' launches alternate, so the
    # k_generate launches of one pass follow each other
    idx = [i for i, (n, _) in enumerate(seq) if "k_generate" in n and (i == 0 or "k_generate" not in seq[i - 1][0])]
    s = idx[which]
    e = idx[which + 1] if which + 1 < 0 and which + 1 < len(idx) and which != -1 else len(seq)
    tot = 0.0
    agg = {}
    for n, v in seq[s:e]:
        print("%-58s %8.3f ms" % (n[:58], v))
        tot += v
        agg[n] = agg.get(n, 0.0) + v
    print("pass total %.3f ms" % tot)
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print("   %-55s %8.3f ms  %5.1f%%" % (n[:55], v, 100 * v / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else -1)
