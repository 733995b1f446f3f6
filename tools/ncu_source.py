"""Per-source-line hot spots of one kernel in an .ncu-rep.
ncu's SASS page gives per-instruction samples; nvdisasm -g on the cubin of the same build maps offsets to file:line.
usage: ncu_source.py REP KERNEL_REGEX CUBIN MANGLED_SUBSTR [top] [--sass]"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def line_map(cubin, mangled):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    m, cur, infn = {}, None, False
    for l in out.splitlines():
        if l.startswith(".text."):
            infn = mangled in l
            continue
        if not infn:
            continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if mm:
            cur = (mm.group(1).split("/")[-1], int(mm.group(2)))
            continue
        mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if mm:
            m[int(mm.group(1), 16)] = (cur, mm.group(2).strip())
    return m


WHICH = 0


def main(rep, kre, cubin, mangled, top=40, sass=False):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "hdr": None}
            blocks.append(cur)
        elif cur is not None:
            if cur["hdr"] is None:
                cur["hdr"] = row
            else:
                cur["rows"].append(row)
    lm = line_map(cubin, mangled)
    blocks = [x for x in blocks if re.search(kre, x["name"])]
    b = blocks[WHICH]
    h = b["hdr"]
    col = {n: i for i, n in enumerate(h)}

    def g(r, n):
        try:
            return float(r[col[n]])
        except Exception:
            return 0.0
    base = int(b["rows"][0][0], 16)
    print(b["name"][:120])
    tot_s = sum(g(r, "# Samples") for r in b["rows"]) or 1
    tot_i = sum(g(r, "Instructions Executed") for r in b["rows"]) or 1
    tot_t = sum(g(r, "Thread Instructions Executed") for r in b["rows"])
    print("samples %d  warp-inst %d  thread-inst %d  avg threads/inst %.2f" % (tot_s, tot_i, tot_t, tot_t / tot_i))
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = {n: sum(g(r, n) for r in b["rows"]) for n in stalls}
    print("stall totals:", ", ".join("%s %.1f%%" % (n[6:], 100 * v / tot_s) for n, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot_s))
    # opcode histogram
    ops = defaultdict(float)
    for r in b["rows"]:
        op = r[1].strip().split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
        ops[op.split(".")[0]] += g(r, "Instructions Executed")
    print("opcodes:", ", ".join("%s %.1f%%" % (k, 100 * v / tot_i) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:24]))
    if sass:
        rows = sorted(b["rows"], key=lambda r: -g(r, "# Samples"))[:top]
        for r in rows:
            off = int(r[0], 16) - base
            st = max(stalls, key=lambda n: g(r, n))
            loc = lm.get(off, (None, ""))[0]
            print("%6.2f%% smp %6.2f%% inst thr %5.1f  %-18s %-22s %s" % (100 * g(r, "# Samples") / tot_s, 100 * g(r, "Instructions Executed") / tot_i,
                  g(r, "Avg. Threads Executed"), st[6:], "%s:%d" % loc if loc else "?", r[1].strip()[:80]))
        return
    per = defaultdict(lambda: [0.0, 0.0, 0.0, defaultdict(float)])
    for r in b["rows"]:
        off = int(r[0], 16) - base
        loc = lm.get(off, (None, ""))[0] or ("?", 0)
        p = per[loc]
        p[0] += g(r, "# Samples"); p[1] += g(r, "Instructions Executed"); p[2] += g(r, "Thread Instructions Executed")
        for n in stalls:
            p[3][n] += g(r, n)
    print("%7s %7s %6s  %-20s %s" % ("smp%", "inst%", "thr", "top stall", "file:line"))
    for loc, p in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        st = max(p[3].items(), key=lambda kv: kv[1]) if p[3] else ("stall_?", 0)
        print("%6.2f%% %6.2f%% %6.1f  %-20s %s:%d" % (100 * p[0] / tot_s, 100 * p[1] / tot_i, p[2] / max(p[1], 1), "%s %.0f%%" % (st[0][6:], 100 * st[1] / max(p[0], 1)), loc[0], loc[1]))


if __name__ == "__main__":
    for x in sys.argv[1:]:
        if x.startswith("--which="):
            WHICH = int(x.split("=")[1])
    a = [x for x in sys.argv[1:] if x != "--sass" and not x.startswith("--which=")]
    main(a[0], a[1], a[2], a[3], int(a[4]) if len(a) > 4 else 40, "--sass" in sys.argv)
