"""Per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels in an `ncu --set full` capture,
averaged per kernel -> JSON for bench.py's roofline.traffic.  usage: ncu_traffic.py REP OUT.json [note]"""
import csv
import io
import json
import subprocess
import sys


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main(rep, out, note=""):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ki, ri, wi, ti = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    agg = {}
    for r in rows[2:]:
        name = r[ki].split("(")[0]
        if "trace_kernel" in r[ki]:
            name = "trace_kernel<any>" if "trace_kernel<(bool)1" in r[ki] or "trace_kernel<1" in r[ki] else "trace_kernel<closest>"
        b = to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi])
        agg.setdefault(name, []).append(b)
    res = {"source": rep.split("/")[-1], "note": note,
           "kernels": {k: {"launches_captured": len(v), "dram_bytes_per_launch": sum(v) / len(v)} for k, v in agg.items()}}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
