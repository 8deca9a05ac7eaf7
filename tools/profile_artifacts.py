"""Developer tool: the two derived artefacts under profiles/.

    python tools/profile_artifacts.py dram  gpurun_out/prof_frame.ncu-rep "<source note>" > profiles/dram_bytes.json
    python tools/profile_artifacts.py launches gpurun_out/launches.csv "<command>"        > profiles/rNN_launches_summary.txt

`dram`: dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel of an `ncu --set full` capture (bench.py
copies the dominant kernel's value into roofline.traffic).  `launches`: totals per kernel of an ncu launch list
(`--metrics gpu__time_duration.sum --csv`), and each kernel's SHARE of the captured time.
"""
import csv
import io
import json
import re
import sys

from ncu_summary import raw_page


def short_name(full):
    n = full.split("(")[0].replace("void ", "").split("::")[-1].strip()
    return n


def to_bytes(v, unit):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    x = float(v.replace(",", ""))
    return x * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit, 1)


def dram(rep, note):
    out = {}
    for k in raw_page(rep):
        name = short_name(k["Kernel Name"][0]).split("<")[0]
        rd, wr = to_bytes(*k["dram__bytes_read.sum"]), to_bytes(*k["dram__bytes_write.sum"])
        e = out.setdefault(name, {"dram_bytes_per_launch": 0.0, "dram_read": 0.0, "dram_write": 0.0, "duration_us": 0.0, "launches": 0, "source": note})
        e["dram_read"] += rd
        e["dram_write"] += wr
        e["duration_us"] += to_us(*k["gpu__time_duration.sum"])
        e["launches"] += 1
    for e in out.values():
        n = e["launches"]
        e["dram_read"] /= n
        e["dram_write"] /= n
        e["duration_us"] /= n
        e["dram_bytes_per_launch"] = e["dram_read"] + e["dram_write"]
    print(json.dumps(out, indent=1))


def launches(path, command):
    txt = open(path, errors="replace").read()
    start = txt.find('"ID"')
    rows = list(csv.reader(io.StringIO(txt[start:])))
    hdr = rows[0]
    ni, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        if len(r) <= vi or not re.match(r"^[0-9.,]+$", r[vi] or "x"):
            continue
        name = short_name(r[ni])
        us = to_us(r[vi], r[ui])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values()) or 1.0
    n = sum(a[0] for a in agg.values())
    print(f"# ncu launch list of `{command}` (gpu__time_duration.sum, --clock-control none)")
    print(f"# {n} launches captured; per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
    print(f"{'kernel':34s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:34s} {c:8d} {t:10.1f} {t / c:9.1f} {100 * t / total:6.1f}%")


if __name__ == "__main__":
    {"dram": dram, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
