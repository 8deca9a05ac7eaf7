"""Developer tool: text summary of an `ncu --set full --import-source on` report (what goes under profiles/).

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [--top 25] > profiles/r01_x.txt

Per kernel: duration, DRAM bytes, throughput percentages, occupancy, issue utilisation, warp execution efficiency, stall
reasons from the warp-state sampler, and the hottest source lines (instructions executed / stall samples).
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads per warp instruction (of 32)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_static", "static smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("sm__inst_executed_pipe_lsu.sum", "LSU instructions"),
]


def ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def raw_page(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append({h: (v, u) for h, v, u in zip(hdr, r, units)})
    return out


def source_page(rep, kernel_id):
    txt = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel_id])
    rows = list(csv.reader(io.StringIO(txt)))
    cur, hdr, agg = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif r[0] not in ("", "Function Name") and hdr and len(r) > 8:
            try:
                agg.append((int(r[7]), int(r[4]), cur, int(r[0]), r[1].strip()[:110], r))
            except ValueError:
                pass
    return hdr, agg


# buckets of the per-line view: (file, first line, last line, name), loaded from --buckets <json> when given
def bucket_of(buckets, fname, line):
    for f, lo, hi, name in buckets:
        if f == fname and lo <= line <= hi:
            return name
    return fname


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    import json
    buckets = json.load(open(sys.argv[sys.argv.index("--buckets") + 1])) if "--buckets" in sys.argv else []
    dump = sys.argv[sys.argv.index("--dump-lines") + 1] if "--dump-lines" in sys.argv else None
    kernels = raw_page(rep)
    print(f"# {rep}: {len(kernels)} profiled launch(es)\n")
    for i, k in enumerate(kernels):
        print(f"## [{i}] {k['Kernel Name'][0]}")
        for key, label in KEYS:
            if key in k:
                print(f"  {label:42s} {k[key][0]} {k[key][1]}")
        print()
    # source-level breakdown, one pass per distinct kernel (ncu prints the first match when not told otherwise)
    seen = set()
    for i, k in enumerate(kernels):
        name = k["Kernel Name"][0]
        if name in seen:
            continue
        seen.add(name)
        short = name.split("(")[0].replace("void ", "").split("::")[-1].strip()
        hdr, agg = source_page(rep, f"regex:{short.split('<')[0]}")
        if not agg:
            continue
        tot = sum(a[0] for a in agg) or 1
        ts = sum(a[1] for a in agg) or 1
        print(f"## source view of {short}: {tot} warp instructions, {ts} stall samples")
        names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        idx = [hdr.index(h) for h in names]
        sums = [0] * len(names)
        for a in agg:
            for j, ix in enumerate(idx):
                try:
                    sums[j] += int(a[5][ix])
                except ValueError:
                    pass
        print("  stall reasons (all samples): " + ", ".join(f"{n[6:]} {100.0 * s / ts:.1f}%" for n, s in sorted(zip(names, sums), key=lambda t: -t[1])[:9]))
        if buckets:
            agg_b = {}
            for a in agg:
                b = bucket_of(buckets, a[2], a[3])
                x = agg_b.setdefault(b, [0, 0])
                x[0] += a[0]
                x[1] += a[1]
            print("  by function (instructions executed / stall samples):")
            for b, (ni, ns) in sorted(agg_b.items(), key=lambda t: -t[1][0]):
                print(f"    {100.0 * ni / tot:5.1f}% inst {100.0 * ns / ts:5.1f}% stall  {b}")
        if dump:
            with open(f"{dump}_{short.split('<')[0]}.csv", "w") as fh:
                fh.write("file,line,inst,stall_samples\n")
                for a in sorted(agg, key=lambda t: (t[2] or "", t[3])):
                    fh.write(f"{a[2]},{a[3]},{a[0]},{a[1]}\n")
        print(f"  top {top} source lines by instructions executed:")
        for a in sorted(agg, key=lambda t: -t[0])[:top]:
            print(f"    {100.0 * a[0] / tot:5.1f}% inst {100.0 * a[1] / ts:5.1f}% stall  {a[2]}:{a[3]:<5d} {a[4]}")
        print(f"  top {top} source lines by stall samples:")
        for a in sorted(agg, key=lambda t: -t[1])[:top]:
            print(f"    {100.0 * a[1] / ts:5.1f}% stall {100.0 * a[0] / tot:5.1f}% inst  {a[2]}:{a[3]:<5d} {a[4]}")
        print()


if __name__ == "__main__":
    main()
