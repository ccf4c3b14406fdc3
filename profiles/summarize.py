#!/usr/bin/env python
"""Turn an ncu report (--set full, --import-source on) into the text summary committed under profiles/.
usage: python profiles/summarize.py gpurun_out/r02_quick.ncu-rep k_move > profiles/r02_k_move.txt
(the second argument picks one kernel of a report that holds several)

       python profiles/summarize.py --traffic WORKLOAD OUT.json REPORT [REPORT ...]
writes the DRAM bytes per launch of every kernel in the reports (dram__bytes_read.sum + dram__bytes_write.sum; first launch
of each kernel name) together with the sha256 of the kernel sources they were captured on: bench.py reports
`roofline.traffic` / `roofline.physical` from that file, and only while the sources are still the same."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg"]


def main(rep, kern=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    vals = next(r for r in rows[2:] if kern is None or kern in r[ki])
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print(f"# {rep}")
    print(f"kernel: {d.get('Kernel Name', ('?',))[0]}   grid {d.get('Grid Size', ('?',))[0]} block {d.get('Block Size', ('?',))[0]}")
    for k in KEYS:
        if k in d:
            print(f"{k:70s} {d[k][0]:>18s} {d[k][1]}")
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv"] + (["-k", "regex:" + kern] if kern else [])
    src = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hi = next((i for i, r in enumerate(rows) if "Source" in r), None)
    if hi is None:
        print("\n(no source page in this report)")
        return
    hdr = rows[hi]
    rows = rows[hi - 1:]
    st = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    tot = {hdr[i]: 0 for i in st}
    for r in rows[2:]:
        for i in st:
            try:
                tot[hdr[i]] += int(r[i])
            except Exception:
                pass
    ssum = sum(tot.values()) or 1
    print("\nwarp stall sampling (share of all samples):")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {k:40s} {v / ssum * 100:5.1f} %")
    isrc, ismp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    seen, data = set(), []
    for r in rows[2:]:
        if len(r) <= iex or (r[0], r[isrc]) in seen:
            continue
        seen.add((r[0], r[isrc]))
        try:
            data.append((r[isrc].strip(), int(r[iex]), int(r[ismp])))
        except ValueError:
            pass
    S = sum(x[2] for x in data) or 1
    print("\nhottest SASS instructions by samples:")
    for s, c, sm in sorted(data, key=lambda t: -t[2])[:12]:
        print(f"  {sm / S * 100:5.1f} %  x{c:<10d} {s[:100]}")
    tma = [s for s, _, _ in data if "UBLKCP" in s or "UTMA" in s or "SYNCS" in s]
    if tma:
        print("\nTMA / mbarrier SASS present:", sorted(set(x.split()[0] if not x.startswith('@') else x.split()[1] for x in tma)))


def traffic(workload, out, reps):
    import hashlib
    import json
    import re
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    h = hashlib.sha256()
    for f in sorted((root / "flowamok_b200" / "csrc").iterdir()):
        if f.suffix in (".h", ".cuh", ".cu", ".inc"):
            h.update(f.name.encode())
            h.update(f.read_bytes())
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tunit = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
    dram, dur = {}, {}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        ki, ir, iw, it = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        for r in rows[2:]:
            name = re.sub(r"<.*", "", r[ki].replace("void ", "").split("(")[0]).replace("fa::", "").strip()
            if name in dram:
                continue
            dram[name] = float(r[ir].replace(",", "")) * unit[units[ir]] + float(r[iw].replace(",", "")) * unit[units[iw]]
            dur[name] = float(r[it].replace(",", "")) * tunit[units[it]]
    Path(out).write_text(json.dumps({"workload": workload, "csrc_sha256": h.hexdigest()[:16], "dram_bytes": dram,
                                     "ncu_duration_us": dur, "source": "ncu --set full --clock-control none, one launch per kernel: " +
                                     ", ".join(Path(r).name for r in reps)}, indent=1) + "\n")
    print(Path(out).read_text())


if __name__ == "__main__":
    if sys.argv[1] == "--traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4:])
    else:
        main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
