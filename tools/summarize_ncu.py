#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/<tag>_launches.csv profiles/<tag>_launches_summary.txt
  python tools/summarize_ncu.py full gpurun_out/<tag>_x.ncu-rep profiles/<tag>_x_ncu_full_summary.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg"]


def launches(src, dst):
    rows = []
    with open(src) as f:
        txt = f.read()
    start = txt.find('"ID"')
    rd = csv.DictReader(io.StringIO(txt[start:]))
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            v_ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
            rows.append((re.sub(r"\(.*", "", r["Kernel Name"])[:60], v_ms))
    agg = collections.OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(v for _, v in rows)
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none over bench.py (per-launch times are cold-cache + serialised: compare SHARES)\n")
        f.write(f"# launches={len(rows)} total_ms={tot:.2f}\n")
        f.write(f"{'kernel':60s} {'launches':>8s} {'total_ms':>10s} {'share':>7s} {'avg_us':>9s}\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:60s} {n:8d} {v:10.3f} {v / tot:7.3f} {v / n * 1e3:9.1f}\n")
    print(open(dst).read())


def full(src, dst, note=""):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; source {src}\n")
        if note:
            f.write(f"# {note}\n")
        for i, row in enumerate(data):
            d = dict(zip(hdr, row))
            f.write(f"--- launch {i}: {d.get('Kernel Name', '')[:90]}\n")
            for m in METRICS:
                if m in d:
                    f.write(f"  {m:72s} {d[m]} {units[hdr.index(m)]}\n")
            stalls = sorted(((float(v.replace(',', '')), k) for k, v in d.items() if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and v), reverse=True)[:6]
            for v, k in stalls:
                f.write(f"  {k:72s} {v:.0f} warp\n")
    print(open(dst).read())


def _num(v):
    try:
        return float(v.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def roofline(src, dst, hbm_gbs="6545", tensor_tf="1399.3"):
    """One line per captured launch: duration, DRAM bytes read + written (ncu), achieved DRAM GB/s against the measured copy
    bandwidth (MEASURED_PEAKS.json), tensor-pipe activity, occupancy and the two leading stall reasons."""
    hbm = float(hbm_gbs)
    tables = []                    # (header, units, rows) per report: ncu picks the unit of a column per report
    if src.endswith(".csv"):      # `ncu -i x.ncu-rep --page raw --csv` exported on the GPU box (several files: comma-separated)
        for part in src.split(","):
            rows = list(csv.reader(open(part)))
            tables.append((rows[0], rows[1], rows[2:]))
    else:
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        tables.append((rows[0], rows[1], rows[2:]))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "nsecond": 1e-9, "usecond": 1e-6,
             "msecond": 1e-3, "second": 1.0}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; source {src}\n")
        f.write(f"# achieved GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum; peak = {hbm:.0f} GB/s (measured copy bandwidth)\n")
        f.write(f"# {'kernel':46s} {'grid':>7s} {'regs':>4s} {'us':>8s} {'dram MB':>8s} {'GB/s':>6s} {'%HBM':>5s} {'tensor%':>7s} {'warps%':>6s}  top stalls\n")
        for hdr, units, row in [(h, u_, r) for h, u_, rows_ in tables for r in rows_]:
            d = dict(zip(hdr, row))
            u = dict(zip(hdr, units))

            def val(m):
                v = _num(d.get(m))
                return None if v is None else v * scale.get(u.get(m, ""), 1.0)
            t = val("gpu__time_duration.sum")
            rdb, wrb = val("dram__bytes_read.sum") or 0.0, val("dram__bytes_write.sum") or 0.0
            name = re.sub(r"\(.*", "", d.get("Kernel Name", "")).replace("rsb::", "")[:46]
            gbs = (rdb + wrb) / t / 1e9 if t else 0.0
            tens = _num(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")) or 0.0
            warps = _num(d.get("sm__warps_active.avg.pct_of_peak_sustained_active")) or 0.0
            stalls = sorted(((_num(v) or 0.0, k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k, v in d.items()
                             if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k), reverse=True)[:2]
            f.write(f"  {name:46s} {d.get('launch__grid_size', ''):>7s} {d.get('launch__registers_per_thread', ''):>4s} {t * 1e6:8.1f} "
                    f"{(rdb + wrb) / 1e6:8.1f} {gbs:6.0f} {100 * gbs / hbm:5.1f} {tens:7.1f} {warps:6.1f}  "
                    + ", ".join(f"{k} {int(v)}" for v, k in stalls) + "\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full, "roofline": roofline}[sys.argv[1]](*sys.argv[2:])
