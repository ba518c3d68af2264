"""One line per captured kernel of an ncu report: duration, DRAM bytes, achieved HBM GB/s against the measured copy
bandwidth, pipe activity and the top stall reason.   python tools/ncu_kernels_table.py report.ncu-rep [hbm_peak_GBs]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6540.2
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, k, default=0.0):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return default


def scaled(r, k):
    """value in bytes / seconds regardless of the unit ncu picked"""
    v, u = val(r, k), units[ix[k]] if k in ix else ""
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0,
            "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}
    return v * mult.get(u, 1.0)


stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")]
print("%-44s %10s %10s %9s %7s %7s %7s %7s %5s  %s" % ("kernel", "time us", "DRAM MB", "GB/s", "of HBM", "fp64 %", "dmma %", "lsu %", "regs",
                                                        "top stalls"))
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[ix["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "")
    name = name.split("(")[0][:44]
    t = scaled(r, "gpu__time_duration.sum")
    b = scaled(r, "dram__bytes_read.sum") + scaled(r, "dram__bytes_write.sum")
    tot = sum(val(r, s) for s in stalls) or 1.0
    top = sorted(((val(r, s), s.rsplit("stalled_", 1)[1]) for s in stalls), reverse=True)[:3]
    print("%-44s %10.1f %10.2f %9.1f %6.1f%% %7.1f %7.1f %7.1f %5d  %s"
          % (name, t * 1e6, b * 1e-6, b / t * 1e-9 if t else 0.0, 100 * b / t * 1e-9 / peak if t else 0.0,
             val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
             val(r, "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
             val(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
             int(val(r, "launch__registers_per_thread")),
             ", ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in top)))
