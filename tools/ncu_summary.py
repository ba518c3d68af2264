"""Summarise an .ncu-rep (raw + source pages) for the covariance kernel: python tools/ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
out = []
for h, u, v in zip(hdr, units, vals):
    if h in want:
        out.append("%-90s %12s %s" % (h, v, u))
for h, u, v in zip(hdr, units, vals):
    if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"):
        out.append("%-90s %12s" % (h, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samp = collections.Counter()
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
for r in rows[2:]:
    s = r[idx["Source"]].split()
    if not s: continue
    op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
    ops[op] += f(r, "Instructions Executed"); samp[op] += f(r, "# Samples")
te, ts = sum(ops.values()), sum(samp.values())
out.append("opcode mix (warp instructions executed / stall samples):")
for op, c in ops.most_common(18):
    out.append("  %-10s exec %6.2f%%  samples %6.2f%%" % (op, 100 * c / te, 100 * samp[op] / ts))
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
