"""Digest of an ncu report (raw + source pages): headline metrics, stall mix, per-segment sample shares. Usage: ncu_digest.py rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum"]
for k in keys:
    print(f"{k:75s} {d.get(k)}")
for k in rows[0]:
    if "utchmma" in k.lower() or "tensor_op" in k.lower() or "pipe_tensor" in k.lower() or "pipe_tc" in k.lower() or "tmem" in k.lower():
        if d[k] not in ("0", "", "n/a"): print(f"{k:75s} {d[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S = ix["# Samples"]
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S]) for r in data)
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stall}
print("samples", tot, " ".join(f"{h[6:]}={100*v/tot:.1f}%" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
EX = ix["Instructions Executed"]; W = ix["L1 Wavefronts Shared"]
f = lambda x: float(x) if x not in ("", "-") else 0.0
print("warp-instructions", sum(f(r[EX]) for r in data), "smem wavefronts (LDS/STS)", sum(f(r[W]) for r in data),
      "SHFL", sum(f(r[EX]) for r in data if "SHFL" in r[1]), "REDUX", sum(f(r[EX]) for r in data if "REDUX" in r[1]))
# segments between barriers
start, acc = 0, 0
for i, r in enumerate(data):
    acc += int(r[S])
    if " BAR." in " " + r[1] or "SYNCS" in r[1] and "TRYWAIT" in r[1]:
        if acc > 0.01 * tot:
            seg = data[start:i + 1]
            sa = {h: sum(int(x[ix[h]]) for x in seg) for h in stall}
            top = " ".join(f"{h[6:]}={100*v/acc:.0f}%" for h, v in sorted(sa.items(), key=lambda x: -x[1])[:4])
            print(f"  SASS {start:5d}-{i:5d} {100*acc/tot:5.1f}% of samples  ends at {r[1].strip()[:40]:40s} {top}")
        start, acc = i + 1, 0
