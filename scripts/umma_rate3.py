import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import _lib
lib = C.CDLL(os.path.join(ROOT, "scripts", "probes", "libdp_probe.so"))  # python -m dragposer_b200.build --probes
fn = lib.dp_selftest_umma_rate
fn.restype = C.c_longlong
fn.argtypes = [C.c_int] * 4
for kind, ts, name in ((0, 0, "tf32 SS"), (0, 1, "tf32 TS"), (1, 0, "bf16 SS")):
    for nacc in (1, 2):
        row = []
        for N in (16, 32, 48, 64, 128, 256):
            if nacc == 2 and N > 64: continue
            c = fn(kind, ts, nacc, N)
            kper = 16 if kind else 8
            row.append(f"N={N}: {c} cyc ({128 * N * kper / max(c,1):.0f} MAC/cyc)")
        print(f"{name} nacc={nacc}: " + "  ".join(row))
