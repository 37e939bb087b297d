import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import _lib
lib = _lib.load()
fn = lib.dp_selftest_umma
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32] + [C.c_uint32] * 6 + [C.c_int] * 8 + [C.c_void_p, C.c_void_p]
def bench(kind, N, a_tmem, nacc, ksteps=8, passes=64):
    esz = 2 if kind else 4
    kper = 16 if kind else 8
    K = kper * ksteps
    a_lbo, a_sbo = 128, 128 * (K * esz // 16)
    b_lbo, b_sbo = 128 * (N // 8), 128
    a = np.zeros(max(16 * a_sbo, 128 * K * 4) // 4 + 64, np.float32)
    b = np.zeros((K * esz // 16) * b_lbo // 4 + 64, np.float32)
    d = np.zeros((128, N), np.float32); cyc = np.zeros(1, np.int64)
    rc = fn(a.ctypes.data, a.nbytes, b.ctypes.data, b.nbytes, a_lbo, a_sbo, b_lbo, b_sbo, 2 * a_lbo, 2 * b_lbo, N, ksteps, 0, 0, passes, kind, a_tmem, nacc, d.ctypes.data, cyc.ctypes.data)
    n = ksteps * passes
    return rc, cyc[0] / n, 128 * N * kper / (cyc[0] / n)
for kind, a_tmem, name in ((0, 1, "tf32 TS"), (0, 0, "tf32 SS"), (1, 0, "bf16 SS")):
    for N in (32, 64):
        for nacc in (1, 2, 3, 4, 6, 8):
            if nacc * 32 < N * nacc // max(1, N // 32) and False: continue
            if N == 64 and nacc > 4: continue
            rc, c, rate = bench(kind, 32 if True else N, a_tmem, nacc) if N == 32 else bench(kind, 32, a_tmem, nacc, ksteps=4)
            print(f"{name} N=32 ksteps={'8' if N == 32 else '4'} nacc={nacc}: rc={rc} {c:7.1f} cycles/MMA {rate:7.0f} MAC/cycle")
