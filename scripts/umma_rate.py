"""Measure tcgen05.mma issue->retire cycles per instruction for tf32 / bf16, SS / TS, several N."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import _lib
lib = _lib.load()
fn = lib.dp_selftest_umma
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32] + [C.c_uint32] * 6 + [C.c_int] * 7 + [C.c_void_p, C.c_void_p]
def bench(kind, N, a_tmem, ksteps=8, passes=64):
    esz = 2 if kind else 4
    kper = 16 if kind else 8
    K = kper * ksteps
    a_lbo, a_sbo = 128, 128 * (K * esz // 16)
    b_lbo, b_sbo = 128 * (N // 8), 128
    a = np.zeros(max(16 * a_sbo, 128 * K * 4) // 4 + 64, np.float32)
    b = np.zeros((K * esz // 16) * b_lbo // 4 + 64, np.float32)
    d = np.zeros((128, N), np.float32); cyc = np.zeros(1, np.int64)
    rc = fn(a.ctypes.data, a.nbytes, b.ctypes.data, b.nbytes, a_lbo, a_sbo, b_lbo, b_sbo, 2 * a_lbo, 2 * b_lbo, N, ksteps, 0, 0, passes, kind, a_tmem, d.ctypes.data, cyc.ctypes.data)
    n = ksteps * passes
    macs = 128 * N * kper
    return rc, cyc[0] / n, macs / (cyc[0] / n)
for kind, name in ((0, "tf32"), (1, "bf16")):
    for a_tmem in ((0, 1) if kind == 0 else (0,)):
        for N in (32, 48, 64, 128, 256):
            rc, c, rate = bench(kind, N, a_tmem)
            print(f"{name} {'TS' if a_tmem else 'SS'} M=128 N={N:3d}: rc={rc} {c:7.1f} cycles/MMA  {rate:7.0f} MAC/cycle")
