"""A operand from TMEM (TS mode), tf32."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import _lib
lib = C.CDLL(os.path.join(ROOT, "scripts", "probes", "libdp_probe.so"))  # python -m dragposer_b200.build --probes
fn = lib.dp_selftest_umma
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32] + [C.c_uint32] * 6 + [C.c_int] * 7 + [C.c_void_p]
def kmajor_image(mat, lbo, sbo):
    R, K = mat.shape
    size = ((R + 7) // 8 - 1) * sbo + ((K + 3) // 4 - 1) * lbo + 128
    img = np.zeros((size + 15) // 16 * 4, np.float32)
    r, k = np.meshgrid(np.arange(R), np.arange(K), indexing="ij")
    img[((r // 8) * sbo + (k // 4) * lbo + (r % 8) * 16 + (k % 4) * 4) // 4] = mat
    return img
rng = np.random.default_rng(0)
q = lambda *s: (rng.integers(-8, 9, s) / 8.0).astype(np.float32)
for K, N in ((48, 64), (64, 48), (16, 32)):
    A, B = q(128, K), q(N, K)
    b_lbo, b_sbo = 128 * (N // 8), 128
    d = np.zeros((128, N), np.float32)
    a = np.ascontiguousarray(A)
    bi = kmajor_image(B, b_lbo, b_sbo)
    rc = fn(a.ctypes.data, a.nbytes, bi.ctypes.data, bi.nbytes, 0, 0, b_lbo, b_sbo, 0, 2 * b_lbo, N, K // 8, 0, 0, 1, 0, 1, d.ctypes.data)
    print(f"TS tf32 K={K} N={N}: rc={rc} max err {np.abs(d - A @ B.T).max():.3e} (|ref| {np.abs(A @ B.T).max():.1f})")
