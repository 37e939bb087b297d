"""fp16 operands, A from tensor memory (TS): packing probe + MN-major B."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import _lib
lib = C.CDLL(os.path.join(ROOT, "scripts", "probes", "libdp_probe.so"))  # python -m dragposer_b200.build --probes
fn = lib.dp_selftest_umma
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32] + [C.c_uint32] * 6 + [C.c_int] * 8 + [C.c_void_p, C.c_void_p]
def image16(mat_bits, lbo, sbo):
    R, K = mat_bits.shape
    size = ((R + 7) // 8 - 1) * sbo + ((K + 7) // 8 - 1) * lbo + 128
    img = np.zeros((size + 15) // 16 * 8, np.uint16)
    r, k = np.meshgrid(np.arange(R), np.arange(K), indexing="ij")
    img[((r // 8) * sbo + (k // 8) * lbo + (r % 8) * 16 + (k % 8) * 2) // 2] = mat_bits
    return img
rng = np.random.default_rng(0)
q = lambda *s: (rng.integers(-8, 9, s) / 8.0).astype(np.float32)
for K, N in ((16, 32), (48, 64), (64, 48)):
    A, B = q(128, K), q(N, K)
    a = np.ascontiguousarray(A.astype(np.float16)).view(np.uint16)
    b_lbo, b_sbo = 128 * (N // 8), 128
    bi = image16(B.astype(np.float16).view(np.uint16), b_lbo, b_sbo)
    d = np.zeros((128, N), np.float32); cyc = np.zeros(1, np.int64)
    rc = fn(a.ctypes.data, a.nbytes, bi.ctypes.data, bi.nbytes, 0, 0, b_lbo, b_sbo, 0, 2 * b_lbo, N, K // 16, 0, 0, 1, 2, 2, 1, d.ctypes.data, cyc.ctypes.data)
    print(f"TS fp16 K={K} N={N}: rc={rc} max err {np.abs(d - A @ B.T).max():.3e} (|ref| {np.abs(A @ B.T).max():.1f})")
    # MN-major B from TMEM-A
    lbo_i, sbo_i = 128 * (K // 8), 128
    img = image16(np.ascontiguousarray(B.T).astype(np.float16).view(np.uint16), lbo_i, sbo_i)
    rc = fn(a.ctypes.data, a.nbytes, img.ctypes.data, img.nbytes, 0, 0, sbo_i, lbo_i, 0, 2 * sbo_i, N, K // 16, 0, 1, 1, 2, 2, 1, d.ctypes.data, cyc.ctypes.data)
    print(f"   with MN-major B: rc={rc} max err {np.abs(d - A @ B.T).max():.3e}")
