"""Exploratory probe of the tcgen05 descriptor conventions on a real B200 (run under gpurun)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import _lib

lib = C.CDLL(os.path.join(ROOT, "scripts", "probes", "libdp_probe.so"))  # python -m dragposer_b200.build --probes
fn = lib.dp_selftest_umma
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32] + [C.c_uint32] * 6 + [C.c_int] * 5 + [C.c_void_p]


def kmajor_image(mat, lbo, sbo, nbytes=None):
    """mat (R, K) -> byte image with element (r,k) at (r//8)*sbo + (k//4)*lbo + (r%8)*16 + (k%4)*4."""
    R, K = mat.shape
    size = ((R + 7) // 8 - 1) * sbo + ((K + 3) // 4 - 1) * lbo + 128
    size = max(size, nbytes or 0)
    size = (size + 15) // 16 * 16
    img = np.zeros(size // 4, np.float32)
    r, k = np.meshgrid(np.arange(R), np.arange(K), indexing="ij")
    off = (r // 8) * sbo + (k // 4) * lbo + (r % 8) * 16 + (k % 4) * 4
    img[off // 4] = mat
    return img


def run(a_img, b_img, a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep, N, ksteps, a_mn=0, b_mn=0, passes=1):
    d = np.zeros((128, N), np.float32)
    rc = fn(a_img.ctypes.data, a_img.nbytes, b_img.ctypes.data, b_img.nbytes, a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep, N, ksteps,
            a_mn, b_mn, passes, d.ctypes.data)
    assert rc == 0, rc
    return d


rng = np.random.default_rng(0)
q = lambda *s: (rng.integers(-8, 9, s) / 8.0).astype(np.float32)  # exactly representable in tf32

# ---- 1. K-major A (128 x K), K-major B (N x K)
for K, N, b_lbo in ((24, 32, 528), (64, 32, 528), (96, 16, 128 * 2 + 16), (40, 64, 128 * 8)):
    A, B = q(128, K), q(N, K)
    a_lbo, a_sbo = 128, 128 * (K // 4)
    b_sbo = 128
    d = run(kmajor_image(A, a_lbo, a_sbo), kmajor_image(B, b_lbo, b_sbo), a_lbo, a_sbo, b_lbo, b_sbo, 2 * a_lbo, 2 * b_lbo, N, K // 8)
    print(f"K-major A/B K={K} N={N} b_lbo={b_lbo}: max err {np.abs(d - A @ B.T).max():.3e}")
    d2 = run(kmajor_image(A, a_lbo, a_sbo), kmajor_image(B, b_lbo, b_sbo), a_lbo, a_sbo, b_lbo, b_sbo, 2 * a_lbo, 2 * b_lbo, N, K // 8, passes=2)
    print(f"   accumulate x2: max err {np.abs(d2 - 2 * (A @ B.T)).max():.3e}")

# ---- 2. MN-major A: weight image W (Kp=out rows, 128 in cols) stored K-major; A'(m=i,k=o) = W[o][i]
for Kp, N in ((24, 32), (64, 32), (96, 16)):
    Wm, B = q(Kp, 128), q(N, Kp)
    w_lbo, w_sbo = 128, 128 * (128 // 4)
    b_lbo, b_sbo = 528, 128
    ref = Wm.T @ B.T
    for name, (lbo, sbo) in (("LBO=w_sbo,SBO=w_lbo", (w_sbo, w_lbo)), ("LBO=w_lbo,SBO=w_sbo", (w_lbo, w_sbo))):
        try:
            d = run(kmajor_image(Wm, w_lbo, w_sbo), kmajor_image(B, b_lbo, b_sbo), lbo, sbo, b_lbo, b_sbo, w_sbo, 2 * b_lbo, N, Kp // 8, a_mn=1)
            print(f"MN-major A Kp={Kp} N={N} {name}: max err {np.abs(d - ref).max():.3e}")
        except AssertionError as e:
            print("failed", name, e)

# ---- 3. precision of one tf32 MMA chain and of the 3xTF32 split on random fp32 data (host-side split)
K, N = 64, 32
A, B = rng.standard_normal((128, K)).astype(np.float32), rng.standard_normal((N, K)).astype(np.float32)
def tf32(x):
    u = x.view(np.uint32).astype(np.uint64)
    u = ((u + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return u.view(np.float32)
a_lbo, a_sbo, b_lbo, b_sbo = 128, 128 * (K // 4), 528, 128
ref = A.astype(np.float64) @ B.astype(np.float64).T
d = run(kmajor_image(A, a_lbo, a_sbo), kmajor_image(B, b_lbo, b_sbo), a_lbo, a_sbo, b_lbo, b_sbo, 256, 2 * b_lbo, N, K // 8)
print("plain tf32 rel err", np.abs(d - ref).max() / np.abs(ref).max())
Ah, Bh = tf32(A), tf32(B)
Al, Bl = A - Ah, B - Bh
# emulate 3 accumulating chains by concatenating along K: [Al|Ah|Ah] x [Bh|Bl|Bh]
A3, B3 = np.concatenate((Al, Ah, Ah), 1), np.concatenate((Bh, Bl, Bh), 1)
a_sbo3 = 128 * (3 * K // 4)
d3 = run(kmajor_image(A3, a_lbo, a_sbo3), kmajor_image(B3, b_lbo, b_sbo), a_lbo, a_sbo3, b_lbo, b_sbo, 256, 2 * b_lbo, N, 3 * K // 8)
print("3xTF32 rel err", np.abs(d3 - ref).max() / np.abs(ref).max(), " fp32 matmul rel err", np.abs((A @ B.T) - ref).max() / np.abs(ref).max())
