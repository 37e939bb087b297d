"""bf16 operands: K-major and MN-major no-swizzle conventions + 6-term bf16x3 accuracy."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragposer_b200 import _lib
lib = C.CDLL(os.path.join(ROOT, "scripts", "probes", "libdp_probe.so"))  # python -m dragposer_b200.build --probes
fn = lib.dp_selftest_umma
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32] + [C.c_uint32] * 6 + [C.c_int] * 6 + [C.c_void_p]

def bf16_bits(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return u.astype(np.uint16)
def bf16_val(bits):
    return (bits.astype(np.uint32) << 16).view(np.float32)

def image16(mat_bits, lbo, sbo, min_bytes=0):
    """K-major bf16 image: element (r,k) at (r//8)*sbo + (k//8)*lbo + (r%8)*16 + (k%8)*2."""
    R, K = mat_bits.shape
    size = ((R + 7) // 8 - 1) * sbo + ((K + 7) // 8 - 1) * lbo + 128
    size = (max(size, min_bytes) + 15) // 16 * 16
    img = np.zeros(size // 2, np.uint16)
    r, k = np.meshgrid(np.arange(R), np.arange(K), indexing="ij")
    img[((r // 8) * sbo + (k // 8) * lbo + (r % 8) * 16 + (k % 8) * 2) // 2] = mat_bits
    return img

def run(a_img, b_img, a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep, N, ksteps, a_mn=0, b_mn=0, passes=1, kind=1):
    d = np.zeros((128, N), np.float32)
    rc = fn(a_img.ctypes.data, a_img.nbytes, b_img.ctypes.data, b_img.nbytes, a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep, N, ksteps, a_mn, b_mn, passes, kind, d.ctypes.data)
    assert rc == 0, rc
    return d

rng = np.random.default_rng(0)
q = lambda *s: (rng.integers(-8, 9, s) / 8.0).astype(np.float32)
for K, N, b_lbo in ((32, 32, 528), (64, 16, 272), (96, 32, 528)):
    A, B = q(128, K), q(N, K)
    a_lbo, a_sbo, b_sbo = 128, 128 * (K // 8), 128
    d = run(image16(bf16_bits(A), a_lbo, a_sbo), image16(bf16_bits(B), b_lbo, b_sbo), a_lbo, a_sbo, b_lbo, b_sbo, 2 * a_lbo, 2 * b_lbo, N, K // 16)
    print(f"bf16 K-major K={K} N={N}: max err {np.abs(d - A @ B.T).max():.3e}")
# MN-major A from a K-major weight image W (Kp=out rows x 128 in cols): A'(m=i,k=o) = W[o][i]
for Kp, N in ((32, 32), (64, 32), (96, 16)):
    Wm, B = q(Kp, 128), q(N, Kp)
    w_lbo, w_sbo = 128, 128 * (128 // 8)
    b_lbo, b_sbo = 528, 128
    ref = Wm.T @ B.T
    for name, (lbo, sbo) in (("LBO=w_sbo,SBO=w_lbo", (w_sbo, w_lbo)), ("LBO=w_lbo,SBO=w_sbo", (w_lbo, w_sbo))):
        d = run(image16(bf16_bits(Wm), w_lbo, w_sbo), image16(bf16_bits(B), b_lbo, b_sbo), lbo, sbo, b_lbo, b_sbo, 2 * w_sbo, 2 * b_lbo, N, Kp // 16, a_mn=1)
        print(f"bf16 MN-major A Kp={Kp} N={N} {name}: max err {np.abs(d - ref).max():.3e}  (|ref| max {np.abs(ref).max():.1f}, |d| max {np.abs(d).max():.1f})")
# accuracy of the 6-term bf16x3 product
K, N = 64, 32
A, B = rng.standard_normal((128, K)).astype(np.float32), rng.standard_normal((N, K)).astype(np.float32)
def split3(x):
    b1 = bf16_bits(x); r1 = x - bf16_val(b1)
    b2 = bf16_bits(r1); r2 = r1 - bf16_val(b2)
    b3 = bf16_bits(r2)
    return b1, b2, b3
A1, A2, A3 = split3(A); B1, B2, B3 = split3(B)
ref = A.astype(np.float64) @ B.astype(np.float64).T
# small terms first: A3B1, A2B2, A1B3, A2B1, A1B2, A1B1  -> concatenate along K
Acat = np.concatenate((A3, A2, A1, A2, A1, A1), 1); Bcat = np.concatenate((B1, B2, B3, B1, B2, B1), 1)
a_lbo, a_sbo, b_lbo, b_sbo = 128, 128 * (6 * K // 8), 528, 128
d = run(image16(Acat, a_lbo, a_sbo), image16(Bcat, b_lbo, b_sbo), a_lbo, a_sbo, b_lbo, b_sbo, 256, 2 * b_lbo, N, 6 * K // 16)
print("bf16x3 (6 terms) rel err", np.abs(d - ref).max() / np.abs(ref).max(), " fp32 matmul", np.abs(A @ B.T - ref).max() / np.abs(ref).max())
Acat = np.concatenate((A2, A1, A1), 1); Bcat = np.concatenate((B1, B2, B1), 1)
a_sbo = 128 * (3 * K // 8)
d = run(image16(Acat, a_lbo, a_sbo), image16(Bcat, b_lbo, b_sbo), a_lbo, a_sbo, b_lbo, b_sbo, 256, 2 * b_lbo, N, 3 * K // 16)
print("bf16x2 (3 terms) rel err", np.abs(d - ref).max() / np.abs(ref).max())
