"""bf16x3 accuracy + MN-major B (activations) probe."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import importlib.util
spec = importlib.util.spec_from_file_location("p3", os.path.join(os.path.dirname(os.path.abspath(__file__)), "umma_probe3.py"))
src = open(spec.origin).read().split("rng = np.random.default_rng(0)")[0]
exec(src)
rng = np.random.default_rng(1)
K, N = 64, 32
A, B = rng.standard_normal((128, K)).astype(np.float32), rng.standard_normal((N, K)).astype(np.float32)
def split3(x):
    b1 = bf16_bits(x); r1 = x - bf16_val(b1)
    b2 = bf16_bits(r1); r2 = r1 - bf16_val(b2)
    b3 = bf16_bits(r2)
    return b1, b2, b3
A1, A2, A3 = split3(A); B1, B2, B3 = split3(B)
ref = A.astype(np.float64) @ B.astype(np.float64).T
a_lbo, b_lbo, b_sbo = 128, 528, 128
Acat = np.concatenate((A3, A2, A1, A2, A1, A1), 1); Bcat = np.concatenate((B1, B2, B3, B1, B2, B1), 1)
a_sbo = 128 * (6 * K // 8)
d = run(image16(Acat, a_lbo, a_sbo), image16(Bcat, b_lbo, b_sbo), a_lbo, a_sbo, b_lbo, b_sbo, 256, 2 * b_lbo, N, 6 * K // 16)
print("bf16x3 (6 terms) rel err", np.abs(d - ref).max() / np.abs(ref).max(), " fp32 matmul", np.abs(A @ B.T - ref).max() / np.abs(ref).max())
Acat = np.concatenate((A2, A1, A1), 1); Bcat = np.concatenate((B1, B2, B1), 1)
a_sbo = 128 * (3 * K // 8)
d = run(image16(Acat, a_lbo, a_sbo), image16(Bcat, b_lbo, b_sbo), a_lbo, a_sbo, b_lbo, b_sbo, 256, 2 * b_lbo, N, 3 * K // 16)
print("bf16x2 (3 terms) rel err", np.abs(d - ref).max() / np.abs(ref).max())
# MN-major B: image of B^T, i.e. stored as [K rows][N cols] K-major-style blocks: element (k, n) at (k//8)*lbo_img + (n//8)*sbo_img...
q = lambda *s: (rng.integers(-8, 9, s) / 8.0).astype(np.float32)
for K, N in ((32, 32), (96, 16)):
    A, B = q(128, K), q(N, K)
    a_lbo, a_sbo = 128, 128 * (K // 8)
    # store Bt = B.T (K x N) with image16(rows=K, cols=N): element (k,n) at (k//8)*sbo_i + (n//8)*lbo_i + (k%8)*16 + (n%8)*2
    lbo_i, sbo_i = 128 * (K // 8), 128   # n-groups far apart, k-groups adjacent
    img = image16(bf16_bits(B.T.copy()), lbo_i, sbo_i)
    # MN-major descriptor: LBO = stride between K 8-groups (= sbo_i), SBO = stride between N 8-groups (= lbo_i)
    d = run(image16(bf16_bits(A), a_lbo, a_sbo), img, a_lbo, a_sbo, sbo_i, lbo_i, 2 * a_lbo, 2 * sbo_i, N, K // 16, b_mn=1)
    print(f"bf16 MN-major B K={K} N={N}: max err {np.abs(d - A @ B.T).max():.3e} (|ref| {np.abs(A@B.T).max():.1f})")
