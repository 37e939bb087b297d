import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from umma_probe import run, kmajor_image  # noqa  (runs probe 1 too; fine)

print("---- MN-major address map probe")
N = 32
def probe(lbo, sbo, ksteps, kstep_bytes, img_floats=16384):
    img = np.arange(img_floats, dtype=np.float32)  # value = float index in the image (exact in tf32 up to 2048; use small)
    img = (img % 2048).astype(np.float32)
    B = np.zeros((N, 8 * ksteps), np.float32)
    for o in range(min(N, 8 * ksteps)):
        B[o, o] = 1.0
    b_lbo, b_sbo = 528, 128
    d = run(img, kmajor_image(B, b_lbo, b_sbo), lbo, sbo, b_lbo, b_sbo, kstep_bytes, 2 * b_lbo, N, ksteps, a_mn=1)
    return d  # d[m][k] = float index read for A'(m,k)

for lbo, sbo in ((4096, 128), (128, 4096), (256, 4096), (4096, 256), (128, 1024), (1024, 128)):
    d = probe(lbo, sbo, 1, 0)
    print(f"LBO={lbo} SBO={sbo}: A'(m,k) float-index map, rows m=0..9,16,32,64; cols k=0..7")
    for m in (0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 32, 64, 127):
        print("   m=%3d:" % m, d[m, :8].astype(int).tolist())
