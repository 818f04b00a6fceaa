"""Accuracy of a radix-32 DFT stage computed as a TF32 GEMM (numpy emulation of the 10-bit TF32 mantissa, float32
accumulation as mma.sync does): plain TF32, and the 3 x TF32 split a_hi b_hi + a_hi b_lo + a_lo b_hi, against float64.
The matcher's gates are 1e-4 relative on the confidence and 0.02 px on the sub-pixel fit; the float32 SIMT path sits at
a few 1e-7 per stage."""
import numpy as np


def tf32(x):
    b = np.asarray(x, dtype=np.float32).view(np.uint32)
    b = (b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)          # round to 10 mantissa bits
    return b.view(np.float32)


def gemm_f32acc(a, b):
    return (a.astype(np.float32)[:, :, None] * b.astype(np.float32)[None, :, :]).sum(axis=1, dtype=np.float32)


rng = np.random.default_rng(0)
n = 32
k = np.arange(n)
w = np.exp(-2j * np.pi * np.outer(k, k) / n)
wr = np.block([[w.real, -w.imag], [w.imag, w.real]])            # 64 x 64
x = rng.standard_normal((n, 32)) + 1j * rng.standard_normal((n, 32))
xr = np.concatenate([x.real, x.imag], axis=0)                    # 64 x 32
ref = wr @ xr
a32, b32 = wr.astype(np.float32), xr.astype(np.float32)
a_hi, b_hi = tf32(a32), tf32(b32)
a_lo, b_lo = tf32(a32 - a_hi), tf32(b32 - b_hi)
one = gemm_f32acc(a_hi, b_hi)
three = gemm_f32acc(a_hi, b_hi) + gemm_f32acc(a_hi, b_lo) + gemm_f32acc(a_lo, b_hi)
f32 = gemm_f32acc(a32, b32)
scale = np.abs(ref).max()
print('radix-32 stage, max error relative to the largest output:')
print(f'  1 x TF32      {np.abs(one - ref).max() / scale:.2e}')
print(f'  3 x TF32      {np.abs(three - ref).max() / scale:.2e}')
print(f'  float32 FMA   {np.abs(f32 - ref).max() / scale:.2e}')
