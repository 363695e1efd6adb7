"""Times every tile configuration of var_gemm_kernel (gple_tune_variance_gemm) on the current GPU.
Usage (GPU box): python profiles/tune_var_gemm.py > gpurun_out/tune_var_gemm.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_liouville_equation_b200 import _lib as L

NAMES = ["BK16 S4 2x4 (64x32)", "BK32 S3 2x4 (64x32)", "BK16 S4 4x4 (32x32)", "BK32 S3 4x4 (32x32)", "BK16 S5 2x4 (64x32)", "BK16 S4 2x8 (64x16)", "BK16 S4 4x2 (32x64)"]
ctx = L.Context(0)
dmma, dfma = ctx.fp64_peak()
print(f"measured peaks: DMMA {dmma:.2f} TFLOP/s, DFMA {dfma:.2f} TFLOP/s")
print(f"register-tile DMMA ceiling (8 warps/SM, 8x4 tile, changing operands, no memory): {ctx.dmma_tile_peak():.2f} TFLOP/s")
for n in (2048, 4096, 8192):
    rows = 148 * 128
    flops = rows * n * (n + 128.0)
    for v, name in enumerate(NAMES):
        ms = min(ctx.tune_variance_gemm(v, rows, n, 3 if n > 4096 else 5) for _ in range(2))
        print(f"n={n:5d} variant {v} {name:22s}: {ms:8.3f} ms  {flops / ms / 1e9:6.2f} TFLOP/s  ({flops / ms / 1e9 / dmma * 100:5.1f}% of DMMA peak)")
