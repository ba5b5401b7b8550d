"""Diagnostic (GPU): where do the out-of-tolerance gradient entries of the widened model at batch 1000 sit?  Reproduces
tests/test_gpu_wide.py::test_wide_step_losses_and_gradients[hidden1-1000-v] and prints, for every encoder / generator gradient
tensor with entries outside 1e-3 of the tensor scale, how many entries miss, in which rows / columns, and how large the misses are -
a LeakyReLU derivative flipped by a pre-activation within round-off of zero shows up as ONE feature: one entry of the following
BatchNorm's bias gradient and one row of the Linear's weight gradient."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity as P  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "v"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    orc, eng, g = P.make_pair(10, 5, B, seed=31 + B, hidden=(1024, 512, 256))
    x, y = P.make_data(10, 5, [B] * 5, seed=7)
    xb = x[y == 1][:B].contiguous()
    eng.zero_grads()
    ref, got, grads = P.run_step(kind, orc, eng, xb, 1, g, lambda_class=0.25, update=False, twin=orc.twin64())
    print("losses", ref, got)
    g64 = P.run_step.last_twin_grads
    for name in ("encoder", "generator"):
        i = P.NETS.index(name)
        for j, (key, g_ref) in enumerate(zip(orc.param_keys(name), grads[name])):
            a = eng.view(i, key, "grads").double().cpu()
            b = g64[name][j]
            scale = float(b.abs().max())
            err = (a - b).abs()
            bad = err > 1e-3 * b.abs() + 1e-3 * scale
            if int(bad.sum()) == 0:
                continue
            idx = bad.nonzero()
            rows = sorted(set(idx[:, 0].tolist()))
            cols = sorted(set(idx[:, 1].tolist())) if idx.shape[1] > 1 else []
            print(f"{name}/{key}: shape {tuple(a.shape)} scale {scale:.3g} bad {int(bad.sum())} max err {float(err.max()):.3g} "
                  f"rows {rows[:8]}{'...' if len(rows) > 8 else ''} ncols {len(cols)}; oracle32-vs-64 max {float((g_ref.double() - b).abs().max()):.3g}")
    # the suspect: pre-activations of the encoder's second BatchNorm output closest to zero (oracle float64)
    eng.close()


if __name__ == "__main__":
    main()
