"""Tensor-core (tcgen05, 3xTF32) eval chains against the FFMA layer kernels and the oracle: generate, generate -> classify
-> filter -> compact, classifier / encoder forward.  Covers ragged tails (n not a multiple of the 64-row tile), a feature
count that is not a multiple of 8 (zero-padded contraction), the F = 30 test shape of the reference, Philox rows keyed by
the global row index, and the compile-time-class-count filter kernels for several K."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cvae_gan_oracle as O  # noqa: E402
from tests import parity as P  # noqa: E402

pytestmark = pytest.mark.gpu


def _pair_engines(F_, K, seed):
    """The same parameters in a tensor-core engine and an FFMA-only engine (CVG_DISABLE_TC is read at creation)."""
    os.environ.pop("CVG_DISABLE_TC", None)
    orc, eng_tc, g = P.make_pair(F_, K, 256, seed=seed)
    os.environ["CVG_DISABLE_TC"] = "1"
    try:
        _, eng_ff, _ = P.make_pair(F_, K, 256, seed=seed)
    finally:
        os.environ.pop("CVG_DISABLE_TC", None)
    return orc, eng_tc, eng_ff, g


@pytest.mark.parametrize("mode", ["128", "64", "pp"])
@pytest.mark.parametrize("F_,K,n", [(10, 5, 1), (10, 5, 63), (10, 5, 65), (10, 5, 129), (10, 5, 1000), (30, 5, 333), (7, 3, 200),
                                    (16, 9, 129), (40, 20, 70)])
def test_tc_chain_matches_ffma_and_oracle(F_, K, n, mode, monkeypatch):
    """mode: 128-row tiles, 64-row tiles (used when F > 32 or K > 16), or two 64-row tiles in flight (ping-pong)."""
    monkeypatch.setenv("CVG_TC_MODE", mode)
    orc, eng_tc, eng_ff, g = _pair_engines(F_, K, seed=50 + F_ + K)
    label = K - 1
    z = torch.randn(n, 128, generator=g)
    with torch.no_grad():
        ref = O.generator_forward(orc.sd["generator"], z, label, False)
    a = eng_tc.generate(label, n, z=z.cuda())
    b = eng_ff.generate(label, n, z=z.cuda())
    ok, worst, mx = P.close(a, ref)
    assert ok, ("tc vs oracle", worst, mx)
    assert (a - b).abs().max().item() < 2e-5, "tensor-core (3xTF32) and FFMA generator outputs differ beyond fp32 round-off"
    # fused generate -> classify -> filter -> compact
    xo, lo, keep_o = orc.generate_filter_stream(label, z, 0.3)
    xg, idx, cnt, lg, kg = eng_tc.generate_filter(label, n, 0.3, z=z.cuda(), want_logits=True, want_keep=True)
    ok, worst, mx = P.close(lg, lo)
    assert ok, ("logits", worst, mx)
    assert torch.equal(kg.bool().cpu(), O.filter_logits(lg.cpu(), label, 0.3))       # bit-exact on identical logits
    c = int(cnt.item())
    assert c == int(kg.sum())
    order = torch.argsort(idx[:c])
    assert torch.equal(idx[:c][order].cpu(), torch.nonzero(kg.cpu()).flatten())
    assert torch.equal(xg[:c][order].cpu(), a.cpu()[kg.bool().cpu()])                 # rows moved verbatim
    # single-network forwards
    x = torch.rand(n, F_, generator=g)
    lt, lf = eng_tc.classifier_forward(x.cuda()), eng_ff.classifier_forward(x.cuda())
    assert (lt - lf).abs().max().item() < 5e-5
    mt, vt = eng_tc.encoder_forward(x.cuda(), label)
    mf, vf = eng_ff.encoder_forward(x.cuda(), label)
    assert (mt - mf).abs().max().item() < 5e-5 and (vt - vf).abs().max().item() < 5e-5
    eng_tc.close()
    eng_ff.close()


def test_tc_empty_and_wide_inputs():
    """n = 0 is a no-op; F > 32 / K > 16 take the 64-row kernel whatever CVG_TC_MODE says."""
    _, eng, g = P.make_pair(40, 20, 64, seed=62)
    assert eng.generate(3, 0).shape == (0, 40)
    xg, idx, cnt, _, _ = eng.generate_filter(3, 0, 0.1)
    assert int(cnt.item()) == 0
    x = torch.rand(5, 40, generator=g).cuda()
    assert eng.classifier_forward(x).shape == (5, 20)
    eng.close()


def test_tc_philox_rows_do_not_depend_on_tiling():
    _, eng, _ = P.make_pair(10, 5, 256, seed=61)
    a = eng.generate(2, 1000, seed=11, row_offset=500)
    b1 = eng.generate(2, 37, seed=11, row_offset=500)
    b2 = eng.generate(2, 963, seed=11, row_offset=537)
    assert torch.equal(a, torch.cat([b1, b2]))
    xg, idx, cnt, _, kg = eng.generate_filter(2, 1000, 0.0, seed=11, row_offset=500, want_keep=True)
    c = int(cnt.item())
    order = torch.argsort(idx[:c])
    assert torch.equal(xg[:c][order], a[kg.bool()])
    assert torch.equal(idx[:c][order].cpu(), 500 + torch.nonzero(kg.cpu()).flatten())
    eng.close()


@pytest.mark.parametrize("K", [2, 3, 4, 5, 8, 12, 13, 32])
def test_filter_compact_all_class_counts(K):
    """Streaming filter (compile-time class count for K <= 12, generic kernel above) == torch softmax/max on the same logits."""
    from cvae_gan_b200.engine import Engine
    eng = Engine(10, min(K, 32), 128, max_batch=64)
    g = torch.Generator().manual_seed(K)
    n = 5000
    x = torch.rand(n, 10, generator=g).cuda()
    lg = (2.0 * torch.randn(n, K, generator=g)).cuda()
    lg[::7, 1] = lg[::7, 0]                      # exact ties: the first maximal index must win
    label, thr = 0, 1.0 / K + 0.05
    p = torch.softmax(lg, 1)
    m, i = p.max(1)
    keep = (m > thr) & (i == label)
    xo, idx, cnt = eng.filter_compact(x, lg, label, thr, row_offset=100)
    c = int(cnt.item())
    assert c == int(keep.sum())
    order = torch.argsort(idx[:c])
    assert torch.equal(idx[:c][order].cpu(), 100 + torch.nonzero(keep.cpu()).flatten())
    assert torch.equal(xo[:c][order], x[keep])
    assert torch.equal(eng.filter_logits(lg, label, thr).cpu(), keep.cpu())
    eng.close()
