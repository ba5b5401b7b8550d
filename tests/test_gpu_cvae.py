"""SURVEY 8 f4, sibling trainer CVAE (/root/reference/src/cvae.py): the encoder/generator step (engine flag CVG_STEP_CVAE,
train.cu step_g_cvae) against the oracle restatement `step_g_cvae` (pinned on tests/golden/ref_cvae_*.npz, which were made
from the unmodified reference); its classifier step is the CVAE-GAN's (tests/test_gpu_parity.py).  Then label visits
(graph replay == per-step calls) and the host class `CVAE` end to end.  Run on a B200 with `pytest -m gpu`."""
import os

import pytest
import torch

from tests import parity as P

pytestmark = pytest.mark.gpu

# gan_config.cvae_config (gan_config.py:51-56)
CVAE_CFG = dict(lambda_recon=1.0, lambda_kl=0.01, lambda_class=0.1)


@pytest.fixture(autouse=True)
def _cpu_threads():
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    yield


@pytest.mark.parametrize("F_,K,B", [(10, 5, 64), (10, 4, 333), (30, 5, 128), (10, 5, 4096)])
@pytest.mark.parametrize("lam", [0.0, 0.25])
def test_cvae_step_losses_and_gradients(F_, K, B, lam):
    """Losses {recon, kl, 0, class} and every encoder / generator gradient of one step without update; lam = 0 is the first 200
    epochs (class loss reported, not back-propagated, cvae.py:141-147), lam != 0 sends the classifier's input gradient through
    the RECONSTRUCTION into G and E."""
    orc, eng, g = P.make_pair(F_, K, B, seed=13 + B, **CVAE_CFG)
    x, y = P.make_data(F_, K, [B] * K, seed=4)
    xb = x[y == (K - 2)][:B].contiguous()
    eng.zero_grads()
    crit0, clf0 = eng.params[2].clone(), eng.params[3].clone()
    twin = orc.twin64() if B >= 4096 else None
    ref, got, grads = P.run_step("v", orc, eng, xb, K - 2, g, lambda_class=lam, update=False, twin=twin)
    assert P.losses_close(ref, got), (ref, got)
    report = []
    P.compare_grads(eng, orc, ["encoder", "generator"], grads, report, grads64=P.run_step.last_twin_grads)
    report = [r for r in report if not any(r[0].endswith(k) for ks in P.PRE_BN_BIASES.values() for k in ks)]
    P.assert_report(report, "CVAE encoder/generator step gradients")
    # no critic in CVAE; the classifier is read, never updated by this step
    assert float(eng.grads[2].abs().max()) == 0.0 and torch.equal(eng.params[2], crit0)
    assert float(eng.grads[3].abs().max()) == 0.0 and torch.equal(eng.params[3], clf0)
    rep2 = []
    P.compare_state(eng, orc, rep2, loose_prebn_atol=1e-3)       # BatchNorm running statistics of E and G (one pass each)
    P.assert_report(rep2, "CVAE encoder/generator step state")
    eng.close()


def test_cvae_two_label_visits_trajectory():
    """CVAE.fit's step sequence (5 classifier + 3 encoder/generator steps per label visit, cvae.py:86-166) with Adam updates."""
    F_, K, B = 10, 5, 256
    orc, eng, g = P.make_pair(F_, K, B, seed=29, **CVAE_CFG)
    x, y = P.make_data(F_, K, [400, 256, 100, 300, 300], seed=2)
    orc.divide_samples(x, y)
    twin = orc.twin64()
    for label in (1, 4):
        for kind, reps in (("c", 5), ("v", 3)):
            for _ in range(reps):
                idx = torch.randperm(len(orc.samples[label]), generator=g)[:B]
                xb = orc.samples[label][idx].contiguous()
                ref, got, _ = P.run_step(kind, orc, eng, xb, label, g, lambda_class=0.25, update=True, twin=twin)
                assert P.losses_close(ref, got, rtol=2e-3, atol=5e-4), (kind, ref, got)
    report = []
    P.compare_state(eng, orc, report, loose_prebn_atol=6 * 2e-4 * 1.5, atol_frac=2e-3, outlier_frac=2e-3, hard_atol=10 * 2e-4,
                    twin=twin)
    P.assert_report(report, "parameters after two CVAE label visits")
    assert [eng.get_adam_step(n) for n in range(4)] == [6, 6, 0, 10]
    eng.close()


def test_cvae_visit_equals_steps_and_fit():
    """CVG_STEP_CVAE visits: graph replay == per-step calls; the flag combinations the header rules out fail; `CVAE().fit`,
    generation, the confidence filter and `reconstruct_samples` end to end."""
    from cvae_gan_b200._lib import STEP_CVAE, STEP_PRIOR_ONLY, CvgError
    from tests.test_gpu_visit import _engine, _flat
    F_, K, B = 10, 5, 256
    g = torch.Generator().manual_seed(0)
    rows = torch.rand(5000, F_, generator=g).cuda()
    loops = (0, 2, 2)
    outs = []
    for mode in ("graph", "steps"):
        eng = _engine()
        eng.ctl_set(seed=79, counter=30, lambda_class=0.25)
        loss = torch.zeros(sum(loops), 4, device="cuda")
        if mode == "graph":
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                eng.visit(3, B, class_rows=rows, loops=loops, loss_out=loss, flags=STEP_CVAE)
            gr.replay()
            gr.replay()
        else:
            c = 30
            for _ in range(2):
                i = 0
                for kind, reps in zip("dcg", loops):
                    for _ in range(reps):
                        x = eng.sample_rows(rows, B, seed=79, counter=c)
                        if kind == "g":
                            eng.step_g_cvae(x, 3, 0.25, seed=79, counter=c + 1, loss_out=loss[i])
                        else:
                            eng.step_c(x, 3, seed=79, counter=c + 1, loss_out=loss[i])
                        c += 2
                        i += 1
        torch.cuda.synchronize()
        outs.append((loss.clone(), _flat(eng), [eng.get_adam_step(n) for n in range(4)]))
        if mode == "steps":
            with pytest.raises(CvgError):
                eng.visit(3, B, class_rows=rows, loops=(1, 1, 1), flags=STEP_CVAE)              # a CVAE visit has no critic steps
            with pytest.raises(CvgError):
                eng.visit(3, B, class_rows=rows, loops=loops, flags=STEP_CVAE | STEP_PRIOR_ONLY)
        eng.close()
    assert outs[0][2] == outs[1][2] == [4, 4, 0, 4]
    assert torch.allclose(outs[1][0], outs[0][0], rtol=2e-3, atol=2e-4)
    assert torch.allclose(outs[1][1], outs[0][1], rtol=1e-3, atol=4 * 2e-4 * 1.5)
    assert float(outs[0][0][2:, 2].abs().max()) == 0.0           # no adversarial term
    assert float(outs[0][0][2:, 0].min()) > 0.0 and float(outs[0][0][2:, 1].min()) > 0.0

    import cvae_gan_b200 as cg
    from tests.parity import make_data
    x, y = make_data(F_, K, [300, 260, 64, 40, 300], seed=6)
    perm = torch.randperm(len(y), generator=torch.Generator().manual_seed(2))
    cg.datasets.tr_samples, cg.datasets.tr_labels = x[perm], y[perm]
    cg.datasets.feature_num, cg.datasets.label_num = F_, K
    cg.config.gan_config.batch_size, cg.config.gan_config.epochs = 64, 3
    torch.manual_seed(0)
    gan = cg.CVAE()
    assert not hasattr(gan, "discriminator") and not hasattr(gan, "lambda_adv")
    assert sorted(gan.loss_history) == ["class_loss", "kl_loss", "recon_loss"]
    assert (gan.lambda_recon, gan.lambda_kl, gan.lambda_class) == (1.0, 0.01, 0.1)
    crit0 = gan.engine.params[2].clone()
    gan.fit(cg.datasets.TrDataset())
    assert [len(v) for v in gan.loss_history.values()] == [3, 3, 3]
    assert all(abs(v) < 1e3 and v == v for vs in gan.loss_history.values() for v in vs)
    assert torch.equal(gan.engine.params[2], crit0)                                                 # no critic in CVAE
    assert int(gan.generator.state_dict()["main_model.1.num_batches_tracked"]) == 3 * K * (5 + 3)   # cvae.py:99,131
    assert int(gan.encoder.state_dict()["encoder.1.num_batches_tracked"]) == 3 * K * 3
    assert not gan.generator.training and not gan.classifier.training and not gan.encoder.training
    s = gan.generate_samples(2, 50)
    assert s.shape == (50, F_) and s.device.type == "cpu" and float(s.min()) >= 0.0 and float(s.max()) <= 1.0
    q = gan.generate_qualified_samples(0, 20, confidence_threshold=0.0)
    assert q.numel() == 0 or q.shape[1] == F_
    # reconstruct_samples works in the reference's CVAE (cvae.py:300-319): mixed labels, eval-mode E and G, both left in train mode
    rec = gan.reconstruct_samples(x[::40], y[::40])
    assert rec.shape == (len(x[::40]), F_) and rec.device.type == "cpu" and float(rec.min()) >= 0.0 and float(rec.max()) <= 1.0
    assert gan.encoder.training and gan.generator.training
    # against the engine's own pieces, row by row (eval mode: rows are independent): same mu / logvar -> same decoder output
    lab = int(y[0])
    mu, lv = gan.engine.encoder_forward(x[:16].cuda(), lab)
    direct = gan.engine.generate(lab, 16, z=mu.contiguous(), train_mode=False)
    assert direct.shape == (16, F_)
    assert sorted(gan.state_dict()) == ["classifier", "encoder", "generator"]
