"""SURVEY 8 f4, sibling trainer VAE-GAN (/root/reference/src/vae_gan.py): unconditional encoder / generator / critic
(`CvgConfig.unconditional`: the CVAE-GAN's stacks without the one-hot label columns), critic step and encoder/generator step
against the oracle restatement (`step_d(x, None, ...)`, `step_g_vaegan`; pinned on tests/golden/ref_vaegan.npz, which was made
from the unmodified reference), an epoch as one engine visit, and the host class `VAEGAN` end to end.
Run on a B200 with `pytest -m gpu`."""
import os

import pytest
import torch

from oracle import cvae_gan_oracle as O
from tests import parity as P

pytestmark = pytest.mark.gpu

# gan_config.vae_gan_config (gan_config.py:33-38)
VAEGAN_CFG = dict(lambda_recon=1.0, lambda_kl=0.01, lambda_adv=0.1, unconditional=True)


@pytest.fixture(autouse=True)
def _cpu_threads():
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    yield


@pytest.mark.parametrize("F_,B", [(10, 64), (10, 333), (30, 128), (10, 4096)])
@pytest.mark.parametrize("kind", ["d", "u"])
def test_vaegan_step_losses_and_gradients(kind, F_, B):
    K = 5        # only sizes the (unused) classifier
    orc, eng, g = P.make_pair(F_, K, B, seed=17 + B, **VAEGAN_CFG)
    assert eng.tables[0]["encoder.0.weight"][1] == (256, F_) and eng.tables[1]["main_model.0.weight"][1] == (256, 128)
    assert eng.tables[2]["discriminator_network.0.parametrizations.weight.original"][1] == (256, F_)
    x, _ = P.make_data(F_, 1, [B], seed=5)
    eng.zero_grads()
    clf0 = eng.params[3].clone()
    twin = orc.twin64() if B >= 4096 else None
    ref, got, grads = P.run_step(kind, orc, eng, x.contiguous(), None, g, update=False, twin=twin)
    assert P.losses_close(ref, got), (ref, got)
    report = []
    nets = {"d": ["discriminator"], "u": ["encoder", "generator"]}[kind]
    P.compare_grads(eng, orc, nets, grads, report, grads64=P.run_step.last_twin_grads)
    report = [r for r in report if not any(r[0].endswith(k) for ks in P.PRE_BN_BIASES.values() for k in ks)]
    P.assert_report(report, f"VAE-GAN step_{kind} gradients")
    assert float(eng.grads[3].abs().max()) == 0.0 and torch.equal(eng.params[3], clf0)      # no classifier in a VAE-GAN
    rep2 = []
    P.compare_state(eng, orc, rep2, loose_prebn_atol=1e-3)
    P.assert_report(rep2, f"VAE-GAN step_{kind} state")
    eng.close()


def test_vaegan_two_epochs_trajectory():
    """VAEGAN.fit's step sequence (per epoch 5 critic + 3 encoder/generator steps on batches from all rows) with Adam updates."""
    F_, B = 10, 256
    orc, eng, g = P.make_pair(F_, 5, B, seed=41, **VAEGAN_CFG)
    x, _ = P.make_data(F_, 3, [400, 300, 300], seed=2)
    twin = orc.twin64()
    for _ in range(2):
        for kind, reps in (("d", 5), ("u", 3)):
            for _ in range(reps):
                xb = x[torch.randperm(len(x), generator=g)[:B]].contiguous()
                ref, got, _ = P.run_step(kind, orc, eng, xb, None, g, update=True, twin=twin)
                assert P.losses_close(ref, got, rtol=2e-3, atol=5e-4), (kind, ref, got)
    report = []
    P.compare_state(eng, orc, report, loose_prebn_atol=6 * 2e-4 * 1.5, atol_frac=2e-3, outlier_frac=2e-3, hard_atol=10 * 2e-4,
                    twin=twin)
    P.assert_report(report, "parameters after two VAE-GAN epochs")
    assert [eng.get_adam_step(n) for n in range(4)] == [6, 6, 10, 0]
    eng.close()


def test_vaegan_epoch_visit_equals_steps_and_fit():
    """An epoch as one engine visit (label 0, no classifier steps, lambda_class 0): graph replay == per-step calls; then
    `VAEGAN().fit`, generation and reconstruction end to end."""
    from cvae_gan_b200 import models
    from cvae_gan_b200._lib import VISIT_LAMBDA_ZERO
    from cvae_gan_b200.engine import Engine
    F_, B = 10, 256
    g = torch.Generator().manual_seed(0)
    rows = torch.rand(5000, F_, generator=g).cuda()
    loops = (2, 0, 2)

    def engine():
        torch.manual_seed(3)
        eng = Engine(F_, 4, 128, max_batch=B, lambda_recon=1.0, lambda_kl=0.01, lambda_adv=0.1, unconditional=True)
        mods = [models.VAEGANEncoderModel(F_, 128), models.VAEGANGeneratorModel(128, F_), models.VAEGANDiscriminatorModel(F_),
                models.CVAEGANClassifierModel(F_, 4)]
        for net, m in enumerate(mods):
            eng.load_state(net, m.state_dict())
        return eng

    outs = []
    for mode in ("graph", "steps"):
        eng = engine()
        eng.ctl_set(seed=80, counter=40, lambda_class=0.0)
        loss = torch.zeros(sum(loops), 4, device="cuda")
        if mode == "graph":
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                eng.visit(0, B, class_rows=rows, loops=loops, loss_out=loss, flags=VISIT_LAMBDA_ZERO)
            gr.replay()
            gr.replay()
        else:
            c = 40
            for _ in range(2):
                i = 0
                for kind, reps in zip("dcg", loops):
                    for _ in range(reps):
                        x = eng.sample_rows(rows, B, seed=80, counter=c)
                        if kind == "d":
                            eng.step_d(x, 0, seed=80, counter=c + 1, loss_out=loss[i])
                        else:
                            eng.step_g(x, 0, 0.0, seed=80, counter=c + 1, loss_out=loss[i])
                        c += 2
                        i += 1
        torch.cuda.synchronize()
        flat = torch.cat([eng.params[n].clone() for n in range(4)] + [eng.state[n].clone() for n in range(4)])
        outs.append((loss.clone(), flat, [eng.get_adam_step(n) for n in range(4)]))
        eng.close()
    assert outs[0][2] == outs[1][2] == [4, 4, 4, 0]
    assert torch.allclose(outs[1][0][:, :3], outs[0][0][:, :3], rtol=2e-3, atol=2e-4)
    assert torch.allclose(outs[1][1], outs[0][1], rtol=1e-3, atol=4 * 2e-4 * 1.5)

    import cvae_gan_b200 as cg
    x, y = P.make_data(F_, 4, [300, 260, 64, 40], seed=6)
    perm = torch.randperm(len(y), generator=torch.Generator().manual_seed(2))
    cg.datasets.tr_samples, cg.datasets.tr_labels = x[perm], y[perm]
    cg.datasets.feature_num, cg.datasets.label_num = F_, 4
    cg.config.gan_config.batch_size, cg.config.gan_config.epochs = 64, 4
    torch.manual_seed(0)
    gan = cg.VAEGAN()
    assert not hasattr(gan, "classifier") and not hasattr(gan, "lambda_class")
    assert sorted(gan.loss_history) == ["adv_loss", "kl_loss", "recon_loss"]
    assert (gan.lambda_recon, gan.lambda_kl, gan.lambda_adv) == (1.0, 0.01, 0.1)
    assert gan.encoder.state_dict()["encoder.0.weight"].shape == (256, F_)
    gan.fit(cg.datasets.TrDataset())
    assert torch.equal(gan.samples.cpu(), x[perm])                                            # all rows, labels ignored
    assert [len(v) for v in gan.loss_history.values()] == [4, 4, 4]
    assert all(abs(v) < 1e3 and v == v for vs in gan.loss_history.values() for v in vs)
    assert int(gan.generator.state_dict()["main_model.1.num_batches_tracked"]) == 4 * (5 + 2 * 3)   # vae_gan.py:87,119,122
    assert int(gan.encoder.state_dict()["encoder.1.num_batches_tracked"]) == 4 * 3
    assert not gan.generator.training and not gan.encoder.training and not gan.discriminator.training
    s = gan.generate_samples(50)
    assert s.shape == (50, F_) and s.device.type == "cpu" and float(s.min()) >= 0.0 and float(s.max()) <= 1.0
    rec = gan.reconstruct_samples(x[:20])
    assert rec.shape == (20, F_) and float(rec.min()) >= 0.0 and float(rec.max()) <= 1.0
    assert gan.encoder.training and gan.generator.training                                    # reference quirk (vae_gan.py:258-259)
    # the eval-mode chains against the oracle on the trained state
    st = {n: {k: v.detach().cpu() for k, v in getattr(gan, n).state_dict().items()} for n in ("encoder", "generator", "discriminator")}
    st["classifier"] = {k: v.detach().cpu() for k, v in gan._classifier_module.state_dict().items()}
    orc = O.OracleCVAEGAN(F_, 4, O.OracleConfig(unconditional=True)).load_state(st)
    z = torch.randn(64, 128, generator=g)
    xo = O.generator_forward(orc.sd["generator"], z, None, False, None)
    ok, worst, mx = P.close(gan.engine.generate(0, 64, z=z.cuda(), train_mode=False), xo.detach())
    assert ok, ("generator", worst, mx)
    mu_o, lv_o = O.encoder_forward(orc.sd["encoder"], x[:32], None, False, None)
    mu, lv = gan.engine.encoder_forward(x[:32].cuda(), 0)
    assert P.close(mu, mu_o.detach())[0] and P.close(lv, lv_o.detach())[0]
    assert sorted(gan.state_dict()) == ["discriminator", "encoder", "generator"]
