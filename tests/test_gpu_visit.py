"""Label-visit entry point (cvg_visit): eager == CUDA-graph replay == the per-step C-ABI calls with the same
Philox key/counters; `CVAEGAN.fit` end to end; generation/filter semantics of the reference class."""
import pytest
import torch

pytestmark = pytest.mark.gpu

F_, K, B = 10, 5, 256


def _engine(seed=3):
    from cvae_gan_b200 import models
    from cvae_gan_b200.engine import Engine
    torch.manual_seed(seed)
    eng = Engine(F_, K, 128, max_batch=B)
    mods = [models.CVAEGANEncoderModel(F_, K), models.CVAEGANGeneratorModel(128, K, F_),
            models.CVAEGANDiscriminatorModel(F_, K), models.CVAEGANClassifierModel(F_, K)]
    for net, m in enumerate(mods):
        eng.load_state(net, m.state_dict())
    return eng


def _flat(eng):
    return torch.cat([eng.params[n].clone() for n in range(4)] + [eng.state[n].clone() for n in range(4)])


def test_visit_eager_equals_graph_equals_steps():
    g = torch.Generator().manual_seed(0)
    rows = torch.rand(5000, F_, generator=g).cuda()
    loops = (2, 2, 2)
    outs = []
    for mode in ("eager", "graph", "steps"):
        eng = _engine()
        eng.ctl_set(seed=77, counter=10, lambda_class=0.25)
        loss = torch.zeros(sum(loops), 4, device="cuda")
        if mode == "eager":
            for _ in range(2):
                eng.visit(2, B, class_rows=rows, loops=loops, loss_out=loss)
        elif mode == "graph":
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                eng.visit(2, B, class_rows=rows, loops=loops, loss_out=loss)
            gr.replay()
            gr.replay()
        else:
            c = 10
            for _ in range(2):
                i = 0
                for kind, reps in zip("dcg", loops):
                    for _ in range(reps):
                        x = eng.sample_rows(rows, B, seed=77, counter=c)
                        if kind == "d":
                            eng.step_d(x, 2, seed=77, counter=c + 1, loss_out=loss[i])
                        elif kind == "c":
                            eng.step_c(x, 2, seed=77, counter=c + 1, loss_out=loss[i])
                        else:
                            eng.step_g(x, 2, 0.25, seed=77, counter=c + 1, loss_out=loss[i])
                        c += 2
                        i += 1
        torch.cuda.synchronize()
        outs.append((loss.clone(), _flat(eng), [eng.get_adam_step(n) for n in range(4)]))
        eng.close()
    for other in outs[1:]:
        assert other[2] == outs[0][2] == [4, 4, 4, 4]
        assert torch.allclose(other[0], outs[0][0], rtol=2e-3, atol=2e-4)
        # float atomics make summation order vary between runs; Adam turns that into <= lr-sized differences
        assert torch.allclose(other[1], outs[0][1], rtol=1e-3, atol=4 * 2e-4 * 1.5)


def test_cgan_visit_equals_steps_and_fit():
    """CVG_STEP_PRIOR_ONLY visits (CGAN): graph replay == per-step calls; `CGAN().fit` end to end."""
    from cvae_gan_b200._lib import STEP_PRIOR_ONLY
    g = torch.Generator().manual_seed(0)
    rows = torch.rand(5000, F_, generator=g).cuda()
    loops = (2, 2, 2)
    outs = []
    for mode in ("graph", "steps"):
        eng = _engine()
        eng.ctl_set(seed=78, counter=20, lambda_class=0.25)
        loss = torch.zeros(sum(loops), 4, device="cuda")
        if mode == "graph":
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                eng.visit(1, B, class_rows=rows, loops=loops, loss_out=loss, flags=STEP_PRIOR_ONLY)
            gr.replay()
            gr.replay()
        else:
            c = 20
            for _ in range(2):
                i = 0
                for kind, reps in zip("dcg", loops):
                    for _ in range(reps):
                        if kind == "g":
                            eng.step_g_prior(B, 1, 0.25, seed=78, counter=c + 1, loss_out=loss[i])
                        else:
                            x = eng.sample_rows(rows, B, seed=78, counter=c)
                            (eng.step_d if kind == "d" else eng.step_c)(x, 1, seed=78, counter=c + 1, loss_out=loss[i])
                        c += 2
                        i += 1
        torch.cuda.synchronize()
        outs.append((loss.clone(), _flat(eng), [eng.get_adam_step(n) for n in range(4)]))
        eng.close()
    assert outs[0][2] == outs[1][2] == [0, 4, 4, 4]
    assert torch.allclose(outs[1][0], outs[0][0], rtol=2e-3, atol=2e-4)
    assert torch.allclose(outs[1][1], outs[0][1], rtol=1e-3, atol=4 * 2e-4 * 1.5)
    assert float(outs[0][0][4:, :2].abs().max()) == 0.0          # no reconstruction / KL terms

    import cvae_gan_b200 as cg
    from tests.parity import make_data
    x, y = make_data(F_, K, [300, 260, 64, 40, 300], seed=6)
    perm = torch.randperm(len(y), generator=torch.Generator().manual_seed(2))
    cg.datasets.tr_samples, cg.datasets.tr_labels = x[perm], y[perm]
    cg.datasets.feature_num, cg.datasets.label_num = F_, K
    cg.config.gan_config.batch_size, cg.config.gan_config.epochs = 64, 3
    torch.manual_seed(0)
    gan = cg.CGAN()
    assert not hasattr(gan, "encoder") and not hasattr(gan, "lambda_recon") and sorted(gan.loss_history) == ["adv_loss", "class_loss"]
    enc0 = gan.engine.params[0].clone()
    gan.fit(cg.datasets.TrDataset())
    assert [len(v) for v in gan.loss_history.values()] == [3, 3]
    assert all(abs(v) < 1e3 and v == v for vs in gan.loss_history.values() for v in vs)
    assert torch.equal(gan.engine.params[0], enc0)                                                  # no encoder in CGAN
    assert int(gan.generator.state_dict()["main_model.1.num_batches_tracked"]) == 3 * K * (5 + 5 + 3)
    assert not gan.generator.training and not gan.classifier.training
    s = gan.generate_samples(2, 50)
    assert s.shape == (50, F_) and s.device.type == "cpu" and float(s.min()) >= 0.0 and float(s.max()) <= 1.0
    q = gan.generate_qualified_samples(0, 20, confidence_threshold=0.0)
    assert q.numel() == 0 or q.shape[1] == F_
    assert sorted(gan.state_dict()) == ["classifier", "discriminator", "generator"]


def test_cvaegan_fit_generate_and_filter():
    import cvae_gan_b200 as cg
    from tests.parity import make_data
    x, y = make_data(F_, K, [300, 260, 64, 40, 300], seed=5)
    perm = torch.randperm(len(y), generator=torch.Generator().manual_seed(1))
    cg.datasets.tr_samples, cg.datasets.tr_labels = x[perm], y[perm]
    cg.datasets.feature_num, cg.datasets.label_num = F_, K
    cg.config.gan_config.batch_size, cg.config.gan_config.epochs = 64, 3
    torch.manual_seed(0)
    gan = cg.CVAEGAN()
    init = {k: v.clone() for k, v in gan.generator.state_dict().items()}
    gan.fit(cg.datasets.TrDataset())
    # key order = first occurrence, rows partitioned like cvae_gan.py:238-245
    first = []
    for lab in y[perm].tolist():
        if lab not in first:
            first.append(lab)
    assert list(gan.samples.keys()) == first
    for lab in range(K):
        assert torch.equal(gan.samples[lab].cpu(), x[perm][y[perm] == lab])
    assert all(len(v) == 3 for v in gan.loss_history.values())
    assert all(abs(v) < 1e3 and v == v for vs in gan.loss_history.values() for v in vs)
    sd = gan.generator.state_dict()
    assert not torch.equal(sd["main_model.0.weight"].cpu(), init["main_model.0.weight"].cpu())     # trained in place
    assert int(sd["main_model.1.num_batches_tracked"]) == 3 * K * (5 + 5 + 2 * 3)               # 16 per visit
    assert int(gan.encoder.state_dict()["encoder.1.num_batches_tracked"]) == 3 * K * 3
    assert not gan.generator.training and not gan.classifier.training
    # generation surface
    s = gan.generate_samples(1, 37)
    assert s.shape == (37, F_) and s.device.type == "cpu" and float(s.min()) >= 0 and float(s.max()) <= 1
    # after 3 epochs the classifier is barely trained: find a label it actually predicts for generated rows
    got_any = False
    for lab in range(K):
        q = gan.generate_qualified_samples(lab, 25, 0.0)
        if q.numel() == 0:
            continue                                              # 20 empty chunks of 10: patience ran out
        got_any = True
        assert q.shape[1] == F_ and 0 < q.shape[0] <= 25 and q.device.type == "cpu"
        with torch.no_grad():
            logits = gan.engine.classifier_forward(q.cuda()).cpu()
        assert (logits.argmax(1) == lab).all()                    # every returned row passes the filter
    assert got_any
    assert gan.classifier.training                                # reference quirk (cvae_gan.py:363)
    empty = gan.generate_qualified_samples(0, 5, 1.0)             # nothing has max prob > 1: patience runs out
    assert empty.numel() == 0
    with pytest.raises(ValueError):
        gan.reconstruct_samples(x[:4], y[:4])
    assert gan.reconstruct(x[:8], 2).shape == (8, F_)
    # same seed, graphs off -> same training (up to float-atomic ordering)
    torch.manual_seed(0)
    gan2 = cg.CVAEGAN()
    gan2.use_cuda_graphs = False
    gan2.fit(cg.datasets.TrDataset())
    for k in gan.loss_history:
        assert gan.loss_history[k] == pytest.approx(gan2.loss_history[k], rel=5e-3, abs=5e-4)
