"""Host logic of the four trainer classes on CPU (no GPU, no CUDA library calls): the engine is replaced by a recording stand-in
with the Engine's Python surface, so what is checked is exactly what the host classes decide - which label visits they issue in
which order, with which loop counts and flags (lambda_class schedule, sibling step flags), how the Philox counter and the
BatchNorm call counters advance, which loss columns reach `loss_history`, train / eval mode hand-over, and the refusal to train
with stale hyper-parameters.  Reference: src/cvae_gan.py:59-236, src/cgan.py:51-196, src/cvae.py:51-179, src/vae_gan.py:42-157.
The arithmetic behind `visit` is covered by the `-m gpu` parity tests."""
from collections import OrderedDict
from types import SimpleNamespace

import pytest
import torch

from oracle import cvae_gan_oracle as O


class FakeEngine:
    """Records the calls a host class makes; parameters live in flat CPU buffers with the real tensor tables' keys / shapes."""
    created = []

    def __init__(self, feature_num, label_num, z_size=128, max_batch=4096, *, lambda_recon=1.0, lambda_kl=0.1, lambda_adv=1.0,
                 g_lr=2e-4, d_lr=2e-4, c_lr=1e-4, world_size=1, rank=0, hidden=None, unconditional=False, **_):
        self.F, self.K, self.Z, self.max_batch = feature_num, label_num, z_size, max_batch
        self.world_size, self.rank, self.hidden, self.unconditional = world_size, rank, hidden, unconditional
        self.device = torch.device("cpu")
        self.cfg = SimpleNamespace(lambda_recon=lambda_recon, lambda_kl=lambda_kl, lambda_adv=lambda_adv, g_lr=g_lr, d_lr=d_lr, c_lr=c_lr)
        self.tables, self.params, self.state, self.grads, self.adam_m, self.adam_v = [], [], [], [], [], []
        for net in O.NETS:
            tab, po, so = OrderedDict(), 0, 0
            for key, shape, kind in O.tensor_table(net, feature_num, label_num, z_size, hidden, unconditional):
                if kind == "buffer_i64":
                    continue
                n = 1
                for s in shape:
                    n *= s
                if kind == "param":
                    tab[key] = (0, tuple(shape), po)
                    po += n
                else:
                    tab[key] = (1, tuple(shape), so)
                    so += n
            self.tables.append(tab)
            self.params.append(torch.zeros(po))
            self.grads.append(torch.ones(po))
            self.adam_m.append(torch.ones(po))
            self.adam_v.append(torch.ones(po))
            self.state.append(torch.zeros(max(so, 1)))
        self.adam_steps = [7, 7, 7, 7]
        self.calls = []
        FakeEngine.created.append(self)

    def view(self, net, key, which="params"):
        kind, shape, off = self.tables[net][key]
        n = 1
        for s in shape:
            n *= s
        buf = self.state[net] if kind == 1 else getattr(self, which)[net]
        return buf[off:off + n].view(shape)

    def load_state(self, net, sd):
        for key in self.tables[net]:
            self.view(net, key).copy_(sd[key].float().reshape(self.tables[net][key][1]))

    def set_adam_step(self, net, t):
        self.adam_steps[net] = t

    def ctl_set(self, seed=None, counter=None, lambda_class=None):
        self.calls.append(("ctl_set", seed, counter, lambda_class))

    def verify_replicas(self):
        self.calls.append(("verify_replicas",))

    def visit(self, label, batch_global, class_rows=None, x_batches=None, loops=(5, 5, 3), flags=0, loss_out=None):
        n = sum(loops)
        k = sum(1 for c in self.calls if c[0] == "visit")
        self.calls.append(("visit", label, batch_global, tuple(class_rows.shape), tuple(loops), flags))
        # the last step's row is what the host reads back: {recon, kl, adv, class} tagged with the visit number
        loss_out[n - 1] = torch.tensor([k + 0.1, k + 0.2, k + 0.3, k + 0.4])
        return loss_out


@pytest.fixture
def host(monkeypatch):
    import cvae_gan_b200 as cg
    from cvae_gan_b200 import cvae_gan, vae_gan
    FakeEngine.created.clear()
    monkeypatch.setattr(cvae_gan, "Engine", FakeEngine)
    monkeypatch.setattr(vae_gan, "Engine", FakeEngine)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    gc = cg.config.gan_config
    saved = (gc.epochs, gc.batch_size, gc.d_loop_num, gc.c_loop_num, gc.g_loop_num, gc.g_lr, cg.datasets.tr_samples,
             cg.datasets.tr_labels, cg.datasets.feature_num, cg.datasets.label_num)
    g = torch.Generator().manual_seed(1)
    y = torch.tensor([2, 2, 0, 1, 0, 2, 1, 1, 0, 2] * 4)                 # first occurrence order: 2, 0, 1
    cg.datasets.tr_samples, cg.datasets.tr_labels = torch.rand(40, 6, generator=g), y
    cg.datasets.feature_num, cg.datasets.label_num = 6, 3
    gc.batch_size, gc.d_loop_num, gc.c_loop_num, gc.g_loop_num = 16, 5, 4, 3
    yield cg
    (gc.epochs, gc.batch_size, gc.d_loop_num, gc.c_loop_num, gc.g_loop_num, gc.g_lr, cg.datasets.tr_samples, cg.datasets.tr_labels,
     cg.datasets.feature_num, cg.datasets.label_num) = saved


def _visits(eng):
    return [c for c in eng.calls if c[0] == "visit"]


@pytest.mark.parametrize("cls_name,loops,flag_name,history,g_fwd,uses_encoder",
                         [("CVAEGAN", (5, 4, 3), None, ("recon_loss", "kl_loss", "adv_loss", "class_loss"), 2, True),
                          ("CGAN", (5, 4, 3), "STEP_PRIOR_ONLY", ("adv_loss", "class_loss"), 1, False),
                          ("CVAE", (0, 4, 3), "STEP_CVAE", ("recon_loss", "kl_loss", "class_loss"), 1, True)])
def test_label_visit_trainers_fit_loop(host, cls_name, loops, flag_name, history, g_fwd, uses_encoder):
    cg = host
    from cvae_gan_b200 import _lib
    gc = cg.config.gan_config
    gc.epochs = 3
    torch.manual_seed(4)
    gan = getattr(cg, cls_name)()
    eng = gan.engine
    gan.use_cuda_graphs = False
    gan.fit(cg.datasets.TrDataset())
    # partition: key order = first occurrence, rows keep their order (cvae_gan.py:238-245)
    assert list(gan.samples.keys()) == [2, 0, 1]
    for lab in (0, 1, 2):
        assert torch.equal(gan.samples[lab], cg.datasets.tr_samples[cg.datasets.tr_labels == lab])
    # fresh optimisers per fit (cvae_gan.py:75-97)
    assert eng.adam_steps == [0, 0, 0, 0] and all(float(m.abs().sum()) == 0.0 for m in eng.adam_m + eng.adam_v + eng.grads)
    # one visit per (epoch, label) in partition order; loops per trainer; e < 200 -> lambda_class = 0 (cvae_gan.py:198-204)
    flag = getattr(_lib, flag_name) if flag_name else 0
    vs = _visits(eng)
    assert [v[1] for v in vs] == [2, 0, 1] * 3
    assert all(v[2] == 16 and v[4] == loops and v[5] == (_lib.VISIT_LAMBDA_ZERO | flag) for v in vs)
    assert [v[3] for v in vs[:3]] == [tuple(gan.samples[lab].shape) for lab in (2, 0, 1)]
    lam_calls = [c[3] for c in eng.calls if c[0] == "ctl_set" and c[3] is not None]
    assert lam_calls == [0.0, 0.0, 0.0]
    # the epoch's record is the LAST label's last generator step (cvae_gan.py:219-222): visits 2, 5, 8
    assert tuple(gan.loss_history) == history
    col = {"recon_loss": 0.1, "kl_loss": 0.2, "adv_loss": 0.3, "class_loss": 0.4}
    for key in history:
        assert gan.loss_history[key] == pytest.approx([k + col[key] for k in (2, 5, 8)])
    # Philox counter: two values per optimiser step; BatchNorm forward counts
    n_steps = sum(loops)
    assert gan._counter == 2 * n_steps * 9
    nbt = int(gan.generator.state_dict()["main_model.1.num_batches_tracked"])
    assert nbt == 9 * (loops[0] + loops[1] + g_fwd * loops[2])
    enc = gan.encoder if uses_encoder else gan._encoder_module
    assert int(enc.state_dict()["encoder.1.num_batches_tracked"]) == (9 * loops[2] if uses_encoder else 0)
    assert all(not m.training for m in gan._networks())
    assert hasattr(gan, "encoder") == uses_encoder and hasattr(gan, "discriminator") == (cls_name != "CVAE")


def test_lambda_class_ramp_reaches_the_engine(host):
    """Epochs >= 200 switch the visits to the lambda_class != 0 graph and hand the scheduled weight to the control block."""
    cg = host
    from cvae_gan_b200 import _lib
    from cvae_gan_b200 import cvae_gan as mod
    gc = cg.config.gan_config
    gc.epochs = 2
    gan = cg.CVAE()
    gan.use_cuda_graphs = False
    real = mod.lambda_class_at
    try:
        mod.lambda_class_at = lambda e, lam: real(e + 349, lam)       # epochs 349, 350 of the schedule
        gan.fit(cg.datasets.TrDataset())
    finally:
        mod.lambda_class_at = real
    lam = [c[3] for c in gan.engine.calls if c[0] == "ctl_set" and c[3] is not None]
    assert lam == pytest.approx([0.1 * 149 / 300, 0.1 * 150 / 300])      # cvae_config lambda_class = 0.1
    assert all(v[5] == _lib.STEP_CVAE for v in _visits(gan.engine))


def test_stale_hyper_parameters_are_refused(host):
    cg = host
    gc = cg.config.gan_config
    gc.epochs = 1
    gan = cg.CVAEGAN()
    gan.use_cuda_graphs = False
    gc.g_lr = 5e-4                                        # the reference would build its optimisers with this value in fit()
    with pytest.raises(ValueError, match="changed after"):
        gan.fit(cg.datasets.TrDataset())
    gan2 = cg.CVAEGAN()
    gan2.lambda_kl = 0.7                                  # the reference reads self.lambda_kl at every step
    with pytest.raises(ValueError, match="changed after"):
        gan2.fit(cg.datasets.TrDataset())


def test_vaegan_fit_loop(host):
    """vae_gan.py:42-157: no partition, one (d_loop, 0, g_loop) visit per EPOCH over all rows with the class weight at 0."""
    cg = host
    from cvae_gan_b200 import _lib
    gc = cg.config.gan_config
    gc.epochs = 4
    gan = cg.VAEGAN()
    eng = gan.engine
    assert eng.unconditional and eng.tables[0]["encoder.0.weight"][1] == (256, 6) and eng.tables[1]["main_model.0.weight"][1] == (256, 128)
    assert eng.cfg.lambda_kl == 0.01 and eng.cfg.lambda_adv == 0.1            # vae_gan_config
    gan.use_cuda_graphs = False
    gan.fit(cg.datasets.TrDataset())
    assert torch.equal(gan.samples, cg.datasets.tr_samples)
    vs = _visits(eng)
    assert len(vs) == 4 and all(v[1] == 0 and v[2] == 16 and v[3] == (40, 6) and v[4] == (5, 0, 3) and v[5] == _lib.VISIT_LAMBDA_ZERO
                                for v in vs)
    assert tuple(gan.loss_history) == ("recon_loss", "kl_loss", "adv_loss")
    assert gan.loss_history["adv_loss"] == pytest.approx([k + 0.3 for k in range(4)])
    assert gan._counter == 2 * 8 * 4
    assert int(gan.generator.state_dict()["main_model.1.num_batches_tracked"]) == 4 * (5 + 2 * 3)
    assert int(gan.encoder.state_dict()["encoder.1.num_batches_tracked"]) == 4 * 3
    assert eng.adam_steps == [0, 0, 0, 0] and all(not m.training for m in gan._networks())
    assert not hasattr(gan, "classifier")


# ---------------------------------------------------------------------------------------------------------------------
# generate_qualified_samples (cvae_gan.py:347-378): the chunk-of-10 / patience-20 loop replayed on fused batches
# ---------------------------------------------------------------------------------------------------------------------
def _stream_keep(rows: torch.Tensor, period: int, hits: int) -> torch.Tensor:
    return ((rows * 2654435761) % 4294967296 // 65536) % period < hits


class FilterEngine(FakeEngine):
    """generate_filter over a deterministic row stream: row r is the vector [r, r, ...] and is accepted iff _stream_keep(r);
    accepted rows come back compacted in REVERSE order (the kernel's order is whatever the atomics give)."""
    period, hits = 7, 2

    def generate_filter(self, label, n, thr, z=None, seed=0, row_offset=0, capacity=None, want_logits=False, want_keep=False):
        rows = torch.arange(row_offset, row_offset + n)
        keep = _stream_keep(rows, self.period, self.hits)
        acc = rows[keep].flip(0)
        x_out = torch.full((n, self.F), -1.0)
        idx_out = torch.full((n,), -1, dtype=torch.int64)
        x_out[:len(acc)] = acc[:, None].float().expand(-1, self.F)
        idx_out[:len(acc)] = acc
        self.calls.append(("generate_filter", label, n, thr, row_offset))
        return x_out, idx_out, torch.tensor([len(acc)]), None, keep.to(torch.uint8)


    # train-mode path (generator not yet fitted): chunk-by-chunk primitives
    def generate(self, label, n, z=None, seed=0, row_offset=0, train_mode=False):
        self.calls.append(("generate", label, n, row_offset, train_mode))
        return torch.arange(row_offset, row_offset + n)[:, None].float().expand(-1, self.F).contiguous()

    def classifier_forward(self, x):
        return x

    def filter_logits(self, logits, label, thr):
        return _stream_keep(logits[:, 0].long(), self.period, self.hits)


@pytest.mark.parametrize("period,hits,num", [(7, 2, 25), (3, 2, 200), (1000, 1, 30), (50, 0, 5), (2, 1, 1), (400, 1, 3000)])
def test_generate_qualified_samples_replays_the_reference_loop(host, monkeypatch, period, hits, num):
    cg = host
    from cvae_gan_b200 import cvae_gan
    monkeypatch.setattr(cvae_gan, "Engine", FilterEngine)
    monkeypatch.setattr(FilterEngine, "period", period)
    monkeypatch.setattr(FilterEngine, "hits", hits)
    gan = cg.CVAEGAN()
    gan.generator.eval()
    gan._gen_rows = 12345                                  # the generation stream continues where earlier calls stopped
    out = gan.generate_qualified_samples(1, num, 0.5)
    # the reference's loop, literally, on the same stream
    result, pos, patience = [], 12345, 20
    while len(result) < num and patience > 0:
        n = min(10, num - len(result))
        rows = torch.arange(pos, pos + n)
        valid = rows[_stream_keep(rows, period, hits)]
        pos += n
        result.extend(valid.tolist())
        if len(valid) == 0:
            patience -= 1
    if result:
        assert out.shape == (len(result), 6) and out[:, 0].tolist() == [float(r) for r in result]      # same rows, same order
    else:
        assert out.numel() == 0 and out.shape == torch.tensor([]).shape
    assert gan._gen_rows == pos                                # the stream position the loop would have reached
    assert gan.classifier.training                             # reference quirk: C is left in train mode (cvae_gan.py:363)
    assert all(c[3] == 0.5 and c[1] == 1 for c in gan.engine.calls if c[0] == "generate_filter")


# ---------------------------------------------------------------------------------------------------------------------
# Classifier.fit (classifier.py:24-45): the batches are the ones DataLoader(dataset, batch_size, shuffle=True) would yield
# ---------------------------------------------------------------------------------------------------------------------
def test_classifier_fit_batches_follow_dataloader_shuffle(host):
    from torch.utils.data import DataLoader, TensorDataset
    cg = host
    cc = cg.config.classifier_config
    saved = (cc.epochs, cc.batch_size)
    try:
        cc.epochs, cc.batch_size = 3, 16
        eng = FakeEngine(6, 3)
        seen = []
        eng.max_batch = 64
        eng.reset_adam = lambda net: eng.calls.append(("reset_adam", net))
        eng.step_classifier = lambda x, y, lr, seed, counter, loss_out: seen.append((x.clone(), y.clone(), lr, counter))
        clf = cg.Classifier("T")
        clf.model.attach(eng, 3)
        torch.manual_seed(11)
        clf.fit(cg.datasets.TrDataset())
        assert ("reset_adam", 3) in eng.calls and len(clf.loss_history) == 3 and not clf.model.training
        # the reference's loader under the same generator state (one draw precedes the epochs here: the dropout stream's seed)
        torch.manual_seed(11)
        torch.empty((), dtype=torch.int64).random_()
        loader = DataLoader(TensorDataset(cg.datasets.tr_samples, cg.datasets.tr_labels), batch_size=16, shuffle=True)
        want = [(xb, yb) for _ in range(3) for xb, yb in loader]
        assert len(seen) == len(want) == 9 and [int(c) for *_, c in seen] == list(range(9))
        for (x, y, lr, _), (xb, yb) in zip(seen, want):
            assert torch.equal(x, xb) and torch.equal(y, yb) and lr == cc.lr         # incl. the last partial batch of 8 rows
        with pytest.raises(ValueError, match="labels must lie"):
            cg.datasets.tr_labels = cg.datasets.tr_labels + 5
            clf.fit(cg.datasets.TrDataset())
    finally:
        cc.epochs, cc.batch_size = saved


# ---------------------------------------------------------------------------------------------------------------------
# plotting surface (cvae_gan.py:263-337, cgan.py:216-262, cvae.py:203-261, vae_gan.py:180-236, classifier.py:202-303):
# matplotlib is not in this image; a stand-in module checks that the methods run and save under the reference's file names
# ---------------------------------------------------------------------------------------------------------------------
def test_plot_methods_run_and_use_the_reference_file_names(host, monkeypatch, tmp_path):
    import sys
    from unittest import mock
    cg = host
    plt = mock.MagicMock()
    mpl = mock.MagicMock()
    mpl.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    monkeypatch.setattr(cg.config, "path_config", SimpleNamespace(gan_outs=tmp_path), raising=False)
    want = {"CVAEGAN": ["cvae_gan_loss_history.jpg", "cvae_gan_combined_loss.jpg"], "CGAN": ["cgan_loss_history.jpg", "cgan_combined_loss.jpg"],
            "CVAE": ["cvae_loss_history.jpg", "cvae_combined_loss.jpg"], "VAEGAN": ["vae_gan_loss_history.jpg", "vae_gan_combined_loss.jpg"]}
    for name, files in want.items():
        plt.reset_mock()
        gan = getattr(cg, name)()
        for k in gan.loss_history:
            gan.loss_history[k] = [0.3, -0.2, 0.1]
        gan.plot_loss_history()
        saved = [str(c.args[0]) for c in plt.savefig.call_args_list]
        assert [s.split("/")[-1] for s in saved] == files, (name, saved)
    # ROC curves: scores come from the engine's classifier forward (the reference scores with the raw network outputs)
    eng = FakeEngine(6, 3)
    g = torch.Generator().manual_seed(3)
    scores = torch.randn(40, 3, generator=g)
    scores[torch.arange(40), cg.datasets.tr_labels] += 2.0
    eng.classifier_forward = lambda x: scores
    clf = cg.Classifier("CVAE_GAN")
    clf.model.attach(eng, 3)
    cg.datasets.te_samples, cg.datasets.te_labels = cg.datasets.tr_samples, cg.datasets.tr_labels
    plt.reset_mock()
    curves = clf.plot_roc_curve(cg.datasets.TeDataset())
    assert sorted(curves) == [0, 1, 2] and all(0.8 < c[2] <= 1.0 for c in curves.values())
    assert str(plt.savefig.call_args_list[-1].args[0]).endswith("CVAE_GAN_roc_curve_multiclass.jpg")
    binary = clf.plot_roc_curve(cg.datasets.TeDataset(), is_binary=True)
    assert list(binary) == ["binary"] and str(plt.savefig.call_args_list[-1].args[0]).endswith("CVAE_GAN_roc_curve_binary.jpg")


@pytest.mark.parametrize("period,hits,num", [(7, 2, 25), (3, 2, 120), (50, 0, 5), (1000, 1, 40), (2, 1, 31)])
def test_generate_qualified_samples_in_train_mode_runs_the_literal_chunk_loop(host, monkeypatch, period, hits, num):
    """Before fit() the generator is in train mode: every chunk of <= 10 rows is normalised with its own batch statistics, so
    the loop of cvae_gan.py:355-376 is run chunk by chunk (train-mode generator forward, classifier, filter); a trailing chunk
    of ONE row fails like torch's BatchNorm."""
    cg = host
    from cvae_gan_b200 import cvae_gan
    monkeypatch.setattr(cvae_gan, "Engine", FilterEngine)
    monkeypatch.setattr(FilterEngine, "period", period)
    monkeypatch.setattr(FilterEngine, "hits", hits)
    gan = cg.CVAEGAN()
    assert gan.generator.training
    gan._gen_rows = 777
    result, pos, patience, chunks, fails = [], 777, 20, [], False
    while len(result) < num and patience > 0:
        n = min(10, num - len(result))
        if n < 2:
            fails = True
            break
        rows = torch.arange(pos, pos + n)
        valid = rows[_stream_keep(rows, period, hits)]
        chunks.append((n, pos))
        pos += n
        result.extend(valid.tolist())
        if len(valid) == 0:
            patience -= 1
    if fails:
        with pytest.raises(ValueError, match="more than 1 value per channel"):
            gan.generate_qualified_samples(2, num, 0.5)
        return
    out = gan.generate_qualified_samples(2, num, 0.5)
    gen_calls = [c for c in gan.engine.calls if c[0] == "generate"]
    assert [(c[2], c[3]) for c in gen_calls] == chunks and all(c[1] == 2 and c[4] is True for c in gen_calls)
    if result:
        assert out.shape == (len(result), 6) and out[:, 0].tolist() == [float(r) for r in result]
    else:
        assert out.numel() == 0
    assert gan._gen_rows == pos and gan.classifier.training
    # one train-mode generator forward per chunk reaches num_batches_tracked (BatchNorm counts forwards, also under no_grad)
    assert int(gan.generator.state_dict()["main_model.1.num_batches_tracked"]) == len(chunks)
