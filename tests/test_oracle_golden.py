"""Pins oracle/cvae_gan_oracle.py to the UNMODIFIED reference (fixtures made by
oracle/make_golden.py from /root/reference; see that file for what each fixture is)."""
import os

import numpy as np
import pytest
import torch

from oracle import cvae_gan_oracle as O

NETS = O.NETS


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _states(npz, prefix):
    out = {n: {} for n in NETS}
    for k in npz.files:
        if k.startswith(prefix + "/"):
            _, net, key = k.split("/", 2)
            out[net][key] = torch.from_numpy(npz[k])
    return out


def _replay_fit(npz, epoch_offset):
    F_, K, B, fit_seed, gen_seed, _ = [int(v) for v in npz["meta"]]
    cfg = O.OracleConfig(batch_size=B, epochs=2, epoch_offset=epoch_offset)
    orc = O.OracleCVAEGAN(F_, K, cfg).load_state(_states(npz, "init"))
    torch.manual_seed(fit_seed)
    return orc, gen_seed


@pytest.fixture(scope="module")
def fit_a(golden_dir):
    torch.set_num_threads(1)
    npz = _load(golden_dir, "ref_fit_a.npz")
    orc, gen_seed = _replay_fit(npz, 0)
    orc.fit(torch.from_numpy(npz["x"]), torch.from_numpy(npz["y"]))
    return npz, orc, gen_seed


def test_tensor_table_matches_reference_state_dict(golden_dir):
    npz = _load(golden_dir, "ref_fit_a.npz")
    st = _states(npz, "init")
    for net in NETS:
        tab = O.tensor_table(net, 10, 5, 128)
        assert [k for k, _, _ in tab] == list(st[net].keys())        # same keys, same ORDER
        for k, shape, _ in tab:
            assert tuple(st[net][k].shape) == tuple(shape), (net, k)
    n_params = {net: sum(int(np.prod(s)) for _, s, kind in O.tensor_table(net, 10, 5, 128) if kind == "param")
                for net in NETS}
    # E 62 784 / G 77 002 / D 45 313 / C 44 549
    assert n_params == {"encoder": 62784, "generator": 77002, "discriminator": 45313, "classifier": 44549}


def test_fit_epoch0_1_matches_reference(fit_a):
    npz, orc, _ = fit_a
    assert list(orc.samples.keys()) == npz["sample_keys"].tolist()
    for k in orc.loss_history:
        np.testing.assert_allclose(orc.loss_history[k], npz["loss/" + k], rtol=2e-6, atol=1e-7)
    fin = _states(npz, "final")
    st = orc.state()
    for net in NETS:
        for key, ref in fin[net].items():
            got = st[net][key]
            if ref.dtype == torch.int64:
                assert torch.equal(got, ref), (net, key)
            else:
                torch.testing.assert_close(got, ref, rtol=2e-5, atol=2e-7, msg=f"{net}/{key}")


def test_generation_after_fit_matches_reference(fit_a):
    npz, orc, gen_seed = fit_a
    torch.manual_seed(gen_seed)
    s = orc.generate_samples(1, 37)
    torch.testing.assert_close(s, torch.from_numpy(npz["gen/samples_l1_n37"]), rtol=1e-5, atol=1e-6)
    for thr in (0.2, 0.5):
        for lab in (0, 3):
            q = orc.generate_qualified_samples(lab, 25, thr)
            ref = torch.from_numpy(npz[f"gen/qualified_l{lab}_thr{thr}"])
            q = q.reshape(-1, 10) if q.numel() else torch.zeros(0, 10)
            assert q.shape == ref.shape, (thr, lab, q.shape, ref.shape)
            torch.testing.assert_close(q, ref, rtol=1e-5, atol=1e-6)
    # the reference leaves the classifier in train mode after generate_qualified_samples
    assert int(orc.training["classifier"]) == int(npz["gen/classifier_training_after"][0])


def test_fit_epoch350_lambda_class_matches_reference(golden_dir):
    torch.set_num_threads(1)
    a = _load(golden_dir, "ref_fit_a.npz")
    b = _load(golden_dir, "ref_fit_b.npz")
    orc, _ = _replay_fit(a, 350)
    orc.fit(torch.from_numpy(a["x"]), torch.from_numpy(a["y"]))
    assert O.lambda_class_schedule(350, 0.5) == pytest.approx(0.25)
    for k in orc.loss_history:
        np.testing.assert_allclose(orc.loss_history[k], b["loss/" + k], rtol=2e-6, atol=1e-7)
    st = orc.state()
    for net in NETS:
        for key, t in st[net].items():
            f = t.double().ravel().numpy()
            d = b[f"digest/{net}/{key}"]
            np.testing.assert_allclose([f.sum(), (f * f).sum()], d[:2], rtol=1e-5, atol=1e-6, err_msg=f"{net}/{key}")
            n = min(8, f.size)
            np.testing.assert_allclose(f[:n], d[2:2 + n], rtol=2e-5, atol=2e-7)
            np.testing.assert_allclose(f[-n:], d[10:10 + n][-n:] if f.size >= 8 else d[10:10 + n], rtol=2e-5, atol=2e-7)


def test_filter_decisions_match_reference(golden_dir):
    npz = _load(golden_dir, "ref_filter.npz")
    logits = torch.from_numpy(npz["logits"])
    n = logits.shape[0]
    for lab in range(5):
        for thr in (0.0, 0.2, 0.5, 0.9):
            want = np.unpackbits(npz[f"keep_l{lab}_thr{thr}"])[:n].astype(bool)
            got = O.filter_logits(logits, lab, thr).numpy()
            assert np.array_equal(got, want)
    # ties: first maximal index wins (torch.max), and a 5-way tie has p = 0.2 which is NOT > 0.2
    assert O.filter_logits(torch.zeros(3, 5), 0, 0.19).all()
    assert not O.filter_logits(torch.zeros(3, 5), 0, 0.2).any()
    assert not O.filter_logits(torch.zeros(3, 5), 1, 0.0).any()


def test_patience_scan_equals_literal_loop():
    g = torch.Generator().manual_seed(3)
    for p_acc, num in ((0.5, 57), (0.02, 40), (0.0, 10), (1.0, 33)):
        keep = torch.rand(5000, generator=g) < p_acc
        # literal loop
        result, patience, pos = 0, 20, 0
        while result < num and patience > 0:
            n = min(10, num - result)
            k = int(keep[pos:pos + n].sum())
            pos += n
            result += k
            if k == 0:
                patience -= 1
        assert O.patience_scan(keep, num) == (pos, result)


# ---------------------------------------------------------------------------------------------------
# downstream classifier (classifier.py): oracle/classifier_oracle.py against the unmodified reference
# ---------------------------------------------------------------------------------------------------
def test_classifier_fit_and_test_match_reference(golden_dir):
    import os
    import numpy as np
    from oracle import classifier_oracle as CO
    gz = np.load(os.path.join(golden_dir, "ref_clf.npz"))
    F_, K, epochs, bs, seed = [int(v) for v in gz["meta"]]
    sd = {k[len("init/"):]: torch.from_numpy(gz[k]).clone() for k in gz.files if k.startswith("init/")}
    xtr, ytr = torch.from_numpy(gz["xtr"]), torch.from_numpy(gz["ytr"])
    xte, yte = torch.from_numpy(gz["xte"]), torch.from_numpy(gz["yte"])
    torch.set_num_threads(1)
    torch.manual_seed(seed)
    CO.fit(sd, xtr, ytr, epochs, float(gz["lr"][0]), bs)
    for k in sd:
        ref = torch.from_numpy(gz["final/" + k])
        assert torch.allclose(sd[k], ref, rtol=2e-4, atol=2e-6), (k, float((sd[k] - ref).abs().max()))
    m, cm = CO.macro_metrics(yte, CO.predict(sd, xte), K)
    assert np.array_equal(cm.numpy().astype(np.int64), gz["confusion"])
    assert np.allclose([m["Precision"], m["Recall"], m["F1"]], gz["metrics"], atol=1e-12)


# ---------------------------------------------------------------------------------------------------
# sibling trainer CGAN (src/cgan.py) - fixtures made by oracle/make_golden_cgan.py from the unmodified reference
# ---------------------------------------------------------------------------------------------------
NETS3 = ("generator", "discriminator", "classifier")


def _replay_cgan(golden_dir, epoch_offset):
    torch.set_num_threads(1)
    npz = _load(golden_dir, "ref_cgan_a.npz")
    F_, K, B, fit_seed, gen_seed, _ = [int(v) for v in npz["meta"]]
    st = _states(npz, "init")
    st["encoder"] = _states(_load(golden_dir, "ref_fit_a.npz"), "init")["encoder"]      # CGAN has no encoder: never touched
    cfg = O.OracleConfig(batch_size=B, epochs=2, epoch_offset=epoch_offset)
    orc = O.OracleCVAEGAN(F_, K, cfg).load_state(st)
    enc0 = {k: v.detach().clone() for k, v in orc.sd["encoder"].items()}
    torch.manual_seed(fit_seed)
    orc.fit_cgan(torch.from_numpy(npz["x"]), torch.from_numpy(npz["y"]))
    for k, v in orc.sd["encoder"].items():
        assert torch.equal(v.detach(), enc0[k]), k
    return npz, orc, gen_seed


def test_cgan_fit_epoch0_1_and_generation_match_reference(golden_dir):
    npz, orc, gen_seed = _replay_cgan(golden_dir, 0)
    assert list(orc.samples.keys()) == npz["sample_keys"].tolist()
    assert sorted(orc.loss_history) == ["adv_loss", "class_loss"]
    for k in orc.loss_history:
        np.testing.assert_allclose(orc.loss_history[k], npz["loss/" + k], rtol=2e-6, atol=1e-7)
    fin = _states(npz, "final")
    st = orc.state()
    for net in NETS3:
        for key, ref in fin[net].items():
            got = st[net][key]
            if ref.dtype == torch.int64:
                assert torch.equal(got, ref), (net, key)
            else:
                torch.testing.assert_close(got, ref, rtol=2e-5, atol=2e-7, msg=f"{net}/{key}")
    torch.manual_seed(gen_seed)
    s = orc.generate_samples(1, 37)
    torch.testing.assert_close(s, torch.from_numpy(npz["gen/samples_l1_n37"]), rtol=1e-5, atol=1e-6)
    for thr in (0.2, 0.5):
        for lab in (0, 3):
            q = orc.generate_qualified_samples(lab, 25, thr)
            ref = torch.from_numpy(npz[f"gen/qualified_l{lab}_thr{thr}"])
            q = q.reshape(-1, 10) if q.numel() else torch.zeros(0, 10)
            assert q.shape == ref.shape, (thr, lab, q.shape, ref.shape)
            torch.testing.assert_close(q, ref, rtol=1e-5, atol=1e-6)


def test_cgan_fit_epoch350_lambda_class_matches_reference(golden_dir):
    b = _load(golden_dir, "ref_cgan_b.npz")
    _, orc, _ = _replay_cgan(golden_dir, 350)
    for k in orc.loss_history:
        np.testing.assert_allclose(orc.loss_history[k], b["loss/" + k], rtol=2e-6, atol=1e-7)
    st = orc.state()
    for net in NETS3:
        for key, t in st[net].items():
            f = t.double().ravel().numpy()
            d = b[f"digest/{net}/{key}"]
            np.testing.assert_allclose([f.sum(), (f * f).sum()], d[:2], rtol=1e-5, atol=1e-6, err_msg=f"{net}/{key}")
            n = min(8, f.size)
            np.testing.assert_allclose(f[:n], d[2:2 + n], rtol=2e-5, atol=2e-7)


# ---------------------------------------------------------------------------------------------------
# sibling trainer CVAE (src/cvae.py) - fixtures made by oracle/make_golden_cvae.py from the unmodified reference
# ---------------------------------------------------------------------------------------------------
NETS_CVAE = ("encoder", "generator", "classifier")


def _replay_cvae(golden_dir, epoch_offset):
    torch.set_num_threads(1)
    npz = _load(golden_dir, "ref_cvae_a.npz")
    F_, K, B, fit_seed, gen_seed, _ = [int(v) for v in npz["meta"]]
    st = _states(npz, "init")
    st["discriminator"] = _states(_load(golden_dir, "ref_fit_a.npz"), "init")["discriminator"]   # CVAE has no critic: never touched
    # cvae_config (gan_config.py:51-56): lambda_recon 1.0, lambda_kl 0.01, lambda_class 0.1
    cfg = O.OracleConfig(batch_size=B, epochs=2, epoch_offset=epoch_offset, lambda_recon=1.0, lambda_kl=0.01, lambda_class=0.1)
    orc = O.OracleCVAEGAN(F_, K, cfg).load_state(st)
    d0 = {k: v.detach().clone() for k, v in orc.sd["discriminator"].items()}
    torch.manual_seed(fit_seed)
    orc.fit_cvae(torch.from_numpy(npz["x"]), torch.from_numpy(npz["y"]))
    for k, v in orc.sd["discriminator"].items():
        assert torch.equal(v.detach(), d0[k]), k
    return npz, orc, gen_seed


def test_cvae_fit_epoch0_1_generation_and_reconstruction_match_reference(golden_dir):
    npz, orc, gen_seed = _replay_cvae(golden_dir, 0)
    assert list(orc.samples.keys()) == npz["sample_keys"].tolist()
    assert sorted(orc.loss_history) == ["class_loss", "kl_loss", "recon_loss"]
    for k in orc.loss_history:
        np.testing.assert_allclose(orc.loss_history[k], npz["loss/" + k], rtol=2e-6, atol=1e-7)
    fin = _states(npz, "final")
    st = orc.state()
    for net in NETS_CVAE:
        for key, ref in fin[net].items():
            got = st[net][key]
            if ref.dtype == torch.int64:
                assert torch.equal(got, ref), (net, key)
            else:
                torch.testing.assert_close(got, ref, rtol=2e-5, atol=2e-7, msg=f"{net}/{key}")
    torch.manual_seed(gen_seed)
    s = orc.generate_samples(1, 37)
    torch.testing.assert_close(s, torch.from_numpy(npz["gen/samples_l1_n37"]), rtol=1e-5, atol=1e-6)
    for thr in (0.2, 0.5):
        for lab in (0, 3):
            q = orc.generate_qualified_samples(lab, 25, thr)
            ref = torch.from_numpy(npz[f"gen/qualified_l{lab}_thr{thr}"])
            q = q.reshape(-1, 10) if q.numel() else torch.zeros(0, 10)
            assert q.shape == ref.shape, (thr, lab, q.shape, ref.shape)
            torch.testing.assert_close(q, ref, rtol=1e-5, atol=1e-6)
    rec = orc.reconstruct_samples_cvae(torch.from_numpy(npz["rec/x"]), torch.from_numpy(npz["rec/y"]))
    torch.testing.assert_close(rec, torch.from_numpy(npz["rec/out"]), rtol=1e-5, atol=1e-6)
    assert npz["rec/modes_after"].tolist() == [orc.training["encoder"], orc.training["generator"], orc.training["classifier"]]


def test_cvae_fit_epoch350_lambda_class_matches_reference(golden_dir):
    b = _load(golden_dir, "ref_cvae_b.npz")
    _, orc, _ = _replay_cvae(golden_dir, 350)
    for k in orc.loss_history:
        np.testing.assert_allclose(orc.loss_history[k], b["loss/" + k], rtol=2e-6, atol=1e-7)
    st = orc.state()
    for net in NETS_CVAE:
        for key, t in st[net].items():
            f = t.double().ravel().numpy()
            d = b[f"digest/{net}/{key}"]
            np.testing.assert_allclose([f.sum(), (f * f).sum()], d[:2], rtol=1e-5, atol=1e-6, err_msg=f"{net}/{key}")
            n = min(8, f.size)
            np.testing.assert_allclose(f[:n], d[2:2 + n], rtol=2e-5, atol=2e-7)


# ---------------------------------------------------------------------------------------------------
# the width-configurable restatement, checked at the reference's own widths
# ---------------------------------------------------------------------------------------------------
def test_configurable_widths_reproduce_the_reference_at_its_own_widths(golden_dir):
    """`OracleConfig.hidden` (the widened model of BASELINE.json configs[4] is not expressible in the reference): with the
    widths the reference's formulas give for F = 10, K = 5, Z = 128 - (256, 128, 64) in all four networks - the tensor tables
    are the reference's and a 2-epoch fit replays the unmodified reference's losses and final parameters."""
    torch.set_num_threads(1)
    for net in NETS:
        assert O.tensor_table(net, 10, 5, 128, hidden=(256, 128, 64)) == O.tensor_table(net, 10, 5, 128)
        wide = dict((k, s) for k, s, _ in O.tensor_table(net, 10, 5, 128, hidden=(1024, 512, 256)))
        assert all(len(s) < 2 or max(s) <= 1024 for s in wide.values())
    assert dict((k, s) for k, s, _ in O.tensor_table("classifier", 10, 5, 128, hidden=(1024, 512, 256)))["classifier_network.7.weight"] == (256, 512)
    npz = _load(golden_dir, "ref_fit_a.npz")
    F_, K, B, fit_seed, _, _ = [int(v) for v in npz["meta"]]
    cfg = O.OracleConfig(batch_size=B, epochs=2, hidden=(256, 128, 64))
    orc = O.OracleCVAEGAN(F_, K, cfg).load_state(_states(npz, "init"))
    torch.manual_seed(fit_seed)
    orc.fit(torch.from_numpy(npz["x"]), torch.from_numpy(npz["y"]))
    for k in orc.loss_history:
        np.testing.assert_allclose(orc.loss_history[k], npz["loss/" + k], rtol=2e-6, atol=1e-7)
    fin = _states(npz, "final")
    st = orc.state()
    for net in NETS:
        for key, ref in fin[net].items():
            if ref.dtype != torch.int64:
                torch.testing.assert_close(st[net][key], ref, rtol=2e-5, atol=2e-7, msg=f"{net}/{key}")


# ---------------------------------------------------------------------------------------------------
# sibling trainer VAE-GAN (src/vae_gan.py, unconditional networks) - fixture made by oracle/make_golden_vaegan.py
# ---------------------------------------------------------------------------------------------------
def test_vaegan_fit_generation_and_reconstruction_match_reference(golden_dir):
    torch.set_num_threads(1)
    npz = _load(golden_dir, "ref_vaegan.npz")
    F_, K, B, fit_seed, gen_seed, epochs = [int(v) for v in npz["meta"]]
    st = _states(npz, "init")
    st["classifier"] = _states(_load(golden_dir, "ref_fit_a.npz"), "init")["classifier"]     # VAE-GAN has no classifier: never touched
    # vae_gan_config (gan_config.py:33-38): lambda_recon 1.0, lambda_kl 0.01, lambda_adv 0.1
    cfg = O.OracleConfig(batch_size=B, epochs=epochs, lambda_recon=1.0, lambda_kl=0.01, lambda_adv=0.1, unconditional=True)
    orc = O.OracleCVAEGAN(F_, K, cfg).load_state(st)
    assert orc.sd["encoder"]["encoder.0.weight"].shape == (256, F_) and orc.sd["generator"]["main_model.0.weight"].shape == (256, 128)
    c0 = {k: v.detach().clone() for k, v in orc.sd["classifier"].items()}
    torch.manual_seed(fit_seed)
    orc.fit_vaegan(torch.from_numpy(npz["x"]))
    assert len(orc.samples) == int(npz["n_samples"][0])
    for k in orc.loss_history:
        np.testing.assert_allclose(orc.loss_history[k], npz["loss/" + k], rtol=2e-6, atol=1e-7)
    fin = _states(npz, "final")
    stt = orc.state()
    for net in ("encoder", "generator", "discriminator"):
        for key, ref in fin[net].items():
            got = stt[net][key]
            if ref.dtype == torch.int64:
                assert torch.equal(got, ref), (net, key)
            else:
                torch.testing.assert_close(got, ref, rtol=2e-5, atol=2e-7, msg=f"{net}/{key}")
    for k, v in orc.sd["classifier"].items():
        assert torch.equal(v.detach(), c0[k]), k
    torch.manual_seed(gen_seed)
    s = orc.generate_samples(None, 41)
    torch.testing.assert_close(s, torch.from_numpy(npz["gen/samples_n41"]), rtol=1e-5, atol=1e-6)
    rec = orc.reconstruct_samples_vaegan(torch.from_numpy(npz["rec/x"]))
    torch.testing.assert_close(rec, torch.from_numpy(npz["rec/out"]), rtol=1e-5, atol=1e-6)
    assert npz["rec/modes_after"].tolist() == [orc.training["encoder"], orc.training["generator"], orc.training["discriminator"]]
