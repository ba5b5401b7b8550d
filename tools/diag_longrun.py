"""Long-run training behaviour, oracle (CPU, torch RNG) vs CUDA engine (Philox): loss curves of the same configuration.
Not a parity test (different random streams) - a check that both train the same way over hundreds of steps.

    python tools/diag_longrun.py oracle|gpu [epochs] [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
F_, K, Z = 10, 4, 128


def data(rows=(4000, 600, 300, 300)):
    g = torch.Generator().manual_seed(7)
    xs, ys = [], []
    for k, n in enumerate(rows):
        c = torch.rand(F_, generator=g)
        xs.append((c + 0.08 * torch.randn(n, F_, generator=g)).clamp(0, 1))
        ys.append(torch.full((n,), k, dtype=torch.long))
    return torch.cat(xs), torch.cat(ys)


def main():
    which = sys.argv[1]
    epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    x, y = data()
    if which == "oracle":
        from oracle import cvae_gan_oracle as O
        torch.manual_seed(0)
        torch.set_num_threads(8)
        orc = O.OracleCVAEGAN(F_, K, O.OracleConfig(batch_size=B)).init_like_reference(torch.Generator().manual_seed(1))
        orc.make_optimizers()
        orc.divide_samples(x, y)
        noise = O.TorchNoise()
        for e in range(epochs + 1):
            for label in range(K):
                for _ in range(5):
                    orc.step_d(orc.get_target_samples(label, B, noise), label, noise)
                for _ in range(5):
                    orc.step_c(orc.get_target_samples(label, B, noise), label, noise)
                for _ in range(3):
                    l, _ = orc.step_g(orc.get_target_samples(label, B, noise), label, noise, 0.25)
            if e % 10 == 0:
                print(f"oracle epoch {e}: G {[round(v, 4) for v in l.values()]}", flush=True)
    else:
        from cvae_gan_b200.engine import Engine
        from cvae_gan_b200 import models
        eng = Engine(F_, K, Z, max_batch=B)
        torch.manual_seed(1)
        mods = [models.CVAEGANEncoderModel(F_, K, Z), models.CVAEGANGeneratorModel(Z, K, F_), models.CVAEGANDiscriminatorModel(F_, K),
                models.CVAEGANClassifierModel(F_, K)]
        for net, m in enumerate(mods):
            eng.load_state(net, m.state_dict())
        tabs = [x[y == k].cuda().contiguous() for k in range(K)]
        eng.ctl_set(seed=4321, counter=0, lambda_class=0.25)
        loss = torch.zeros(13, 4, device="cuda")
        for e in range(epochs + 1):
            for label in range(K):
                eng.visit(label, B, class_rows=tabs[label], loss_out=loss)
            if e % 10 == 0:
                torch.cuda.synchronize()
                print(f"gpu epoch {e}: G {[round(v, 4) for v in loss[12].tolist()]}", flush=True)


if __name__ == "__main__":
    main()
