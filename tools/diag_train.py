import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvae_gan_b200.engine import Engine
from cvae_gan_b200 import models
F_, K, Z, B = 10, 4, 128, 4096
dev = torch.device('cuda')
eng = Engine(F_, K, Z, max_batch=B)
torch.manual_seed(1)
mods = [models.CVAEGANEncoderModel(F_, K, Z), models.CVAEGANGeneratorModel(Z, K, F_), models.CVAEGANDiscriminatorModel(F_, K), models.CVAEGANClassifierModel(F_, K)]
for net, m in enumerate(mods): eng.load_state(net, m.state_dict())
g = torch.Generator().manual_seed(7)
tabs = []
for k, fr in enumerate((0.90, 0.06, 0.03, 0.01)):
    c = torch.rand(F_, generator=g); n = int(1_000_000 * fr)
    tabs.append((c.to(dev) + 0.08 * torch.randn(n, F_, device=dev)).clamp_(0, 1).contiguous())
eng.ctl_set(seed=4321, counter=0, lambda_class=float(sys.argv[1]) if len(sys.argv) > 1 else 0.25)
loss = torch.zeros(13, 4, device=dev)
for e in range(401):
    for label in range(K):
        eng.visit(label, B, class_rows=tabs[label], loss_out=loss)
    if e % 50 == 0:
        torch.cuda.synchronize()
        accs, conf, gacc = [], [], []
        for label in range(K):
            lg = eng.classifier_forward(tabs[label][:5000])
            p = torch.softmax(lg, 1)
            accs.append(float((p.argmax(1) == label).float().mean()))
            conf.append(float(p.max(1).values.mean()))
            xg = eng.generate(label, 5000, seed=9)
            pg = torch.softmax(eng.classifier_forward(xg), 1)
            gacc.append((round(float((pg.argmax(1) == label).float().mean()), 3), round(float(pg.max(1).values.mean()), 3)))
        l = loss.tolist()
        print(f"epoch {e}: D {l[4][0]:.4f} C {l[9][0]:.4f} G {[round(v,4) for v in l[12]]} real acc {[round(a,3) for a in accs]} conf {[round(a,3) for a in conf]} gen (acc,conf) {gacc}", flush=True)
