"""Development aid: runs the standalone filter kernel (and optionally the fused generation kernel) a few times so that
ncu can capture them.  Not a benchmark."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cvae_gan_b200.engine import Engine
from cvae_gan_b200 import models

F_, K_, Z_ = 10, 5, 128
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
what = sys.argv[2] if len(sys.argv) > 2 else "filter"
dev = torch.device("cuda", 0)
eng = Engine(F_, K_, Z_, max_batch=4096)
torch.manual_seed(0)
mods = [models.CVAEGANEncoderModel(F_, K_, Z_), models.CVAEGANGeneratorModel(Z_, K_, F_),
        models.CVAEGANDiscriminatorModel(F_, K_), models.CVAEGANClassifierModel(F_, K_)]
for net, m in enumerate(mods):
    eng.load_state(net, m.state_dict())
if what == "filter":
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.rand(n, F_, device=dev, generator=g)
    lg = 3.0 * torch.randn(n, K_, device=dev, generator=g)
    thr0 = float(os.environ.get("PF_THR", "0.5"))
    for _ in range(4):
        out = eng.filter_compact(x, lg, 0, thr0)
    torch.cuda.synchronize()
    print("accepted", int(out[2].item()))
else:
    for _ in range(3):
        out = eng.generate_filter(0, n, 0.5, seed=1)
    torch.cuda.synchronize()
    print("accepted", int(out[2].item()))
if what == "filter" and len(sys.argv) > 3:
    for thr in (2.0, 0.9, 0.5, 0.0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            out = eng.filter_compact(x, lg, 0, thr)
        torch.cuda.synchronize()
        from cvae_gan_b200._lib import check
        from cvae_gan_b200.engine import _ptr, _stream
        x_out = torch.empty(n, F_, device=dev); idx_out = torch.empty(n, dtype=torch.int64, device=dev)
        e0.record()
        for _ in range(10):
            eng.count_buf.zero_()
            check(eng.lib.cvg_filter_compact(_ptr(x), _ptr(lg), n, F_, K_, 0, float(thr), 0, _ptr(x_out), _ptr(idx_out), n, _ptr(eng.count_buf), _stream()))
        e1.record()
        torch.cuda.synchronize()
        print("thr", thr, "ms", e0.elapsed_time(e1) / 10, "accepted", int(eng.count_buf.item()))
