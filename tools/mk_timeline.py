"""Per-op timeline of one label visit executed by the step-program kernel (CVG_MK_DBG=1): op kind, work items, cycles CTA 0
spent in the op, and the wall-clock distance to the next op's start (= the phase's critical path + barrier)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["CVG_MK_DBG"] = "1"
from cvae_gan_b200.engine import Engine  # noqa: E402
from cvae_gan_b200 import models  # noqa: E402

KINDS = {1: "MN", 2: "DW", 3: "DWRED", 4: "FILL", 5: "STAGE", 6: "SN_POWER", 7: "SN_DOT", 8: "SN_GRAD", 9: "LN_FWD", 10: "LN_BWD",
         11: "CE", 12: "SEED", 13: "ZERO", 14: "PACK", 15: "UNPACK", 16: "ADAM", 17: "CTL_SET", 18: "FINISH", 19: "NVL32", 20: "NVL64",
         21: "REPARAM", 22: "PREP"}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    loops = tuple(int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else "1,1,1".split(",")))
    F_, K, Z = 10, 5, 128
    torch.manual_seed(0)
    eng = Engine(F_, K, Z, max_batch=B)
    mods = [models.CVAEGANEncoderModel(F_, K, Z), models.CVAEGANGeneratorModel(Z, K, F_), models.CVAEGANDiscriminatorModel(F_, K),
            models.CVAEGANClassifierModel(F_, K)]
    for net, m in enumerate(mods):
        eng.load_state(net, m.state_dict())
    eng.debug_set("train_mode", 1)
    rows = torch.rand(200000, F_, device="cuda")
    eng.ctl_set(seed=1, counter=0, lambda_class=0.25)
    for _ in range(3):
        eng.visit(1, B, class_rows=rows, loops=loops)
    torch.cuda.synchronize()
    cyc = eng.mk_cycles()
    st = eng.mk_starts
    ops = eng.mk_ops
    tot = st[-1] - st[0] + cyc[-1]
    print(f"visit loops {loops} batch {B}: {len(cyc)} ops, {tot} cycles on CTA 0")
    agg = {}
    phase_t0, phase_ops = st[0], []
    for i, (c, (k, bar, items)) in enumerate(zip(cyc, ops)):
        wall = (st[i + 1] - st[i]) if i + 1 < len(st) else c
        name = KINDS.get(k, str(k))
        print(f"{i:4d} {name:9s} bar={bar} items={items:5d} cta0={c:7d} to_next={wall:7d}")
        a = agg.setdefault(name, [0, 0, 0])
        a[0] += 1
        a[1] += c
        a[2] += wall
    print("by kind: count, cta0 cycles, wall cycles")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        print(f"  {name:9s} {a[0]:4d} {a[1]:9d} {a[2]:9d}  {100.0 * a[2] / tot:5.1f}%")


if __name__ == "__main__":
    main()
