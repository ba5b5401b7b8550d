"""A/B check of the two training executors on a B200: the step-program kernel (tcgen05, mega.cuh) against the stand-alone
FFMA layer kernels, same parameters, batch, injected noise and masks.  Prints, per step kind, the loss rows, every
workspace matrix in dataflow order (first divergence = the op to look at), every gradient tensor and the mutated state.

    python tools/mk_ab.py [B] [F] [K]        (exit code 1 when anything differs by more than 2e-4 of its scale)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import parity as P  # noqa: E402
from cvae_gan_b200._lib import STEP_NO_UPDATE  # noqa: E402

BUFS = {
    "d": ["xT", "z", "g_h0", "g_h1", "g_h2", "g_out", "d_a0", "d_a1", "d_a2", "d_s", "d_g2", "d_g1", "d_g0"],
    "c": ["xT", "z", "g_h0", "g_h1", "g_h2", "g_out", "c_a1", "c_h2", "c_a2", "c_a3", "c_logit", "c_dlogit", "c_g2", "c_g1", "c_g0"],
    "g": ["xT", "e_h0", "e_h1", "e_h2", "e_ml", "g_h0", "g_h1", "g_h2", "g_out", "d_a0", "d_a1", "d_a2", "d_s", "c_a1", "c_h2",
          "c_a2", "c_a3", "c_logit", "c_dlogit", "d_g2", "d_g1", "d_g0", "c_g2", "c_g1", "c_g0", "dx", "g_dout", "g_dy2", "g_dy1",
          "g_dy0", "e_dml", "e_dy2", "e_dy1", "e_dy0"],
}
TWO_PASS = {"g_h0", "g_h1", "g_h2", "g_out", "d_a0", "d_a1", "d_a2", "d_s", "d_g0", "d_g1", "d_g2", "c_a1", "c_h2", "c_a2", "c_a3",
            "c_logit", "c_dlogit", "c_g0", "c_g1", "c_g2", "g_dout", "g_dy0", "g_dy1", "g_dy2"}
NPASS = {"d": {"g": 1, "d": 2, "c": 0}, "c": {"g": 1, "d": 0, "c": 2}, "g": {"g": 2, "d": 1, "c": 1}}


def clone_engine(src, B):
    from cvae_gan_b200.engine import Engine
    e = Engine(src.F, src.K, src.Z, max_batch=max(B, 64))
    for net in range(4):
        e.load_state(net, src.export_state(net))
    return e


def rel(a, b):
    sc = float(b.abs().max()) if b.numel() else 0.0
    return float((a - b).abs().max()) / (sc + 1e-30), sc


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    F_ = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    update = os.environ.get("AB_UPDATE") == "1"
    tol = 2e-4
    bad = 0
    for kind in os.environ.get("AB_KINDS", "dcg"):
        orc, ea, g = P.make_pair(F_, K, B, seed=7)
        eb = clone_engine(ea, B)
        ma, mb = (int(v) for v in os.environ.get("AB_MODES", "0,1").split(","))
        ea.debug_set("train_mode", ma)
        eb.debug_set("train_mode", mb)
        x, y = P.make_data(F_, K, [B] * K, seed=1)
        xb = x[y == 1][:B].contiguous().cuda()
        inj, dev = P.draw_noise(kind, B, 128, g)
        flags = 0 if update else STEP_NO_UPDATE
        outs = []
        for e in (ea, eb):
            e.zero_grads()
            lo = torch.zeros(4, device="cuda")
            if kind == "d":
                e.step_d(xb, 1, noise=dev, flags=flags, loss_out=lo)
            elif kind == "c":
                e.step_c(xb, 1, noise=dev, flags=flags, loss_out=lo)
            else:
                e.step_g(xb, 1, 0.25, noise=dev, flags=flags, loss_out=lo)
            torch.cuda.synchronize()
            outs.append(lo.tolist())
        print(f"== step_{kind} B={B}: program ops {eb.debug_get('mk_last_nops')}, launches ffma {ea.launch_count()} / program {eb.launch_count()}")
        print("   loss ffma   ", outs[0])
        print("   loss program", outs[1])
        for a_, b_ in zip(outs[0], outs[1]):
            if abs(a_ - b_) > 1e-4 * max(abs(a_), 1e-2):
                bad += 1
        for name in BUFS[kind]:
            if kind == "g" and name == "z":
                continue
            owner = name[0]
            np_ = NPASS[kind].get(owner, 1) if name in TWO_PASS else 1
            for ps in range(max(np_, 1)):
                a = ea.debug_read(name, B, ps)
                b = eb.debug_read(name, B, ps)
                r, sc = rel(b, a)
                flag = "" if r <= tol else "   <-- DIFF"
                if r > tol:
                    bad += 1
                print(f"   {name:9s} pass {ps}: rel {r:9.3e} scale {sc:9.3e}{flag}")
        for net in range(4):
            for key in ea.tables[net]:
                kd = ea.tables[net][key][0]
                a = ea.view(net, key, "grads") if kd == 0 else ea.view(net, key)
                b = eb.view(net, key, "grads") if kd == 0 else eb.view(net, key)
                r, sc = rel(b, a)
                if sc == 0.0 and float(b.abs().max()) == 0.0:
                    continue
                lim = tol if sc > 1e-7 else 1.0     # pre-BN biases are pure round-off on both sides
                flag = "" if r <= lim else "   <-- DIFF"
                if r > lim:
                    bad += 1
                what = "grad " if kd == 0 else "state"
                print(f"   {what} {P.NETS[net]}/{key}: rel {r:9.3e} scale {sc:9.3e}{flag}")
            if update:
                r, sc = rel(eb.params[net], ea.params[net])
                print(f"   params {P.NETS[net]}: rel {r:9.3e}")
        if kind == "g" and os.environ.get("AB_CHECK_EDY1"):
            # float64 recomputation of the encoder's layer-2 input gradient from engine A's own buffers: who is right?
            for tag, e in (("ffma", ea), ("program", eb)):
                dy2 = e.debug_read("e_dy2", B).double()
                h2 = e.debug_read("e_h2", B).double()
                h1 = e.debug_read("e_h1", B).double()
                W2 = e.view(0, "encoder.6.weight").double()
                g2 = e.view(0, "encoder.7.weight").double()
                g1, b1 = e.view(0, "encoder.4.weight").double(), e.view(0, "encoder.4.bias").double()
                m2, v2 = h2.mean(0), h2.var(0, unbiased=False)
                r2 = 1.0 / torch.sqrt(v2 + 1e-5)
                xh2 = (h2 - m2) * r2
                v = g2 * r2 * (dy2 - dy2.mean(0) - xh2 * (dy2 * xh2).mean(0))
                y = v @ W2
                m1, v1 = h1.mean(0), h1.var(0, unbiased=False)
                pre1 = (h1 - m1) / torch.sqrt(v1 + 1e-5) * g1 + b1
                want = torch.where(pre1 > 0, y, 0.2 * y)
                got = e.debug_read("e_dy1", B).double()
                r, sc = rel(got, want)
                print(f"   e_dy1 of {tag} vs float64 recomputation from its own inputs: rel {r:9.3e} scale {sc:9.3e}; "
                      f"|v| {float(v.abs().max()):.3e} |dy2| {float(dy2.abs().max()):.3e} min var2 {float(v2.min()):.3e}")
        cyc = eb.mk_cycles() if os.environ.get("CVG_MK_DBG") else []
        if cyc:
            print("   op cycles (CTA 0):", cyc)
            print("   mn sections (CTA 0) [preamble, wait, store, arrive, -, -, epilogue+bar, -, chunks, items | epi: fetch0, mma wait, tmem+transpose, groups, stats]:", eb.mk_sections)
        ea.close()
        eb.close()
    print("A/B", "FAIL" if bad else "OK", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
