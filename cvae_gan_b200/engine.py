"""Thin host wrapper over the C ABI: owns the device buffers (as torch tensors) and forwards calls.

PyTorch is plumbing here: device memory, streams, torch.distributed for the NCCL-id broadcast.
All arithmetic of the hot path runs in libcvaegan_b200.so; there is no eager/torch fallback.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import (NET_NAMES, STEP_LOCAL_BN, STEP_NO_UPDATE, CvgConfig, CvgNoise, CvgTensorDesc, check)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """One handle per (process, GPU).  Not thread-safe (like the reference)."""

    def __init__(self, feature_num: int, label_num: int, z_size: int = 128, max_batch: int = 4096, *,
                 lambda_recon=1.0, lambda_kl=0.1, lambda_adv=1.0, g_lr=2e-4, d_lr=2e-4, c_lr=1e-4,
                 betas=(0.5, 0.999), adam_eps=1e-8, world_size: int = 1, rank: int = 0, device=None, hidden=None,
                 unconditional: bool = False):
        """`hidden`: None = the reference's layer widths; (h1, h2, h3) = the widened model (BASELINE.json configs[4]), the
        same three hidden widths for all four networks (multiples of 64, <= 1024; h2 <= 512).
        `unconditional`: encoder, generator and critic without the one-hot label columns (the sibling trainer VAE-GAN's
        networks, vae_gan_models.py); `label` arguments are then ignored."""
        if not torch.cuda.is_available():
            raise _lib.CvgError("cvae_gan_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        torch.cuda.set_device(self.device)
        self.F, self.K, self.Z = int(feature_num), int(label_num), int(z_size)
        self.max_batch = int(max_batch)
        self.world_size, self.rank = int(world_size), int(rank)
        cfg = CvgConfig(self.F, self.K, self.Z, self.max_batch, self.world_size, self.rank,
                        lambda_recon, lambda_kl, lambda_adv, g_lr, d_lr, c_lr, betas[0], betas[1], adam_eps,
                        0.1, 1e-5, 1e-5, 1e-12, 0.2, 0.3,
                        (C.c_int32 * 3)(*([int(v) for v in hidden] if hidden else [0, 0, 0])), 1 if unconditional else 0)
        self.unconditional = bool(unconditional)
        self.hidden = tuple(int(v) for v in hidden) if hidden else None
        self.cfg = cfg
        h = C.c_void_p()
        check(self.lib.cvg_create(C.byref(cfg), C.byref(h)))
        self.h = h
        # ---- buffers (torch owns all memory) ----------------------------------------------------------
        self.params, self.grads, self.adam_m, self.adam_v, self.state = [], [], [], [], []
        self.tables: list = []
        for net in range(4):
            npar, nst = C.c_int64(), C.c_int64()
            check(self.lib.cvg_net_sizes(h, net, C.byref(npar), C.byref(nst)))
            z = lambda n: torch.zeros(max(int(n), 4), dtype=torch.float32, device=self.device)
            self.params.append(z(npar.value))
            self.grads.append(z(npar.value + _lib.GRAD_TAIL))
            self.adam_m.append(z(npar.value))
            self.adam_v.append(z(npar.value))
            self.state.append(z(nst.value))
            check(self.lib.cvg_bind_net(h, net, _ptr(self.params[net]), _ptr(self.grads[net]), _ptr(self.adam_m[net]),
                                        _ptr(self.adam_v[net]), _ptr(self.state[net])))
            cnt = C.c_int32()
            check(self.lib.cvg_tensor_table(h, net, None, 0, C.byref(cnt)))
            arr = (CvgTensorDesc * cnt.value)()
            check(self.lib.cvg_tensor_table(h, net, arr, cnt.value, C.byref(cnt)))
            tab = OrderedDict()
            for d in arr:
                shape = tuple(int(d.shape[i]) for i in range(d.ndim))
                tab[d.key.decode()] = (int(d.kind), shape, int(d.offset))
            self.tables.append(tab)
        nbytes = int(self.lib.cvg_workspace_bytes(h))
        self.workspace = torch.empty(nbytes + 512, dtype=torch.uint8, device=self.device)
        base = (self.workspace.data_ptr() + 255) & ~255
        check(self.lib.cvg_bind_workspace(h, C.c_void_p(base), nbytes, _stream()))
        self.loss_buf = torch.zeros(4, dtype=torch.float32, device=self.device)
        self.count_buf = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.nvl = False
        if self.world_size > 1:
            self._init_comm()

    # ---- lifetime ------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            torch.cuda.synchronize(self.device)
            self.lib.cvg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _init_comm(self):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise _lib.CvgError("world_size > 1 needs torch.distributed to be initialised (NCCL id broadcast)")
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_uint8 * 128)()
            check(self.lib.cvg_comm_unique_id(buf))
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.to(self.device) if dist.get_backend() == "nccl" else uid
        dist.broadcast(uid, src=0)
        raw = bytes(uid.cpu().tolist())
        check(self.lib.cvg_comm_init(self.h, raw, self.rank, self.world_size))
        # latency-bound exchanges over NVLink peer memory (CUDA IPC) instead of NCCL; CVG_DISABLE_NVL=1 keeps NCCL
        import os
        if dist.get_backend() == "nccl" and self.world_size <= 8 and os.environ.get("CVG_DISABLE_NVL") != "1":
            # every rank must take the same decision: all-reduce the success flags of both stages (MIN)
            hbuf = (C.c_uint8 * 64)()
            ok = self.lib.cvg_nvl_local_handle(self.h, hbuf) == 0
            flag = torch.tensor([1 if ok else 0], device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                mine = torch.tensor(list(hbuf), dtype=torch.uint8, device=self.device)
                allh = [torch.empty_like(mine) for _ in range(self.world_size)]
                dist.all_gather(allh, mine)
                raw_h = b"".join(bytes(t.cpu().tolist()) for t in allh)
                ok = self.lib.cvg_nvl_attach(self.h, raw_h) == 0
                flag = torch.tensor([1 if ok else 0], device=self.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if int(flag.item()) != 1:
                    # no peer access between some pair of GPUs: every rank stays on NCCL
                    check(self.lib.cvg_nvl_disable(self.h))
            self.nvl = bool(int(flag.item()) == 1)
            dist.barrier()

    def verify_replicas(self):
        """Data parallel: every rank must start from the same parameters / BatchNorm / spectral-norm state (only gradients
        are all-reduced afterwards).  Compares a checksum across ranks and raises on a mismatch."""
        import torch.distributed as dist
        if self.world_size <= 1:
            return
        cs = torch.stack([t.double().sum() for t in (self.params + self.state)] +
                         [t.double().abs().sum() for t in (self.params + self.state)])
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if not torch.equal(lo, hi):
            raise _lib.CvgError("data-parallel replicas differ before training (parameters / BatchNorm / spectral-norm state): seed every "
                                "rank identically or broadcast rank 0's state (Engine.load_state) before fit()")

    # ---- named views ---------------------------------------------------------------------------------
    def view(self, net: int, key: str, which: str = "params") -> torch.Tensor:
        """View of one tensor (reference state_dict key) inside a flat buffer; `which` in
        {params, grads, adam_m, adam_v} for parameters, or the float state buffer for BN/SN state."""
        kind, shape, off = self.tables[net][key]
        n = 1
        for s in shape:
            n *= s
        buf = self.state[net] if kind == 1 else getattr(self, which)[net]
        return buf[off:off + n].view(shape)

    def load_state(self, net: int, sd: Dict[str, torch.Tensor]):
        for key, (kind, shape, off) in self.tables[net].items():
            if key not in sd:
                raise KeyError(f"{NET_NAMES[net]}: missing key {key}")
            self.view(net, key).copy_(sd[key].to(self.device, torch.float32).reshape(shape))

    def export_state(self, net: int) -> "OrderedDict[str, torch.Tensor]":
        return OrderedDict((k, self.view(net, k).detach().clone()) for k in self.tables[net])

    def zero_grads(self):
        for g in self.grads:
            g.zero_()

    def set_adam_step(self, net: int, t: int):
        check(self.lib.cvg_set_adam_step(self.h, net, int(t)))

    def get_adam_step(self, net: int) -> int:
        return int(self.lib.cvg_get_adam_step(self.h, net))

    # ---- steps ---------------------------------------------------------------------------------------
    @staticmethod
    def _noise(noise: Optional[dict], keep: list):
        """Build a CvgNoise from a dict of CUDA tensors; `keep` pins them until the call returns."""
        if not noise:
            return None
        n = CvgNoise()
        for k in ("z", "eps", "d_mask1", "d_mask2", "c_mask1", "c_mask2"):
            t = noise.get(k)
            if t is not None:
                want = torch.float32 if k in ("z", "eps") else torch.uint8
                if t.dtype != want or not t.is_cuda or not t.is_contiguous():
                    raise ValueError(f"noise['{k}'] must be a contiguous CUDA {want} tensor")
                keep.append(t)
                setattr(n, k, t.data_ptr())
        return C.byref(n)

    def _x(self, x_real: torch.Tensor) -> torch.Tensor:
        if x_real.dim() != 2 or x_real.size(1) != self.F:
            raise ValueError(f"x_real must be [B, {self.F}], got {tuple(x_real.shape)}")
        return x_real.to(self.device, torch.float32).contiguous()

    def step_d(self, x_real, label: int, noise=None, seed=0, counter=0, flags=0, loss_out=None):
        x = self._x(x_real)
        keep = [x]
        out = self.loss_buf if loss_out is None else loss_out
        check(self.lib.cvg_step_d(self.h, _ptr(x), int(label), x.size(0), self._noise(noise, keep), seed, counter,
                                  flags, _ptr(out), _stream()))
        return out

    def step_c(self, x_real, label: int, noise=None, seed=0, counter=0, flags=0, loss_out=None):
        x = self._x(x_real)
        keep = [x]
        out = self.loss_buf if loss_out is None else loss_out
        check(self.lib.cvg_step_c(self.h, _ptr(x), int(label), x.size(0), self._noise(noise, keep), seed, counter,
                                  flags, _ptr(out), _stream()))
        return out

    def step_classifier(self, x, labels, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, noise=None, seed=0, counter=0, flags=0,
                        loss_out=None):
        """One fine-tuning step of the classifier on (x[B,F], labels[B]) - classifier.py:34-44."""
        x = self._x(x)
        labels = labels.to(self.device, torch.int64).contiguous()
        keep = [x, labels]
        out = self.loss_buf if loss_out is None else loss_out
        check(self.lib.cvg_step_classifier(self.h, _ptr(x), _ptr(labels), x.size(0), self._noise(noise, keep), seed, counter,
                                           float(lr), float(betas[0]), float(betas[1]), float(eps), flags, _ptr(out),
                                           _stream()))
        return out

    def reset_adam(self, net: int):
        """Fresh torch.optim.Adam state for one network (zero moments, step 0)."""
        self.adam_m[net].zero_()
        self.adam_v[net].zero_()
        self.set_adam_step(net, 0)

    def step_g(self, x_real, label: int, lambda_class: float, noise=None, seed=0, counter=0, flags=0, loss_out=None):
        x = self._x(x_real)
        keep = [x]
        out = self.loss_buf if loss_out is None else loss_out
        check(self.lib.cvg_step_g(self.h, _ptr(x), int(label), x.size(0), self._noise(noise, keep), seed, counter,
                                  float(lambda_class), flags, _ptr(out), _stream()))
        return out

    def step_g_prior(self, batch: int, label: int, lambda_class: float, noise=None, seed=0, counter=0, flags=0, loss_out=None):
        """The sibling trainer CGAN's generator step (src/cgan.py:138-178): no real batch, generator update only."""
        from ._lib import STEP_PRIOR_ONLY
        keep = []
        out = self.loss_buf if loss_out is None else loss_out
        check(self.lib.cvg_step_g(self.h, None, int(label), int(batch), self._noise(noise, keep), seed, counter,
                                  float(lambda_class), flags | STEP_PRIOR_ONLY, _ptr(out), _stream()))
        return out

    def step_g_cvae(self, x_real, label: int, lambda_class: float, noise=None, seed=0, counter=0, flags=0, loss_out=None):
        """The sibling trainer CVAE's encoder/generator step (src/cvae.py:117-166): x_recon = G(E(x)) only, the classification
        term on x_recon, no critic; loss_out = {recon, kl, 0, class}."""
        from ._lib import STEP_CVAE
        return self.step_g(x_real, label, lambda_class, noise=noise, seed=seed, counter=counter, flags=flags | STEP_CVAE,
                           loss_out=loss_out)

    def ctl_set(self, seed=None, counter=None, lambda_class=None):
        """Write the device control block (Philox key/counter and/or this epoch's lambda_class)."""
        set_rng = seed is not None
        set_l = lambda_class is not None
        check(self.lib.cvg_ctl_set(self.h, int(seed or 0), int(counter or 0), 1 if set_rng else 0,
                                   float(lambda_class or 0.0), 1 if set_l else 0, _stream()))

    def visit(self, label: int, batch_global: int, class_rows=None, x_batches=None, loops=(5, 5, 3), flags=0,
              loss_out=None):
        """One label visit (capturable in a CUDA graph): loops = (d_loop, c_loop, g_loop)."""
        b_local = batch_global // self.world_size
        n = sum(loops)
        if loss_out is None:
            loss_out = torch.zeros(n, 4, dtype=torch.float32, device=self.device)
        check(self.lib.cvg_visit(self.h, int(label), b_local, int(batch_global), _ptr(class_rows),
                                 0 if class_rows is None else class_rows.size(0), _ptr(x_batches),
                                 int(loops[0]), int(loops[1]), int(loops[2]), int(flags), _ptr(loss_out), _stream()))
        return loss_out

    def adam(self, net_mask: int):
        check(self.lib.cvg_adam(self.h, int(net_mask), _stream()))

    def sample_rows(self, class_rows: torch.Tensor, batch_global: int, seed=0, counter=0, want_idx=False):
        """cvae_gan.py:247-260 on the device; returns this rank's shard [batch_global / world, F]."""
        b_local = batch_global // self.world_size
        x = torch.empty(b_local, self.F, dtype=torch.float32, device=self.device)
        idx = torch.empty(b_local, dtype=torch.int64, device=self.device) if want_idx else None
        check(self.lib.cvg_sample_rows(self.h, _ptr(class_rows), class_rows.size(0), batch_global,
                                       self.rank * b_local, b_local, seed, counter, _ptr(x), _ptr(idx), _stream()))
        return (x, idx) if want_idx else x

    # ---- generation ------------------------------------------------------------------------------------
    def generate(self, label: int, n: int, z: Optional[torch.Tensor] = None, seed=0, row_offset=0, train_mode=False):
        out = torch.empty(n, self.F, dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        if z is not None and (tuple(z.shape) != (n, self.Z) or not z.is_cuda or not z.is_contiguous()):
            raise ValueError(f"z must be a contiguous CUDA [n, {self.Z}] tensor")
        check(self.lib.cvg_generate(self.h, int(label), n, _ptr(z), seed, row_offset, 1 if train_mode else 0,
                                    _ptr(out), _stream()))
        return out

    def generate_filter(self, label: int, n: int, thr: float, z: Optional[torch.Tensor] = None, seed=0, row_offset=0,
                        capacity: Optional[int] = None, want_logits=False, want_keep=False):
        """Returns (x_out[capacity,F], idx_out[capacity], count tensor (device int64), logits|None, keep|None).
        Rows of x_out/idx_out beyond min(count, capacity) are undefined."""
        capacity = n if capacity is None else int(capacity)
        x_out = torch.empty(capacity, self.F, dtype=torch.float32, device=self.device)
        idx_out = torch.empty(capacity, dtype=torch.int64, device=self.device)
        logits = torch.empty(n, self.K, dtype=torch.float32, device=self.device) if want_logits else None
        keep = torch.empty(n, dtype=torch.uint8, device=self.device) if want_keep else None
        self.count_buf.zero_()
        if z is not None and (tuple(z.shape) != (n, self.Z) or not z.is_cuda or not z.is_contiguous()):
            raise ValueError(f"z must be a contiguous CUDA [n, {self.Z}] tensor")
        check(self.lib.cvg_generate_filter(self.h, int(label), n, float(thr), _ptr(z), seed, row_offset, _ptr(x_out),
                                           _ptr(idx_out), capacity, _ptr(self.count_buf), _ptr(logits), _ptr(keep),
                                           _stream()))
        return x_out, idx_out, self.count_buf, logits, keep

    def filter_logits(self, logits: torch.Tensor, label: int, thr: float) -> torch.Tensor:
        logits = logits.to(self.device, torch.float32).contiguous()
        keep = torch.empty(logits.size(0), dtype=torch.uint8, device=self.device)
        check(self.lib.cvg_filter_logits(_ptr(logits), logits.size(0), logits.size(1), int(label), float(thr),
                                         _ptr(keep), _stream()))
        return keep.bool()

    def filter_compact(self, x: torch.Tensor, logits: torch.Tensor, label: int, thr: float, row_offset=0, capacity=None):
        n = x.size(0)
        capacity = n if capacity is None else int(capacity)
        x_out = torch.empty(capacity, x.size(1), dtype=torch.float32, device=self.device)
        idx_out = torch.empty(capacity, dtype=torch.int64, device=self.device)
        self.count_buf.zero_()
        check(self.lib.cvg_filter_compact(_ptr(x), _ptr(logits), n, x.size(1), logits.size(1), int(label), float(thr),
                                          row_offset, _ptr(x_out), _ptr(idx_out), capacity, _ptr(self.count_buf),
                                          _stream()))
        return x_out, idx_out, self.count_buf

    def classifier_forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(self.device, torch.float32).contiguous()
        out = torch.empty(x.size(0), self.K, dtype=torch.float32, device=self.device)
        if x.size(0):
            check(self.lib.cvg_classifier_forward(self.h, _ptr(x), x.size(0), _ptr(out), _stream()))
        return out

    def encoder_forward(self, x: torch.Tensor, label: int):
        x = x.to(self.device, torch.float32).contiguous()
        mu = torch.empty(x.size(0), self.Z, dtype=torch.float32, device=self.device)
        lv = torch.empty_like(mu)
        if x.size(0):
            check(self.lib.cvg_encoder_forward(self.h, _ptr(x), int(label), x.size(0), _ptr(mu), _ptr(lv), _stream()))
        return mu, lv

    def debug_read(self, name: str, rows: int, pas: int = 0) -> torch.Tensor:
        """Test hook: one workspace matrix of the last step as a row-major [rows, features] tensor."""
        c = C.c_int()
        check(self.lib.cvg_debug_read(self.h, name.encode(), pas, rows, None, C.byref(c), _stream()))
        out = torch.empty(rows, c.value, dtype=torch.float32, device=self.device)
        check(self.lib.cvg_debug_read(self.h, name.encode(), pas, rows, _ptr(out), C.byref(c), _stream()))
        return out

    def profile(self, enable: bool):
        check(self.lib.cvg_profile_enable(self.h, 1 if enable else 0))

    def profile_read(self):
        """{class: (launches, algorithmic flops, device ms)} for fwd / dx / dw GEMM launches."""
        out = {}
        for cls, name in enumerate(("gemm_fwd", "gemm_dx", "gemm_dw")):
            n, f, ms = C.c_int64(), C.c_double(), C.c_double()
            check(self.lib.cvg_profile_read(self.h, cls, C.byref(n), C.byref(f), C.byref(ms)))
            out[name] = (n.value, f.value, ms.value)
        return out

    def launch_count(self) -> int:
        return int(self.lib.cvg_launch_count(self.h))

    def debug_set(self, key: str, value: int):
        """Executor switches: train_mode (1 = step-program tcgen05 kernel, 0 = stand-alone FFMA kernels), mk_max_ops, ..."""
        check(self.lib.cvg_debug_set(self.h, key.encode(), int(value)))

    def debug_get(self, key: str) -> int:
        v = C.c_int()
        check(self.lib.cvg_debug_get(self.h, key.encode(), C.byref(v)))
        return v.value

    def mk_cycles(self):
        """Per-op cycle counts of the last step program (CVG_MK_DBG=1)."""
        n = C.c_int()
        buf = (C.c_longlong * 4096)()
        check(self.lib.cvg_debug_mk_cycles(self.h, buf, 4096, C.byref(n)))
        self.mk_sections = list(buf[2048:2048 + 16])
        self.mk_starts = list(buf[1024:1024 + n.value])
        self.mk_ops = [(int(v) & 255, (int(v) >> 8) & 1, int(v) >> 16) for v in buf[3072:3072 + n.value]]
        return list(buf[:n.value])


def patience_scan(keep_host: torch.Tensor, num: int, chunk: int = 10, patience: int = 20):
    """Host integer logic of cvae_gan.py:350-376 (see cvg_patience_scan)."""
    lib = _lib.load()
    k = keep_host.to(torch.uint8).contiguous().cpu()
    rc, ra = C.c_int64(), C.c_int64()
    check(lib.cvg_patience_scan(C.c_void_p(k.data_ptr()), k.numel(), int(num), int(chunk), int(patience),
                                C.byref(rc), C.byref(ra)))
    return rc.value, ra.value
