"""SAM, EMA and the fused AdamW on top of the multi-tensor kernels (csrc/optim.cu).

`SAM` and `EMAModel` keep the constructor / method surface of the reference's
`regularization/sam.py` and `utils/ema.py`, so `train/train_transformer.py` drives them unchanged; the
per-parameter Python loops, `.clone()`s, host syncs and the D2H copy of the whole state_dict that the
reference performs every step are replaced by a handful of kernel launches over pointer tables.
"""
from __future__ import annotations

import copy
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Type

import torch

from . import _lib

F32 = torch.float32


def build_chunks(sizes: Sequence[int], chunk: int) -> Tuple[List[int], List[int]]:
    """Host-side chunk table: every tensor is cut into pieces of `chunk` elements.
    Returns (chunk_tensor, chunk_offset)."""
    ct, co = [], []
    for t, n in enumerate(sizes):
        off = 0
        while off < n:
            ct.append(t)
            co.append(off)
            off += chunk
    return ct, co


class MultiTensorTable:
    """Device-side description of a list of same-shaped tensor lists (sizes + chunk table)."""

    def __init__(self, like: Sequence[torch.Tensor], device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("vsn_b200 multi-tensor kernels need CUDA tensors (there is no CPU fallback)")
        _lib.check_device()
        self.n = len(like)
        self.sizes_host = [int(t.numel()) for t in like]
        chunk = _lib.load().vsn_mt_chunk_elems()
        ct, co = build_chunks(self.sizes_host, chunk)
        self.n_chunks = len(ct)
        self.sizes = torch.tensor(self.sizes_host, dtype=torch.int64, device=self.device)
        self.chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=self.device)
        self.chunk_off = torch.tensor(co, dtype=torch.int64, device=self.device)

    def ptr_array(self, tensors: Sequence[torch.Tensor]) -> torch.Tensor:
        """Device array of the tensors' addresses.  Uploaded from a pinned staging ring with a non-blocking copy:
        a `torch.tensor(list, device=...)` would copy from pageable memory, which waits for all queued GPU work --
        a hidden device synchronisation every time gradients are re-allocated (e.g. after zero_grad(set_to_none))."""
        assert len(tensors) == self.n
        for t, n in zip(tensors, self.sizes_host):
            if t.numel() != n or not t.is_contiguous() or not t.is_cuda:
                raise RuntimeError("multi-tensor kernels need contiguous CUDA tensors of matching sizes")
        if not hasattr(self, "_stage"):
            self._stage = [(torch.empty(self.n, dtype=torch.int64).pin_memory(), torch.cuda.Event()) for _ in range(4)]
            self._stage_i = 0
        host, ev = self._stage[self._stage_i]
        self._stage_i = (self._stage_i + 1) % len(self._stage)
        ev.synchronize()                                   # the copy that last used this staging buffer is done
        host.copy_(torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64))
        dev = torch.empty(self.n, dtype=torch.int64, device=self.device)
        dev.copy_(host, non_blocking=True)
        ev.record(torch.cuda.current_stream(self.device))
        return dev

    def _tab(self):
        return self.sizes.data_ptr(), self.chunk_tensor.data_ptr(), self.chunk_off.data_ptr(), self.n_chunks

    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    def _work(self, arrays: int, elem_bytes: float = 4.0) -> None:
        """Algorithmic bytes of the next launch (bench.py's instrumented pass): `arrays` passes over all elements."""
        if _lib.PROFILE is not None:
            tot = sum(self.sizes_host)
            _lib.TAG = f"tensors{self.n} elems{tot}"
            _lib.WORK = (0, int(tot * elem_bytes * arrays))

    def sqnorm(self, g_ptrs, p_ptrs, sq, adaptive):
        self._work(2 if adaptive else 1)
        _lib.call("vsn_mt_sqnorm", g_ptrs.data_ptr(), p_ptrs.data_ptr() if p_ptrs is not None else None, *self._tab(),
                  sq.data_ptr(), int(adaptive), self._stream())

    def sam_perturb(self, p_ptrs, g_ptrs, old_ptrs, scale, flags, adaptive):
        self._work(4)                    # read p, g; write old_p, p  (16 B per parameter)
        _lib.call("vsn_mt_sam_perturb", p_ptrs.data_ptr(), g_ptrs.data_ptr(), old_ptrs.data_ptr(), *self._tab(),
                  scale.data_ptr(), flags.data_ptr(), int(adaptive), self._stream())

    def copy(self, dst_ptrs, src_ptrs):
        self._work(2)
        _lib.call("vsn_mt_copy", dst_ptrs.data_ptr(), src_ptrs.data_ptr(), *self._tab(), self._stream())

    def cast_bf16(self, src_ptrs, dst_ptrs):
        self._work(1, 6.0)               # read fp32, write bf16
        _lib.call("vsn_mt_cast_bf16", src_ptrs.data_ptr(), dst_ptrs.data_ptr(), *self._tab(), self._stream())

    def adamw(self, p_ptrs, g_ptrs, m_ptrs, v_ptrs, wd, ctl, lr, beta1, beta2, eps, zero_grad):
        self._work(7 + int(zero_grad))   # read p, g, m, v; write p, m, v (+ the cleared gradient): 28-32 B per parameter
        _lib.call("vsn_mt_adamw", p_ptrs.data_ptr(), g_ptrs.data_ptr(), m_ptrs.data_ptr(), v_ptrs.data_ptr(),
                  *self._tab(), wd.data_ptr(), ctl.data_ptr(), float(lr), float(beta2), 1.0 - float(beta1),
                  1.0 - float(beta2), float(eps), int(zero_grad), self._stream())

    def ema(self, p_ptrs, new_ptrs, s0_ptrs, s1_ptrs, ema_ptrs, w0, w1, w2):
        self._work(3 + (s0_ptrs is not None) + (s1_ptrs is not None))   # read p (+s0, s1); write the new slot and the average
        _lib.call("vsn_mt_ema", p_ptrs.data_ptr(), new_ptrs.data_ptr(),
                  s0_ptrs.data_ptr() if s0_ptrs is not None else None,
                  s1_ptrs.data_ptr() if s1_ptrs is not None else None, ema_ptrs.data_ptr(), *self._tab(),
                  float(w0), float(w1), float(w2), self._stream())


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW (train/train_transformer.py:2125-2147) as ONE multi-tensor launch per step.

    Same constructor, same update rule and the same per-parameter state (`step`, `exp_avg`, `exp_avg_sq`), so
    `state_dict()` / `load_state_dict()` interchange with torch.optim.AdamW checkpoints
    (train/train_transformer.py:752-820).  What is fused into the pass: the GradScaler's unscale and inf check
    (`scaler.step(opt)` hands `grad_scale` / `found_inf` over as device tensors, :1203-1232 -- a skipped step costs no
    host synchronisation: the step counter lives on the device too) and, with `step(zero_grad=True)`, the gradient
    clear the next accumulation pass needs.  28 B (+4) per parameter, once.  `amsgrad` / `maximize` are not built."""

    _step_supports_amp_scaling = True
    fused_zero_grad = True

    def __init__(self, params, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, amsgrad: bool = False, *, maximize: bool = False, foreach=None,
                 capturable: bool = False, differentiable: bool = False, fused=None):
        if amsgrad or maximize or differentiable:
            raise NotImplementedError("vsn_b200 FusedAdamW: amsgrad / maximize / differentiable are not built")
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError("invalid AdamW hyper-parameter")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        self._plans: Dict[tuple, dict] = {}
        self._step_dev: Optional[torch.Tensor] = None      # one counter for all parameters (they always step together)
        self._ctl: Optional[torch.Tensor] = None

    # -- state ------------------------------------------------------------------------------------
    def _init_state(self, p: torch.nn.Parameter) -> None:
        st = self.state[p]
        if "exp_avg" not in st:
            st["step"] = torch.zeros((), dtype=F32)
            st["exp_avg"] = torch.zeros_like(p.detach(), memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p.detach(), memory_format=torch.contiguous_format)

    def _device_step(self, dev) -> torch.Tensor:
        if self._step_dev is None:
            steps = {float(st["step"]) for st in self.state.values() if "step" in st}
            if len(steps) > 1:
                raise NotImplementedError("vsn_b200 FusedAdamW keeps one step count: the loaded state has several")
            self._step_dev = torch.full((1,), steps.pop() if steps else 0.0, device=dev, dtype=F32)
            self._ctl = torch.zeros(4, device=dev, dtype=F32)
        return self._step_dev

    def state_dict(self) -> Dict[str, Any]:
        if self._step_dev is not None:                      # checkpoint time: one host read of the counter
            n = float(self._step_dev.item())
            for st in self.state.values():
                if "step" in st:
                    st["step"] = torch.tensor(n, dtype=F32)
        return super().state_dict()

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        super().load_state_dict(state_dict)
        self._plans, self._step_dev, self._ctl = {}, None, None

    # -- one launch per set of groups that share (lr, betas, eps) ------------------------------------
    def _plan(self, ps: List[torch.nn.Parameter], wds: List[float]) -> dict:
        key = tuple((p.data_ptr(), p.numel(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr())
                    for p in ps)
        pl = self._plans.get(key)
        if pl is None:
            for p in ps:
                if p.dtype != F32 or not p.is_cuda or not p.is_contiguous():
                    raise RuntimeError("vsn_b200 FusedAdamW needs contiguous fp32 CUDA parameters (there is no CPU fallback)")
            dev = ps[0].device
            tab = MultiTensorTable([p.detach() for p in ps], dev)
            ms = [self.state[p]["exp_avg"] for p in ps]
            vs = [self.state[p]["exp_avg_sq"] for p in ps]
            pl = dict(tab=tab, p=tab.ptr_array([p.detach() for p in ps]), m=tab.ptr_array(ms), v=tab.ptr_array(vs),
                      keep=(ms, vs), gkey=None, g=None, wd=torch.tensor(wds, device=dev, dtype=F32), wds=list(wds))
            if len(self._plans) > 8:
                self._plans.clear()
            self._plans[key] = pl
        if pl["wds"] != list(wds):
            pl["wd"].copy_(torch.tensor(wds, dtype=F32), non_blocking=True)
            pl["wds"] = list(wds)
        gkey = tuple(p.grad.data_ptr() for p in ps)
        if gkey != pl["gkey"]:
            for p in ps:
                if p.grad.dtype != F32 or not p.grad.is_contiguous():
                    raise RuntimeError("vsn_b200 FusedAdamW needs contiguous fp32 gradients")
            pl["g"] = pl["tab"].ptr_array([p.grad for p in ps])
            pl["gkey"] = gkey
        return pl

    @torch.no_grad()
    def step(self, closure: Optional[Callable[[], Any]] = None, *, zero_grad: bool = False):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        sets: Dict[tuple, Tuple[list, list]] = {}
        for group in self.param_groups:
            hp = (float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]))
            ps, wds = sets.setdefault(hp, ([], []))
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("AdamW does not support sparse gradients")
                self._init_state(p)
                ps.append(p)
                wds.append(float(group["weight_decay"]))
        sets = {hp: v for hp, v in sets.items() if v[0]}
        if not sets:
            return loss
        betas = {(hp[1], hp[2]) for hp in sets}
        if len(betas) > 1:
            raise NotImplementedError("vsn_b200 FusedAdamW keeps one step count / bias correction: one betas pair")
        dev = next(iter(sets.values()))[0][0].device
        step = self._device_step(dev)
        found_inf = getattr(self, "found_inf", None)        # set by GradScaler.step (torch/amp/grad_scaler.py)
        grad_scale = getattr(self, "grad_scale", None)
        inv = None
        if grad_scale is not None:
            inv = grad_scale.to(device=dev, dtype=F32).reciprocal().reshape(1)
        if found_inf is not None:
            found_inf = found_inf.to(device=dev, dtype=F32).reshape(1)
        b1, b2 = next(iter(betas))
        _lib.call("vsn_adamw_prepare", step.data_ptr(), found_inf.data_ptr() if found_inf is not None else None,
                  inv.data_ptr() if inv is not None else None, b1, b2, self._ctl.data_ptr(),
                  torch.cuda.current_stream().cuda_stream)
        for (lr, _, _, eps), (ps, wds) in sets.items():
            pl = self._plan(ps, wds)
            pl["tab"].adamw(pl["p"], pl["g"], pl["m"], pl["v"], pl["wd"], self._ctl, lr, b1, b2, eps, zero_grad)
        return loss


class SAM(torch.optim.Optimizer):
    """Sharpness-Aware Minimization with the reference's API (regularization/sam.py:9-165).

    first_step: global gradient norm + `old_p = p; p += rho/(||g||+1e-12) * g` in three launches, no host
    synchronisation.  second_step: `p = old_p` (one launch) then the base optimiser's step.
    Non-finite handling follows the reference: tensors with a non-finite norm are left out of the norm and
    not perturbed; a zero norm disables the perturbation.  (Where the reference returns early without saving
    `old_p` and would later restore a stale copy, this implementation saves `old_p = p`, so second_step is
    always well defined.)"""

    def __init__(self, params, base_optimizer: Type[torch.optim.Optimizer], rho: float = 0.05,
                 adaptive: bool = False, **kwargs):
        assert rho >= 0.0, f"Invalid rho, should be non-negative: {rho}"
        defaults = dict(rho=rho, adaptive=adaptive, **kwargs)
        super().__init__(params, defaults)
        self.base_optimizer = base_optimizer(self.param_groups, **kwargs)
        self.param_groups = self.base_optimizer.param_groups
        self.defaults.update(self.base_optimizer.defaults)
        self._plan_key = None
        self._plan = None

    # -- plan: pointer tables for the parameters that currently have a gradient ------------------
    def _active(self) -> List[torch.nn.Parameter]:
        ps = [p for g in self.param_groups for p in g["params"] if p.grad is not None]
        rhos = {g["rho"] for g in self.param_groups}
        ads = {bool(g["adaptive"]) for g in self.param_groups}
        if len(rhos) > 1 or len(ads) > 1:
            raise NotImplementedError("vsn_b200 SAM expects one rho / adaptive setting for all parameter groups")
        return ps

    def _get_plan(self, ps: List[torch.nn.Parameter]):
        # the key covers the saved-weights tensors too: if optimizer.state was replaced (load_state_dict, a cleared
        # state) the cached addresses would point at freed memory
        key = tuple((p.data_ptr(), p.numel(), self.state[p]["old_p"].data_ptr() if "old_p" in self.state[p] else 0)
                    for p in ps)
        if key != self._plan_key:
            for p in ps:
                if p.dtype != F32 or not p.is_cuda:
                    raise RuntimeError("vsn_b200 SAM needs fp32 CUDA parameters (there is no CPU fallback)")
                if not p.is_contiguous():
                    raise RuntimeError("vsn_b200 SAM needs contiguous parameters")
            dev = ps[0].device
            tab = MultiTensorTable([p.detach() for p in ps], dev)
            olds = []
            for p in ps:
                st = self.state[p]
                o = st.get("old_p")
                if (o is None or o.shape != p.shape or o.device != p.device or o.dtype != F32
                        or not o.is_contiguous()):
                    st["old_p"] = torch.empty_like(p.detach(), memory_format=torch.contiguous_format)
                olds.append(st["old_p"])
            key = tuple((p.data_ptr(), p.numel(), o.data_ptr()) for p, o in zip(ps, olds))
            self._plan = dict(tab=tab, p=tab.ptr_array([p.detach() for p in ps]), old=tab.ptr_array(olds),
                              olds=olds,           # referenced here so the addresses in `old` stay valid
                              gkey=None, g=None,
                              sq=torch.zeros(len(ps), device=dev, dtype=F32),
                              flags=torch.zeros(len(ps), device=dev, dtype=torch.int32),
                              scale=torch.zeros(2, device=dev, dtype=F32))
            self._plan_key = key
        pl = self._plan
        gkey = tuple(p.grad.data_ptr() for p in ps)
        if gkey != pl["gkey"]:
            pl["g"] = pl["tab"].ptr_array([p.grad for p in ps])
            pl["gkey"] = gkey
        return pl

    @torch.no_grad()
    def _grad_norm(self) -> torch.Tensor:
        ps = self._active()
        if not ps:
            dev = self.param_groups[0]["params"][0].device
            return torch.tensor(1e-12, device=dev)
        pl = self._get_plan(ps)
        self._launch_norm(pl)
        return pl["scale"][1].clone()

    def _launch_norm(self, pl) -> None:
        adaptive = bool(self.param_groups[0]["adaptive"])
        pl["sq"].zero_()
        pl["tab"].sqnorm(pl["g"], pl["p"] if adaptive else None, pl["sq"], adaptive)
        _lib.call("vsn_sam_scale", pl["sq"].data_ptr(), len(pl["sq"]), float(self.param_groups[0]["rho"]),
                  pl["scale"].data_ptr(), pl["flags"].data_ptr(), torch.cuda.current_stream().cuda_stream)

    @torch.no_grad()
    def first_step(self, zero_grad: bool = False) -> None:
        ps = self._active()
        if ps:
            pl = self._get_plan(ps)
            self._launch_norm(pl)
            pl["tab"].sam_perturb(pl["p"], pl["g"], pl["old"], pl["scale"], pl["flags"],
                                  bool(self.param_groups[0]["adaptive"]))
            self._perturbed = ps
        else:
            self._perturbed = []
        if zero_grad:
            self.zero_grad()

    @torch.no_grad()
    def second_step(self, zero_grad: bool = False, scaler=None) -> None:
        ps = [p for g in self.param_groups for p in g["params"] if p.grad is not None and "old_p" in self.state[p]]
        if ps:
            pl = self._get_plan(ps)
            pl["tab"].copy(pl["p"], pl["old"])     # back to "w" from "w + e(w)"
        if scaler is None:
            if zero_grad and getattr(self.base_optimizer, "fused_zero_grad", False):
                self.base_optimizer.step(zero_grad=True)      # gradients cleared inside the AdamW pass
                return
            self.base_optimizer.step()
            if zero_grad:
                self.zero_grad()
        else:
            scaler.step(self.base_optimizer)
            scaler.update()
            if zero_grad:
                self.zero_grad()

    @torch.no_grad()
    def step(self, closure: Optional[Callable[[], Any]] = None) -> None:
        assert closure is not None, "Sharpness Aware Minimization requires closure, but it was not provided"
        closure = torch.enable_grad()(closure)
        self.first_step(zero_grad=True)
        closure()
        self.second_step()

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        super().load_state_dict(state_dict)
        self.base_optimizer.param_groups = self.param_groups
        self._plan_key, self._plan = None, None     # state (old_p) was replaced: rebuild the pointer tables


class EMAModel:
    """Weighted average of the last `n_models` parameter snapshots, API of utils/ema.py:10-178.

    The reference copies the whole state_dict to the host every step and averages there; here the snapshot
    ring and the average live on the device and one kernel both takes the new snapshot and recomputes the
    average (`ema = sum_i w_i * s_i`, `w_i = decay^(k-1-i) / sum_j decay^j`, oldest first, fma order of
    `Tensor.add_(state, alpha=w)`).  Non-floating buffers follow the newest snapshot.  `model_state` is the
    same name -> tensor mapping (device tensors) and is what the trainer stores as the checkpoint's "model"."""

    def __init__(self, model: torch.nn.Module, decay: float, n_models: int = 3,
                 device: Optional[torch.device] = None):
        if not 1 <= n_models <= 3:
            raise NotImplementedError("vsn_b200 EMAModel supports n_models in 1..3 (the trainer always uses 3)")
        self.decay = decay
        self.n_models = n_models
        self.device = device or next(model.parameters()).device
        sd = model.state_dict()
        self._fkeys = [k for k, v in sd.items() if v.is_floating_point()]
        self._okeys = [k for k, v in sd.items() if not v.is_floating_point()]
        live = [sd[k] for k in self._fkeys]
        for v in live:
            if not v.is_cuda or v.dtype != F32:
                raise RuntimeError("vsn_b200 EMAModel needs an fp32 CUDA model (there is no CPU fallback)")
        self._table = MultiTensorTable(live, live[0].device)
        mk = lambda: [torch.empty_like(v, memory_format=torch.contiguous_format) for v in live]  # noqa: E731
        self._slots = [mk()]
        self._slot_ptrs = [self._table.ptr_array(self._slots[0])]
        self._order = [0]                       # live snapshots, oldest first (indices into _slots)
        ema = mk()
        self._ema_ptrs = self._table.ptr_array(ema)
        self.model_state: Dict[str, torch.Tensor] = {}
        for k, e in zip(self._fkeys, ema):
            self.model_state[k] = e
        for k in self._okeys:
            self.model_state[k] = sd[k].detach().clone()
        # initial state counts as the first snapshot (utils/ema.py:41-48)
        self._live_key = None
        self._table.ema(self._live_ptrs(sd), self._slot_ptrs[0], None, None, self._ema_ptrs, 0.0, 0.0, 1.0)
        self.collected_params: dict = {}
        self.orig_state = None
        self._stash = None

    def _live_ptrs(self, sd) -> torch.Tensor:
        key = tuple(sd[k].data_ptr() for k in self._fkeys)
        if key != self._live_key:
            self._live_key = key
            self._live = self._table.ptr_array([sd[k].detach() for k in self._fkeys])
        return self._live

    @torch.no_grad()
    def update(self, model: torch.nn.Module) -> None:
        sd = model.state_dict()
        if len(self._order) < self.n_models:
            self._slots.append([torch.empty_like(v) for v in self._slots[0]])
            self._slot_ptrs.append(self._table.ptr_array(self._slots[-1]))
            new = len(self._slots) - 1
            older = list(self._order)
        else:
            new = self._order[0]                # evicted snapshot's storage is recycled
            older = self._order[1:]
        k = len(older) + 1
        w = [self.decay ** i for i in range(k)][::-1]
        tot = sum(w)
        w = [x / tot for x in w]
        s0 = self._slot_ptrs[older[-2]] if len(older) >= 2 else None
        s1 = self._slot_ptrs[older[-1]] if len(older) >= 1 else None
        w0 = w[-3] if len(older) >= 2 else 0.0
        w1 = w[-2] if len(older) >= 1 else 0.0
        self._table.ema(self._live_ptrs(sd), self._slot_ptrs[new], s0, s1, self._ema_ptrs, w0, w1, w[-1])
        self._order = older + [new]
        for name in self._okeys:
            self.model_state[name].copy_(sd[name])

    @torch.no_grad()
    def apply_to(self, model: torch.nn.Module) -> None:
        if self.model_state is None:
            return
        sd = model.state_dict()
        if self._stash is None:
            self._stash = [torch.empty_like(v) for v in self._slots[0]]
            self._stash_ptrs = self._table.ptr_array(self._stash)
        live = self._live_ptrs(sd)
        self._table.copy(self._stash_ptrs, live)
        self.orig_state = {k: s for k, s in zip(self._fkeys, self._stash)}
        for k in self._okeys:
            self.orig_state[k] = sd[k].detach().clone()
        self._table.copy(live, self._ema_ptrs)
        for k in self._okeys:
            sd[k].copy_(self.model_state[k])

    @torch.no_grad()
    def restore(self, model: torch.nn.Module) -> None:
        if self.orig_state is None:
            return
        sd = model.state_dict()
        self._table.copy(self._live_ptrs(sd), self._stash_ptrs)
        for k in self._okeys:
            sd[k].copy_(self.orig_state[k])
        self.orig_state = None

    @torch.no_grad()
    def update_bn_stats(self, model: torch.nn.Module, data_loader) -> None:
        """utils/ema.py:144-178 — only meaningful for BatchNorm models (not on the Swin/ViT path)."""
        original_state = copy.deepcopy(model.state_dict())
        self.apply_to(model)
        model.train()
        for x, _ in data_loader:
            model(x.to(next(model.parameters()).device))
        for name, param in model.state_dict().items():
            if "running_" in name or "batch_norm" in name:
                self.model_state[name].copy_(param)
        model.load_state_dict(original_state)
        self.orig_state = None
