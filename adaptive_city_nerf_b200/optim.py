"""Optimizer tail of a training step behind the reference's own entry points.

Reference: common/utils.py:16-76 `get_optimizer` (Adam / AdamW over the model's named parameter groups),
pipelines/offline_stage/meta_core.py:123-141 `maml_meta_update` and :181-190 `clip_all_grads`,
pipelines/online_stage/runtime_adapt.py:262-268 -- i.e. per step

    scaler.unscale_(opt); clip_grad_norm_(params, c); scaler.step(opt); scaler.update()

which in PyTorch is 12 passes over the 64 MiB hash table per expert and one host read (the inf check).  `FusedAdam`
does it in 8 passes and three launches (`acn_grad_sqnorm`, `acn_adam_prepare`, `acn_adam_apply`) and reads nothing
back: skip / clip coefficient / bias corrections are decided on the device.  The arithmetic per element is
torch.optim.Adam's (same op order, constants rounded from the same doubles).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional

import torch

from . import _lib
from ._lib import AdamTensor, check, ctx, lib, ptr, stream


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam (adamw=False) / AdamW (adamw=True) with the gradient unscale, the global-norm clip and the
    non-finite skip folded into the update.

    Three ways to drive it, all without a host sync:
      opt.step(max_norm=c)                      fp32 training: clip + step
      scaler.step(opt, max_norm=c); scaler.update()    public GradScaler protocol (the scaler hands over its scale and
                                                its own inf check through `opt.grad_scale` / `opt.found_inf`)
      opt.step_scaled(scaler, max_norm=c)       everything in the three launches, including the inf check; also
                                                performs `scaler.update()`
    `opt.last_norm` is the unscaled global gradient norm of the last step (a device scalar, what clip_grad_norm_
    returns).

    Step counts are per parameter, as in torch.optim.Adam: a parameter whose `.grad` is None on a step (an expert that
    received no ray after `zero_grad(set_to_none=True)`) keeps its count, so its bias corrections pick up where they
    left off; `state[p]["step"]` is that count (a device scalar) and round-trips through `state_dict()`.

    CUDA graphs: lr, weight decay, betas, eps and max_norm are kernel ARGUMENTS -- a captured step replays the values
    it was captured with (an lr scheduler has no effect on replays; re-capture after changing them)."""

    _step_supports_amp_scaling = True

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 adamw: bool = False, max_norm: Optional[float] = None, write_grads: bool = False,
                 skip_zero_grads: bool = False, norm_group=None):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError(f"invalid Adam hyper-parameters lr={lr} betas={betas} eps={eps} weight_decay={weight_decay}")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.adamw, self.max_norm, self.write_grads = bool(adamw), max_norm, bool(write_grads)
        #: treat a parameter whose gradient is identically zero like torch treats `grad is None` (leave it and its step
        #: count alone).  The container's sync-free routed path cannot return None for an expert that received no row
        #: (it would have to read the row count back); it returns zeros, and this flag restores the reference's behaviour
        #: (zero_grad(set_to_none=True) + experts without rays keep their Adam state).  `get_optimizer` sets it.
        self.skip_zero_grads = bool(skip_zero_grads)
        #: torch.distributed group over which the squared gradient norm (and the non-finite flag) is summed before the
        #: step is decided: expert-sharded training, where every rank holds different parameters but the reference clips
        #: ONE global norm (meta_core.py:181-190).  Two doubles per step, on the device.
        self.norm_group = norm_group
        self._dev: Optional[torch.device] = None
        self._state8 = self._acc2 = self._found_inf = None
        self._bias: Dict[torch.Tensor, torch.Tensor] = {}      # per parameter: [bias_correction1, sqrt(bias_correction2)] workspace

    # ---- device-side step state ---------------------------------------------------------------------------------
    def _device_state(self, dev: torch.device):
        if self._dev is None:
            self._dev = dev
            self._state8 = torch.zeros(8, dtype=torch.float64, device=dev)
            self._acc2 = torch.zeros(2, dtype=torch.float64, device=dev)
            self._found_inf = torch.zeros(1, dtype=torch.float32, device=dev)
        elif dev != self._dev:
            raise RuntimeError(f"FusedAdam: parameters on {dev} and {self._dev}; use one optimizer per device")

    @property
    def last_norm(self) -> torch.Tensor:
        return self._state8[5].float()

    @property
    def steps_taken(self) -> torch.Tensor:
        """Number of optimizer steps that were not skipped (device scalar; per-parameter counts are in state[p]["step"])."""
        return self._state8[0]

    def _tensors(self):
        betas, eps = self.param_groups[0]["betas"], self.param_groups[0]["eps"]
        out, keep = [], []
        for group in self.param_groups:
            if tuple(group["betas"]) != tuple(betas) or group["eps"] != eps:
                raise RuntimeError("FusedAdam: all parameter groups must share betas and eps (lr and weight_decay may differ)")
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam needs CUDA parameters; there is no CPU path")
                if g.is_sparse or p.dtype != torch.float32 or g.dtype != torch.float32:
                    raise RuntimeError("FusedAdam: dense fp32 parameters and gradients only")
                if not p.is_contiguous() or not g.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters and gradients must be contiguous")
                self._device_state(p.device)
                st = self.state[p]
                if "exp_avg" not in st:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)   # advanced on the device, per parameter
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                bias = self._bias.get(p)
                if bias is None:
                    bias = self._bias[p] = torch.zeros(4, dtype=torch.float64, device=p.device)
                t = AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                               p.numel(), float(group["lr"]), float(group["weight_decay"]), st["step"].data_ptr(),
                               bias.data_ptr())
                out.append(t)
                keep.append((p, g))
        return out, keep, betas, eps

    def _run(self, grad_scale, found_inf_in, max_norm):
        tensors, keep, betas, eps = self._tensors()
        if not tensors:
            return
        dev = self._dev
        c, s, L = ctx(dev), stream(dev), lib()
        chunks = []
        for i in range(0, len(tensors), _lib.ADAM_MAX_TENSORS):
            part = tensors[i:i + _lib.ADAM_MAX_TENSORS]
            chunks.append(((AdamTensor * len(part))(*part), len(part)))
        gs = None if grad_scale is None else grad_scale.reshape(-1)
        if gs is not None and (gs.dtype != torch.float32 or gs.device != dev):
            gs = gs.to(device=dev, dtype=torch.float32)
        fi = None if found_inf_in is None else found_inf_in.reshape(-1).to(device=dev, dtype=torch.float32)
        for arr, n in chunks:
            check(L.acn_grad_sqnorm(c, C.cast(arr, C.c_void_p), n, ptr(gs), ptr(self._acc2), s))
        if self.norm_group is not None:
            import torch.distributed as dist
            dist.all_reduce(self._acc2, group=self.norm_group)
        check(L.acn_adam_prepare(c, ptr(self._acc2), ptr(gs), ptr(fi), float(max_norm) if max_norm else 0.0,
                                 float(betas[0]), float(betas[1]), ptr(self._state8), ptr(self._found_inf),
                                 C.cast(chunks[0][0], C.c_void_p), chunks[0][1], int(self.skip_zero_grads), s))
        for arr, n in chunks[1:]:
            check(L.acn_adam_advance(c, C.cast(arr, C.c_void_p), n, ptr(self._state8), float(betas[0]), float(betas[1]),
                                     int(self.skip_zero_grads), s))
        for arr, n in chunks:
            check(L.acn_adam_apply(c, C.cast(arr, C.c_void_p), n, ptr(self._state8), float(betas[0]), float(betas[1]),
                                   float(eps), int(self.adamw), int(self.write_grads), s))
        del keep

    @torch.no_grad()
    def step(self, closure=None, *, max_norm: Optional[float] = ...):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if max_norm is ...:
            max_norm = self.max_norm
        # torch.amp.GradScaler.step() sets these two attributes on optimizers that declare _step_supports_amp_scaling
        self._run(getattr(self, "grad_scale", None), getattr(self, "found_inf", None), max_norm)
        return loss

    @torch.no_grad()
    def step_scaled(self, scaler, *, max_norm: Optional[float] = ...):
        """`scaler.unscale_(opt); clip_grad_norm_(...); scaler.step(opt); scaler.update()` in three launches plus the
        scale update; call it after `scaler.scale(loss).backward()`.  A disabled scaler degrades to `step()`."""
        if max_norm is ...:
            max_norm = self.max_norm
        if not scaler.is_enabled():
            self._run(None, None, max_norm)
            return
        if scaler._scale is None:
            raise RuntimeError("step_scaled: call scaler.scale(loss).backward() first")
        from torch.amp.grad_scaler import OptState
        st = scaler._per_optimizer_states[id(self)]
        if st["stage"] is OptState.STEPPED:
            raise RuntimeError("step_scaled: this optimizer has already stepped since the last scaler.update()")
        if st["stage"] is OptState.UNSCALED:
            # the caller already ran scaler.unscale_(opt) (maml_meta_update's order): the gradients are real, the
            # scaler's own inf check stands -- do not divide by the scale a second time
            found = [v.to(self._dev or v.device) for v in st["found_inf_per_device"].values()]
            self._run(None, torch.stack(found).sum().reshape(1) if found else None, max_norm)
        else:
            self._run(scaler._scale, None, max_norm)
        if self._found_inf is not None:
            torch._amp_update_scale_(scaler._scale, scaler._growth_tracker, self._found_inf, scaler.get_growth_factor(),
                                     scaler.get_backoff_factor(), scaler.get_growth_interval())
        scaler._per_optimizer_states.pop(id(self), None)       # what scaler.update() does: the next step starts READY

    # ---- checkpoints in torch.optim.Adam's layout (reference: utils.py save/load_checkpoint) ------------------------
    def state_dict(self) -> Dict[str, Any]:
        return super().state_dict()        # state[p]["step"] IS the per-parameter count: torch.optim.Adam's layout

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        """Accepts torch.optim.Adam / AdamW checkpoints: per-parameter `step` (a CPU tensor there) moves to the
        parameter's device and keeps its value."""
        super().load_state_dict(state_dict)
        steps = []
        for p, st in self.state.items():
            if "step" in st:
                st["step"] = torch.as_tensor(st["step"], dtype=torch.float32).to(p.device).reshape(()).clone()
                steps.append(float(st["step"]))
        self._bias.clear()
        if steps:
            p0 = next(p for p in self.state if "step" in self.state[p])
            self._dev = None
            self._device_state(p0.device)
            self._state8[0] = max(steps)


def get_optimizer(P, model: torch.nn.Module) -> FusedAdam:
    """common/utils.py:16-76: parameter groups 'encoding' / 'sigma' / 'color' / 'background' from
    `model.get_param_groups()` with `P.encoding_lr`, `P.sigma_lr`, `P.color_lr`, `P.bg_lr` (fallback `P.lr`)."""
    base_lr = float(getattr(P, "lr", 1e-3))
    weight_decay = float(getattr(P, "weight_decay", 0.0))
    group_dict = model.get_param_groups()
    groups = []
    for name, attr in (("encoding", "encoding_lr"), ("sigma", "sigma_lr"), ("color", "color_lr"), ("background", "bg_lr")):
        if name not in group_dict:
            continue
        assert "params" in group_dict[name], f"{name} group must contain 'params'."
        lr = getattr(P, attr, None)
        groups.append({"params": list(group_dict[name]["params"]), "lr": float(base_lr if lr is None else lr), "name": name})
    opt_name = str(getattr(P, "optimizer", "adamw")).lower()
    if opt_name not in ("adam", "adamw"):
        raise ValueError(f"Unknown optimizer for the fused step: {opt_name} (adam | adamw; use torch.optim for sgd)")
    # a container of several experts renders through the sync-free routed path: "no row for this expert" arrives as an
    # all-zero gradient, which must behave like the reference's grad None (common/utils.py + torch.optim.Adam)
    routed = len(getattr(model, "submodules", ())) > 1
    return FusedAdam(groups, lr=base_lr, weight_decay=weight_decay, adamw=opt_name == "adamw", skip_zero_grads=routed)
