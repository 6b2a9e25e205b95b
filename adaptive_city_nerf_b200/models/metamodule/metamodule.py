"""Functional-weights plumbing (reference: models/metamodule/metamodule.py:14-192).

Every MetaModule layer takes an optional `params=` OrderedDict of fast weights keyed RELATIVE to
the module it is handed to; missing keys fall back to the module's own nn.Parameters (with a
warning when a whole submodule is missing).  This is the contract MAML/FOMAML inner loops rely
on (pipelines/offline_stage/meta_core.py:27,61-64) and the fused expert kernels honour it by
taking weight POINTERS: own parameters and fast weights go down the same path."""
from __future__ import annotations

import warnings
from collections import OrderedDict
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..trunc_exp import trunc_exp


class MetaModule(nn.Module):
    """nn.Module whose parameters can be overridden per call through `params=`."""

    def __init__(self):
        super().__init__()
        self._subdict_names: Dict[tuple, tuple] = {}

    def meta_named_parameters(self, prefix: str = "", recurse: bool = True):
        """(name, parameter) pairs owned by MetaModule instances only (e.g. never the hash table,
        whose encoder is a plain nn.Module) -- reference metamodule.py:20-30."""
        seen = set()
        mods = self.named_modules(prefix=prefix) if recurse else [(prefix, self)]
        for mod_name, mod in mods:
            if not isinstance(mod, MetaModule):
                continue
            for pname, p in mod._parameters.items():
                if p is None or id(p) in seen:
                    continue
                seen.add(id(p))
                yield (f"{mod_name}.{pname}" if mod_name else pname), p

    def meta_parameters(self, recurse: bool = True):
        for _, p in self.meta_named_parameters(recurse=recurse):
            yield p

    def get_subdict(self, params: Optional[Dict[str, torch.Tensor]], key: Optional[str] = None):
        """Entries of `params` under the child `key`, with the `key.` prefix stripped
        (reference metamodule.py:37-69; same warning + None when nothing matches)."""
        if params is None:
            return None
        names = tuple(params.keys())
        cache_key = (key, names)
        sub = self._subdict_names.get(cache_key)
        if sub is None:
            if key is None:
                sub = names
            else:
                pre = key + "."
                sub = tuple(n[len(pre):] for n in names if n.startswith(pre))
            self._subdict_names[cache_key] = sub
        if not sub:
            warnings.warn(
                f"Module `{self.__class__.__name__}` has no parameter for submodule `{key}` in `params`.\n"
                f"Using default parameters. Provided keys: [{', '.join(names)}]", stacklevel=2)
            return None
        if key is None:
            return OrderedDict((n, params[n]) for n in sub)
        return OrderedDict((n, params[f"{key}.{n}"]) for n in sub)


class MetaSequential(nn.Sequential, MetaModule):
    """nn.Sequential that hands each MetaModule child its slice of `params`."""

    def forward(self, input, params: Optional[Dict[str, torch.Tensor]] = None):
        for name, module in self._modules.items():
            if isinstance(module, MetaModule):
                input = module(input, params=self.get_subdict(params, name))
            elif isinstance(module, nn.Module):
                input = module(input)
            else:
                raise TypeError(f"The module must be a `nn.Module` or `MetaModule`. Got: {type(module)}")
        return input


class MetaLinear(nn.Linear, MetaModule):
    """y = x W^T + b with optional fast weights {"weight", "bias"} (reference :129-156)."""

    def forward(self, inputs: torch.Tensor, params: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        weight = self.weight if params is None else params.get("weight", self.weight)
        bias = self.bias if params is None else params.get("bias", self.bias)
        if inputs.dtype != weight.dtype:
            inputs = inputs.to(weight.dtype)
        out = inputs.matmul(weight.t())
        return out if bias is None else out + bias


class MetaBatchLinear(nn.Linear, MetaModule):
    """Per-task batched linear: inputs (B,N,in), weight (B,out,in), bias (B,out) (reference :88-126)."""

    def forward(self, inputs: torch.Tensor, params: Optional[Dict[str, torch.Tensor]] = None):
        B = inputs.size(0)
        if params is None:
            weight = self.weight.unsqueeze(0).expand(B, -1, -1)
            bias = None if self.bias is None else self.bias.unsqueeze(0).expand(B, -1)
        else:
            weight, bias = params["weight"], params.get("bias", None)
        if weight.dim() == 2:
            weight = weight.unsqueeze(0)
        if bias is not None:
            if bias.dim() == 1:
                bias = bias.unsqueeze(0)
            elif bias.dim() == 3 and bias.shape[1] == 1:
                bias = bias.squeeze(1)
        out = torch.bmm(inputs, weight.transpose(1, 2).expand(B, -1, -1))
        return out if bias is None else out + bias.unsqueeze(1)


_ACTS = {"relu": nn.ReLU, "sigmoid": nn.Sigmoid, "softplus": nn.Softplus}


class MetaLayerBlock(MetaModule):
    """Linear + activation (reference :159-192); child names `linear` / `act` are part of the
    state_dict contract (e.g. `sigma_trunk.0.linear.weight`)."""

    def __init__(self, dim_in: int, dim_out: int, activation: Optional[str] = None, batched: bool = False):
        super().__init__()
        self.linear = MetaBatchLinear(dim_in, dim_out) if batched else MetaLinear(dim_in, dim_out)
        name = None if activation is None else activation.lower()
        if name is None:
            self.act = nn.Identity()
        elif name in _ACTS:
            self.act = _ACTS[name]()
        elif name == "trunc_exp":
            self.act = trunc_exp
        else:
            raise ValueError(f"Unsupported activation: {activation}")

    def forward(self, x: torch.Tensor, params: Optional[OrderedDict] = None) -> torch.Tensor:
        return self.act(self.linear(x, params=self.get_subdict(params, "linear")))
