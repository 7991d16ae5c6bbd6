"""PEFT-compatible injection surface for MokaLinear (L1 of SURVEY.md).

Mirrors, for the MokA path only:
* AVT: ``get_peft_model(model, LoraConfig)`` -> ``PeftModelForCausalLM`` -> ``LoraModel._find_and_replace``
  (AudioVisualText/peft_hyper/mapping.py:129-151, peft_hyper/tuners/lora.py:117-188, 230-245)
* VT: ``PeftMixedModel(model, cfg, 'image'); add_adapter('text', cfg); set_adapter([...])``
  (VisualText/modified_peft/mixed_model.py:118-132, 230-285; tuners/lora/model.py:173-261;
  tuners/tuners_utils.py:314-395)
Parameter names follow the reference so its checkpoints load:
  AVT ``base_model.model.<path>.q_proj.{weight, lora_A0.weight, ..., lora_B0.weight}``
  VT  ``base_model.model.<path>.q_proj.{base_layer.weight, lora_A.text.weight, lora_B.text.weight, ...}``
"""
from __future__ import annotations

import os
import re
from typing import Dict, Iterable, List, Optional

import torch
import torch.nn as nn

from .config import MokaConfig
from .layer import MokaLinearAVT, MokaLinearVT


def _target_found(cfg: MokaConfig, key: str) -> bool:
    t = cfg.target_modules
    if isinstance(t, str):
        return re.fullmatch(t, key) is not None
    if cfg.variant == 'avt':
        return any(key.endswith(k) for k in t)                       # lora.py:150
    return any(key == k or key.endswith('.' + k) for k in t)         # tuners_utils.py:_check_target_module_exists


def _submodules(model: nn.Module, key: str):
    parent = model.get_submodule('.'.join(key.split('.')[:-1]))
    return parent, model.get_submodule(key), key.split('.')[-1]


def mark_only_lora_as_trainable(model: nn.Module, bias: str = 'none') -> None:
    """lora.py:230-245: everything without ``lora_`` in its name is frozen."""
    for n, p in model.named_parameters():
        if 'lora_' not in n:
            p.requires_grad = False
    if bias == 'all':
        for n, p in model.named_parameters():
            if 'bias' in n:
                p.requires_grad = True
    elif bias != 'none':
        raise NotImplementedError(bias)


def inject_avt(model: nn.Module, cfg: MokaConfig, kernel_flags: int = 0) -> int:
    """lora.py:125-188: swap every nn.Linear whose name ends with a target key, SHARE the frozen
    weight/bias with the old module (lora.py:176-180)."""
    n = 0
    for key in [k for k, _ in model.named_modules()]:
        if not _target_found(cfg, key):
            continue
        parent, target, name = _submodules(model, key)
        if isinstance(target, MokaLinearAVT) or not isinstance(target, nn.Linear):
            continue
        new = MokaLinearAVT(target.in_features, target.out_features, bias=target.bias is not None,
                            rank=cfg.r, lora_alpha=cfg.lora_alpha, lora_nums=cfg.num_modalities,
                            blc_alpha=cfg.blc_alpha, blc_weight=cfg.blc_weight, lora_dropout=cfg.lora_dropout,
                            reserved_modality=cfg.reserved_modality, loramethod=cfg.loramethod,
                            merge_weights=False, kernel_flags=kernel_flags,
                            device='meta')
        new.weight = target.weight
        if target.bias is not None:
            new.bias = target.bias
        dev, dt = target.weight.device, target.weight.dtype
        for i in range(cfg.num_modalities):
            lin = nn.Linear(target.in_features, cfg.r, bias=False, device=dev, dtype=dt)
            nn.init.kaiming_uniform_(lin.weight, a=5 ** 0.5)
            setattr(new, f'lora_A{i}', lin)
        new.lora_B0 = nn.Linear(cfg.r, target.out_features, bias=False, device=dev, dtype=dt)
        nn.init.zeros_(new.lora_B0.weight)
        new.train(target.training)
        setattr(parent, name, new)
        n += 1
    if n == 0:
        raise ValueError(f'Target modules {cfg.target_modules} not found in the base model. '
                         f'Please check the target modules and try again.')
    return n


def inject_vt(model: nn.Module, cfg: MokaConfig, adapter_name: str, kernel_flags: int = 0) -> int:
    """lora/model.py:173-228: first adapter wraps the base layer, later adapters call update_layer."""
    n = 0
    for key in [k for k, _ in model.named_modules()]:
        if key.endswith('.base_layer') or '.lora_' in key:
            continue
        if not _target_found(cfg, key):
            continue
        parent, target, name = _submodules(model, key)
        if isinstance(target, MokaLinearVT):
            target.update_layer(adapter_name, cfg.r, lora_alpha=cfg.lora_alpha, lora_dropout=cfg.lora_dropout,
                                init_lora_weights=True)
            n += 1
        elif isinstance(target, nn.Linear):
            new = MokaLinearVT(target, adapter_name, r=cfg.r, lora_alpha=cfg.lora_alpha,
                               lora_dropout=cfg.lora_dropout, attn_weight=cfg.attn_weight,
                               kernel_flags=kernel_flags)
            new.train(target.training)
            setattr(parent, name, new)
            n += 1
    if n == 0:
        raise ValueError(f'Target modules {cfg.target_modules} not found in the base model. '
                         f'Please check the target modules and try again.')
    return n


class _Tuner(nn.Module):
    """Stands where ``LoraModel`` / ``MixedModel`` stand: ``self.model`` is the wrapped network."""

    def __init__(self, model: nn.Module):
        super().__init__()
        self.model = model

    def forward(self, *a, **k):
        return self.model(*a, **k)


class MokaPeftModel(nn.Module):
    """``PeftModelForCausalLM`` (AVT) / ``PeftMixedModel`` (VT) stand-in: ``base_model.model`` is
    the original network, forward delegates, trainable = names containing ``lora_``."""

    def __init__(self, model: nn.Module, peft_config: MokaConfig, adapter_name: Optional[str] = None,
                 kernel_flags: int = 0):
        super().__init__()
        self.peft_config = peft_config
        self.kernel_flags = kernel_flags
        self.base_model = _Tuner(model)
        self.config = getattr(model, 'config', None)
        if peft_config.variant == 'avt':
            inject_avt(model, peft_config, kernel_flags)
        else:
            inject_vt(model, peft_config, adapter_name or peft_config.adapter_names[0], kernel_flags)
        mark_only_lora_as_trainable(model, peft_config.bias)

    # ---- VT multi-adapter API (mixed_model.py:230-285)
    def add_adapter(self, adapter_name: str, peft_config: Optional[MokaConfig] = None) -> None:
        cfg = peft_config or self.peft_config
        if cfg.variant != 'vt':
            raise ValueError('add_adapter is the VT (modified_peft) surface')
        inject_vt(self.base_model.model, cfg, adapter_name, self.kernel_flags)
        mark_only_lora_as_trainable(self.base_model.model, cfg.bias)

    def set_adapter(self, adapter_name) -> None:
        for m in self.base_model.model.modules():
            if isinstance(m, MokaLinearVT):
                m.set_adapter(adapter_name)

    def forward(self, *args, **kwargs):
        return self.base_model(*args, **kwargs)

    def __getattr__(self, name: str):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.base_model.model, name)

    def moka_layers(self) -> List[nn.Module]:
        return [m for m in self.modules() if isinstance(m, (MokaLinearAVT, MokaLinearVT))]

    def get_nb_trainable_parameters(self):
        tr = sum(p.numel() for p in self.parameters() if p.requires_grad)
        return tr, sum(p.numel() for p in self.parameters())

    def print_trainable_parameters(self) -> None:
        tr, al = self.get_nb_trainable_parameters()
        print(f'trainable params: {tr:,d} || all params: {al:,d} || trainable%: {100 * tr / al:.4f}')

    # ---- checkpoint wire format
    def save_pretrained(self, save_directory: str, state_dict: Optional[Dict[str, torch.Tensor]] = None) -> None:
        """``adapter_model.bin`` = every ``lora_`` tensor (peft_hyper/utils/save_and_load.py:19-62,
        written by finetune.py:195-201)."""
        os.makedirs(save_directory, exist_ok=True)
        torch.save(get_peft_model_state_dict(self, state_dict), os.path.join(save_directory, 'adapter_model.bin'))


def get_peft_model(model: nn.Module, peft_config: MokaConfig, adapter_name: Optional[str] = None,
                   mixed: bool = False, kernel_flags: int = 0) -> MokaPeftModel:
    """mapping.py:129-151 (AVT) / mapping.py:119-149 with ``mixed=True`` (VT)."""
    return MokaPeftModel(model, peft_config, adapter_name, kernel_flags)


def PeftMixedModel(model: nn.Module, peft_config: MokaConfig, adapter_name: str = 'default',
                   kernel_flags: int = 0) -> MokaPeftModel:
    """VT entry point used by VisualText/train/train.py:559."""
    return MokaPeftModel(model, peft_config, adapter_name, kernel_flags)


def get_peft_model_state_dict(model: nn.Module, state_dict: Optional[Dict[str, torch.Tensor]] = None):
    sd = model.state_dict() if state_dict is None else state_dict
    return {k: v for k, v in sd.items() if 'lora_' in k}


def set_peft_model_state_dict(model: nn.Module, peft_model_state_dict: Dict[str, torch.Tensor]):
    """save_and_load.py:65-79: non-strict load of the ``lora_`` tensors."""
    return model.load_state_dict(peft_model_state_dict, strict=False)


def get_non_lora_trainables_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    """``non_lora_trainables.bin`` (AudioVisualText/scripts/finetune/finetune.py:196, 201;
    ``get_peft_state_non_lora_maybe_zero_3`` trainer.py:84-89): every TRAINABLE parameter without ``lora_`` in its
    name, keyed by ``named_parameters()`` of the wrapped model, on the CPU."""
    return {k: t.detach().cpu() for k, t in model.named_parameters() if 'lora_' not in k and t.requires_grad}


def save_reference_checkpoint(model: nn.Module, output_dir: str) -> None:
    """What finetune.py:195-201 leaves in ``output_dir``: ``adapter_model.bin`` + ``non_lora_trainables.bin``."""
    os.makedirs(output_dir, exist_ok=True)
    lora = {k: t.detach().cpu() for k, t in model.named_parameters() if 'lora_' in k}
    torch.save(get_peft_model_state_dict(model, lora), os.path.join(output_dir, 'adapter_model.bin'))
    torch.save(get_non_lora_trainables_state_dict(model), os.path.join(output_dir, 'non_lora_trainables.bin'))


def load_reference_checkpoint(model: nn.Module, ckpt_dir: str) -> Dict[str, List[str]]:
    """The loader of AudioVisualText/scripts/finetune/inference_cut.py:174-184: ``non_lora_trainables.bin`` then
    ``adapter_model.bin``, both ``strict=False``.  Returns the keys of each file that found NO parameter
    (a reference checkpoint must leave both lists empty)."""
    report = {}
    for name in ('non_lora_trainables.bin', 'adapter_model.bin'):
        path = os.path.join(ckpt_dir, name)
        if not os.path.exists(path):
            continue
        sd = torch.load(path, map_location='cpu')
        res = model.load_state_dict(sd, strict=False)
        report[name] = list(res.unexpected_keys)
    return report


def load_safetensors_checkpoint(model: nn.Module, path: str, strict: bool = True):
    """VT evaluation loader (VisualText/eval_benchmarks/mmbench/mmbench.py:576-577): the WHOLE wrapped model from
    ``model.safetensors`` with ``strict=True`` - every key of the reference's PeftMixedModel state dict must exist
    here under the same name (tests/golden/ckpt_keys.json pins them)."""
    from safetensors.torch import load_file
    if os.path.isdir(path):
        path = os.path.join(path, 'model.safetensors')
    return model.load_state_dict(load_file(path), strict=strict)


def save_safetensors_checkpoint(model: nn.Module, path: str) -> None:
    """``model.safetensors`` as HF Trainer writes it for the VT recipe (the whole state dict; shared tensors cloned)."""
    from safetensors.torch import save_file
    if os.path.isdir(path) or not path.endswith('.safetensors'):
        os.makedirs(path, exist_ok=True)
        path = os.path.join(path, 'model.safetensors')
    save_file({k: v.detach().cpu().contiguous().clone() for k, v in model.state_dict().items()}, path)


def get_mm_adapter_state(named_params: Iterable, keys_to_match: List[str]) -> Dict[str, torch.Tensor]:
    """``finetune_weights.bin`` of AudioVisualText/trainer.py:33-36, 183-218: parameters whose name
    contains any of ``save_modules`` (e.g. 'vl_projector,al_projector,lora')."""
    return {k: v.detach().cpu().clone() for k, v in named_params if any(m in k for m in keys_to_match)}


def apply_save_modules_trainable(model: nn.Module, save_modules: str) -> None:
    """finetune.py:154-160: requires_grad = name contains any of the comma-separated keys."""
    keys = save_modules.split(',')
    for n, p in model.named_parameters():
        p.requires_grad_(any(k in n for k in keys))
