"""MokaConfig: the union of the two reference LoraConfig field sets that reach the layer.

AVT fields: AudioVisualText/peft_hyper/tuners/lora.py:31-90 (r as digit string, lora_nums,
blc_alpha, blc_weight, loramethod, reserved_modality).  VT fields:
VisualText/modified_peft/tuners/lora/config.py:43-311 (r, lora_alpha, lora_dropout,
attn_weight).  New: explicit ``num_modalities`` + integer ``r`` so ranks >= 10 are
expressible for the 3-modality variant (the digit encoding cannot: SURVEY.md fact 3)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Union

LLAMA_TARGETS = ['q_proj', 'k_proj', 'v_proj', 'o_proj', 'gate_proj', 'down_proj', 'up_proj']


def parse_avt_rank(r: int, lora_nums: int) -> List[int]:
    """The reference encodes per-modality ranks as decimal digits of one int
    (lora.py:256-259, 309-319): r=444, lora_nums=3 -> [4,4,4]."""
    digits = [int(c) for c in str(r)]
    if len(digits) < lora_nums:
        raise IndexError('list index out of range')           # same failure as lora.py:319
    return digits[:lora_nums]


@dataclass
class MokaConfig:
    variant: str = 'vt'                       # 'avt' (3 modalities, peft_hyper) | 'vt' (2 modalities, modified_peft)
    r: int = 4                                # rank per modality
    num_modalities: Optional[int] = None      # default: 3 for avt, 2 for vt
    lora_alpha: float = 16
    lora_dropout: float = 0.0
    target_modules: Optional[Union[List[str], str]] = field(default_factory=lambda: list(LLAMA_TARGETS))
    # AVT knobs (lora.py:59-66)
    blc_weight: float = 1.0
    blc_alpha: float = 1.0                    # stored, never read by the forward (lora.py:301)
    loramethod: str = 'train'
    reserved_modality: Optional[str] = None   # stored, never read (lora.py:304)
    # VT knob (lora/config.py:166; drivers use 0.05: VisualText/train/train.py:359)
    attn_weight: float = 0.5
    fan_in_fan_out: bool = False
    bias: str = 'none'
    adapter_names: tuple = ('image', 'text')  # VT adapter names, in injection order (train.py:559-561)

    def __post_init__(self):
        if self.variant not in ('avt', 'vt'):
            raise ValueError("variant must be 'avt' or 'vt'")
        if self.num_modalities is None:
            self.num_modalities = 3 if self.variant == 'avt' else 2
        if self.variant == 'vt' and self.num_modalities != 2:
            raise ValueError('the VT variant has exactly two adapters (text, image)')
        if not (1 <= self.r <= 256):
            raise ValueError('r must be in 1..256')

    @classmethod
    def from_avt_lora_config(cls, r: int = 444, lora_nums: int = 3, lora_alpha: float = 16,
                             lora_dropout: float = 0.05, blc_weight: float = 1.0, blc_alpha: float = 1.0,
                             loramethod: str = 'train', target_modules=None, **_ignored) -> 'MokaConfig':
        """Accepts the reference's AVT LoraConfig arguments (finetune.py:87-99), digit-encoded r included."""
        ranks = parse_avt_rank(r, lora_nums)
        if len(set(ranks)) != 1:
            raise ValueError(f'per-modality ranks must be equal (got {ranks}); the reference sum at lora.py:525 needs it too')
        return cls(variant='avt', r=ranks[0], num_modalities=lora_nums, lora_alpha=lora_alpha,
                   lora_dropout=lora_dropout, blc_weight=blc_weight, blc_alpha=blc_alpha,
                   loramethod=loramethod, target_modules=target_modules or list(LLAMA_TARGETS))

    @classmethod
    def from_vt_lora_config(cls, r: int = 4, lora_alpha: float = 16, lora_dropout: float = 0.05,
                            attn_weight: float = 0.5, target_modules=None, **_ignored) -> 'MokaConfig':
        """Accepts the reference's VT LoraConfig arguments (train.py:545-556)."""
        return cls(variant='vt', r=r, lora_alpha=lora_alpha, lora_dropout=lora_dropout,
                   attn_weight=attn_weight, target_modules=target_modules or list(LLAMA_TARGETS))
