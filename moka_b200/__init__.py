"""moka_b200: Blackwell-native MokA adapter layer (MokaLinear) behind the PEFT injection surface.

The compute lives in ``libmoka.so`` (hand-written sm_100a CUDA, C ABI in include/moka.h);
this package is the Python host that mirrors the reference's layer / injection interface."""
from .config import LLAMA_TARGETS, MokaConfig, parse_avt_rank
from ._cabi import MokaError, launch_count

__all__ = ['MokaConfig', 'LLAMA_TARGETS', 'parse_avt_rank', 'MokaError', 'launch_count',
           'MokaLinearAVT', 'MokaLinearVT', 'get_peft_model', 'PeftMixedModel',
           'build_token_meta_avt', 'build_token_meta_vt', 'TokenMeta']


def __getattr__(name):
    # torch-dependent symbols are imported lazily so `import moka_b200` stays cheap
    if name in ('MokaLinearAVT', 'MokaLinearVT', 'moka_linear_function', 'FusedScalars'):
        from . import layer
        return getattr(layer, name)
    if name in ('get_peft_model', 'PeftMixedModel', 'MokaPeftModel', 'mark_only_lora_as_trainable',
                'get_peft_model_state_dict', 'set_peft_model_state_dict'):
        from . import inject
        return getattr(inject, name)
    if name in ('build_token_meta_avt', 'build_token_meta_vt', 'build_token_meta', 'TokenMeta', 'META_CACHE'):
        from . import metadata
        return getattr(metadata, name)
    raise AttributeError(name)
