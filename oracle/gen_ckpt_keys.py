"""Pin the checkpoint wire format to the reference: wrap ``oracle/tiny_tree.TinyCausalLM`` with the
UNMODIFIED reference PEFT forks (build container only) and record, per variant, the ``state_dict()`` key names, shapes,
which parameters train, and the key sets of the files the reference's drivers write:

  AVT  ``adapter_model.bin``          = get_peft_state_maybe_zero_3(named_parameters, 'none') through
                                         PeftModel.save_pretrained   (scripts/finetune/finetune.py:195-200, trainer.py:60-82)
       ``non_lora_trainables.bin``     = get_peft_state_non_lora_maybe_zero_3(named_parameters)  (finetune.py:196,201; trainer.py:84-89)
  VT   ``model.safetensors``           = the whole wrapped model's state dict, loaded with strict=True
                                         (VisualText/eval_benchmarks/mmbench/mmbench.py:576-577)

Output: tests/golden/ckpt_keys.json.  Run from the repo root:  python oracle/gen_ckpt_keys.py
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = '/root/reference'

from oracle import ref_loader as rl  # noqa: E402
from oracle.tiny_tree import TARGETS, TinyCausalLM  # noqa: E402


def describe(model):
    sd = model.state_dict()
    train = {n for n, p in model.named_parameters() if p.requires_grad}
    return [[k, list(v.shape), k in train] for k, v in sd.items()]


def main():
    rl.stub_accelerate()
    out = {}
    # ---- AVT: peft_hyper
    sys.path.insert(0, os.path.join(REF, 'AudioVisualText'))
    import peft_hyper
    cfg = peft_hyper.LoraConfig(r=444, lora_alpha=16, target_modules=TARGETS, lora_dropout=0.05, bias='none',
                                task_type='CAUSAL_LM', lora_nums=3, blc_alpha=1, blc_weight=1.0)
    m = peft_hyper.get_peft_model(TinyCausalLM(), cfg)
    # finetune.py:154-160: requires_grad = name contains a key of save_modules
    for n, p in m.named_parameters():
        p.requires_grad_(any(k in n for k in ('vl_projector', 'al_projector', 'lora')))
    lora = {k: t for k, t in m.named_parameters() if 'lora_' in k}                      # trainer.py:62
    adapter = peft_hyper.get_peft_model_state_dict(m, lora)                              # peft_model.py save_pretrained
    non_lora = {k: t for k, t in m.named_parameters() if 'lora_' not in k and t.requires_grad}   # trainer.py:85-87
    out['avt'] = dict(state_dict=describe(m), adapter_model_bin=sorted(adapter), non_lora_trainables_bin=sorted(non_lora))
    # ---- VT: modified_peft (imports itself as `peft`)
    tmp = tempfile.mkdtemp(prefix='ckpt_vt_')
    os.symlink(os.path.join(REF, 'VisualText', 'modified_peft'), os.path.join(tmp, 'peft'))
    sys.path.insert(0, tmp)
    from peft import LoraConfig, PeftMixedModel
    cfg = LoraConfig(r=4, lora_alpha=16, target_modules=TARGETS, lora_dropout=0.05, task_type='CAUSAL_LM',
                     attn_weight=0.05, inference_mode=False)
    m = PeftMixedModel(TinyCausalLM(), cfg, adapter_name='image')                        # train.py:559-561
    m.add_adapter('text', cfg)
    m.set_adapter(['image', 'text'])
    out['vt'] = dict(state_dict=describe(m), model_safetensors=sorted(m.state_dict()))
    path = os.path.join(ROOT, 'tests', 'golden', 'ckpt_keys.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print('wrote', path, {k: len(v['state_dict']) for k, v in out.items()})


if __name__ == '__main__':
    main()
