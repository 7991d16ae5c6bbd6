"""A tiny Llama-shaped module tree (names and nesting of the modules the MokA injection touches:
``model.layers.N.self_attn.{q,k,v,o}_proj``, ``model.layers.N.mlp.{gate,up,down}_proj``) shared by
``oracle/gen_ckpt_keys.py`` (wrapped by the REFERENCE's PEFT forks) and ``tests/test_ckpt_keys.py`` (wrapped by
``moka_b200.inject``).  Test infrastructure."""
import torch.nn as nn

TARGETS = ['q_proj', 'k_proj', 'v_proj', 'o_proj', 'gate_proj', 'up_proj', 'down_proj']
HIDDEN, INTER, VOCAB, LAYERS = 16, 24, 32, 2


class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q_proj = nn.Linear(d, d, bias=False)
        self.k_proj = nn.Linear(d, d, bias=False)
        self.v_proj = nn.Linear(d, d, bias=False)
        self.o_proj = nn.Linear(d, d, bias=False)


class _MLP(nn.Module):
    def __init__(self, d, i):
        super().__init__()
        self.gate_proj = nn.Linear(d, i, bias=False)
        self.up_proj = nn.Linear(d, i, bias=False)
        self.down_proj = nn.Linear(i, d, bias=False)


class _Layer(nn.Module):
    def __init__(self, d, i):
        super().__init__()
        self.self_attn = _Attn(d)
        self.mlp = _MLP(d, i)
        self.input_layernorm = nn.LayerNorm(d)


class _Inner(nn.Module):
    def __init__(self):
        super().__init__()
        self.embed_tokens = nn.Embedding(VOCAB, HIDDEN)
        self.layers = nn.ModuleList([_Layer(HIDDEN, INTER) for _ in range(LAYERS)])
        self.norm = nn.LayerNorm(HIDDEN)
        # AVT trains two projectors next to the adapters (save_modules 'vl_projector,al_projector,lora')
        self.vl_projector = nn.Linear(8, HIDDEN)
        self.al_projector = nn.Linear(8, HIDDEN)


class _Cfg(dict):
    model_type = 'llama'
    tie_word_embeddings = False

    def to_dict(self):
        return {}


class TinyCausalLM(nn.Module):
    def __init__(self):
        super().__init__()
        self.model = _Inner()
        self.lm_head = nn.Linear(HIDDEN, VOCAB, bias=False)
        self.config = _Cfg()

    def prepare_inputs_for_generation(self, *a, **k):
        return None
