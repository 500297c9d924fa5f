"""OPT-125m-shaped decoder stack (BASELINE config #3: random init, batch 8 x seq 2048 forward).

The reference reaches this graph by FX-tracing ``transformers.OPTForCausalLM`` and swapping
nodes for DmxModules; that tracer is unavailable with the installed
transformers, so -- exactly as the survey prescribes for the oracle --
the stack is assembled by hand from the module types the transformation would produce:
per layer 6 Linear, 2 ActActMatMul (QK^T, PV), 3 ResAdd (mask add + two residuals), Softmax,
2 LayerNorm, ReLU and the Dropouts, plus 2 Embeddings, the embedding ResAdd, the final
LayerNorm and the lm_head Linear.  ``mods`` selects the module family: ``dmx_compressor_b200.nn``
(cast points, configurable with config_rules.BASIC) or a plain-torch twin (the unquantised
forward that the cast overhead is measured against).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import nn as dnn


class _Add(torch.nn.Module):
    def forward(self, a, b):
        return a + b


class _MatMul(torch.nn.Module):
    def forward(self, a, b):
        return torch.matmul(a, b)


plain = SimpleNamespace(Linear=torch.nn.Linear, LayerNorm=torch.nn.LayerNorm, ReLU=torch.nn.ReLU, Softmax=torch.nn.Softmax,
                        Dropout=torch.nn.Dropout, Embedding=torch.nn.Embedding, ResAdd=_Add, ActActMatMul=_MatMul)

OPT_125M = dict(vocab_size=50272, max_position_embeddings=2048, hidden_size=768, num_hidden_layers=12, ffn_dim=3072,
                num_attention_heads=12, dropout=0.1)


class OPTLayer(torch.nn.Module):
    def __init__(self, cfg, m):
        super().__init__()
        d, h = cfg["hidden_size"], cfg["num_attention_heads"]
        self.h, self.dh = h, d // h
        self.scaling = self.dh ** -0.5
        self.self_attn_layer_norm = m.LayerNorm(d)
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = (m.Linear(d, d) for _ in range(4))
        self.qk_matmul, self.pv_matmul = m.ActActMatMul(), m.ActActMatMul()
        self.mask_add, self.attn_resadd, self.ffn_resadd = m.ResAdd(), m.ResAdd(), m.ResAdd()
        self.softmax = m.Softmax(dim=-1)
        self.attn_dropout, self.dropout1, self.dropout2 = m.Dropout(0.0), m.Dropout(cfg["dropout"]), m.Dropout(cfg["dropout"])
        self.final_layer_norm = m.LayerNorm(d)
        self.fc1, self.fc2 = m.Linear(d, cfg["ffn_dim"]), m.Linear(cfg["ffn_dim"], d)
        self.activation_fn = m.ReLU()

    def forward(self, x, mask):
        B, S, d = x.shape
        r = x
        x = self.self_attn_layer_norm(x)
        q = (self.q_proj(x) * self.scaling).view(B, S, self.h, self.dh).transpose(1, 2)
        k = self.k_proj(x).view(B, S, self.h, self.dh).transpose(1, 2)
        v = self.v_proj(x).view(B, S, self.h, self.dh).transpose(1, 2)
        w = self.qk_matmul(q, k.transpose(2, 3))
        w = self.mask_add(w, mask)
        w = self.softmax(w)
        w = self.attn_dropout(w)
        a = self.pv_matmul(w, v).transpose(1, 2).reshape(B, S, d)
        x = self.attn_resadd(r, self.dropout1(self.out_proj(a)))
        r = x
        x = self.final_layer_norm(x)
        x = self.fc2(self.activation_fn(self.fc1(x)))
        return self.ffn_resadd(r, self.dropout2(x))


class OPTStack(torch.nn.Module):
    def __init__(self, cfg=None, mods=dnn):
        super().__init__()
        cfg = dict(OPT_125M, **(cfg or {}))
        self.cfg = cfg
        d = cfg["hidden_size"]
        self.embed_tokens = mods.Embedding(cfg["vocab_size"], d, padding_idx=1)
        self.embed_positions = mods.Embedding(cfg["max_position_embeddings"] + 2, d)  # OPT's offset-2 learned positions
        self.embed_add = mods.ResAdd()
        self.layers = torch.nn.ModuleList(OPTLayer(cfg, mods) for _ in range(cfg["num_hidden_layers"]))
        self.final_layer_norm = mods.LayerNorm(d)
        self.lm_head = mods.Linear(d, cfg["vocab_size"], bias=False)

    def forward(self, input_ids):
        B, S = input_ids.shape
        pos = torch.arange(2, S + 2, device=input_ids.device).unsqueeze(0).expand(B, S)
        x = self.embed_add(self.embed_tokens(input_ids), self.embed_positions(pos))
        mask = torch.full((S, S), torch.finfo(x.dtype).min, device=x.device, dtype=x.dtype).triu(1)[None, None].expand(B, 1, S, S)
        for layer in self.layers:
            x = layer(x, mask)
        from . import elide

        return elide.materialise(self.lm_head(self.final_layer_norm(x)))  # (a deferred output cast runs here at the latest)


def build_pair(cfg=None, device="cuda", dtype=torch.float32, seed=0):
    """(dmx stack in BASIC mode, plain-torch twin) with identical random weights."""
    torch.manual_seed(seed)
    q = OPTStack(cfg, dnn)
    p = OPTStack(cfg, plain)
    p.load_state_dict({k: v for k, v in q.state_dict().items() if k in p.state_dict()}, strict=True)
    q, p = q.to(device=device, dtype=dtype).eval(), p.to(device=device, dtype=dtype).eval()
    dnn.to_basic_mode(q)
    return q, p
