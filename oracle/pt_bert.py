"""TEST INFRASTRUCTURE -- CPU restatement of `pytorch_transformers.modeling_bert`.

The reference imports five names from the un-vendored third-party package
`pytorch_transformers` (reference pythia/models/t2s.py:9-12, m4c.py:8-11):
BertConfig, BertEmbeddings, BertEncoder, BertLayerNorm, BertPreTrainedModel.
The package is absent from /root/reference and its version is not pinned by the
reference (requirement.txt lists none); upstream M4C pinned
pytorch-transformers==1.2.0, whose published algorithm is restated here:

  * BertLayerNorm == torch.nn.LayerNorm (1.2.0 aliases it; class-default eps
    1e-5, BERT-internal LNs pass config.layer_norm_eps = 1e-12)
  * BertEmbeddings: word(pad 0) + position(arange) + token_type(zeros) -> LN
    -> dropout
  * BertSelfAttention: q,k,v Linear -> [B,12,L,64] -> QK^T/sqrt(64) + additive
    mask -> softmax -> dropout -> .V -> merge heads
  * BertSelfOutput / BertOutput: Linear -> dropout -> LN(x + residual)
  * BertIntermediate: Linear -> x*0.5*(1+erf(x/sqrt(2)))
  * BertEncoder.forward(h, mask, head_mask) -> (h,)
  * init: Linear/Embedding N(0, initializer_range), LN weight 1 bias 0,
    Linear bias 0.

"parity unpinned" for this dependency: the reference holds no golden vectors
for it.  tests/test_oracle_cpu.py cross-checks one layer of this restatement
against `transformers` (the renamed descendant of the same code) when that
package is importable.

Used (a) as the import shim that lets the real reference run in the dev
container (tests/golden/make_golden.py) and (b) by nothing in the product.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this package.
"""
import math

import torch
from torch import nn

BertLayerNorm = nn.LayerNorm


class BertConfig:
    """Defaults of BERT-base as in pytorch_transformers 1.2.0; kwargs override
    (call sites: reference t2s.py:25,27,28,46)."""

    def __init__(self, **kwargs):
        self.vocab_size = 30522
        self.hidden_size = 768
        self.num_hidden_layers = 12
        self.num_attention_heads = 12
        self.intermediate_size = 3072
        self.hidden_act = "gelu"
        self.hidden_dropout_prob = 0.1
        self.attention_probs_dropout_prob = 0.1
        self.max_position_embeddings = 512
        self.type_vocab_size = 2
        self.initializer_range = 0.02
        self.layer_norm_eps = 1e-12
        self.output_attentions = False
        self.output_hidden_states = False
        for k, v in kwargs.items():
            setattr(self, k, v)


def gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


class BertEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=0)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, input_ids, token_type_ids=None, position_ids=None):
        seq_length = input_ids.size(1)
        if position_ids is None:
            position_ids = torch.arange(seq_length, dtype=torch.long, device=input_ids.device)
            position_ids = position_ids.unsqueeze(0).expand_as(input_ids)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        e = (self.word_embeddings(input_ids) + self.position_embeddings(position_ids)
             + self.token_type_embeddings(token_type_ids))
        return self.dropout(self.LayerNorm(e))


class BertSelfAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = config.hidden_size // config.num_attention_heads
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        self.key = nn.Linear(config.hidden_size, self.all_head_size)
        self.value = nn.Linear(config.hidden_size, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)

    def _split(self, x):
        b, l, _ = x.shape
        return x.view(b, l, self.num_attention_heads, self.attention_head_size).permute(0, 2, 1, 3)

    def forward(self, hidden_states, attention_mask=None, head_mask=None):
        q = self._split(self.query(hidden_states))
        k = self._split(self.key(hidden_states))
        v = self._split(self.value(hidden_states))
        scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(self.attention_head_size)
        if attention_mask is not None:
            scores = scores + attention_mask
        probs = self.dropout(nn.functional.softmax(scores, dim=-1))
        if head_mask is not None:
            probs = probs * head_mask
        ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous()
        return (ctx.view(ctx.shape[0], ctx.shape[1], self.all_head_size),)


class BertSelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return self.LayerNorm(self.dropout(self.dense(hidden_states)) + input_tensor)


class BertAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self = BertSelfAttention(config)
        self.output = BertSelfOutput(config)

    def forward(self, input_tensor, attention_mask=None, head_mask=None):
        s = self.self(input_tensor, attention_mask, head_mask)
        return (self.output(s[0], input_tensor),)


class BertIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)

    def forward(self, hidden_states):
        return gelu(self.dense(hidden_states))


class BertOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return self.LayerNorm(self.dropout(self.dense(hidden_states)) + input_tensor)


class BertLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = BertAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)

    def forward(self, hidden_states, attention_mask=None, head_mask=None):
        a = self.attention(hidden_states, attention_mask, head_mask)[0]
        return (self.output(self.intermediate(a), a),)


class BertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(config) for _ in range(config.num_hidden_layers)])

    def forward(self, hidden_states, attention_mask=None, head_mask=None):
        for i, layer in enumerate(self.layer):
            hm = head_mask[i] if head_mask is not None else None
            hidden_states = layer(hidden_states, attention_mask, hm)[0]
        return (hidden_states,)


class BertPreTrainedModel(nn.Module):
    def __init__(self, config, *inputs, **kwargs):
        super().__init__()
        self.config = config

    def _init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, BertLayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def init_weights(self):
        self.apply(self._init_weights)

    @classmethod
    def from_pretrained(cls, *args, **kwargs):
        raise RuntimeError("no pretrained weights offline; set text_bert_init_from_bert_base=false")
