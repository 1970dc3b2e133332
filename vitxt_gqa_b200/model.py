"""Registered models `t2s` and `m4c`: the reference's model API in front of libt2s_sm100.

Drop-in for reference pythia/models/t2s.py (`T2S`) and pythia/models/m4c.py (`M4C`):
same registry keys, same constructor / build() / forward(sample_list) contract,
same result-dict keys, same `state_dict` names and shapes, same
`get_optimizer_parameters`.  The module tree below only HOLDS the parameters
(nn.Linear / nn.Embedding / nn.LayerNorm objects under the reference's attribute
names); `forward` never calls them -- it enqueues the kernel schedule of
DESIGN.md on the current CUDA stream through the C ABI.  PyTorch supplies device
memory, the stream and (elsewhere) torch.distributed; there is no PyTorch or CPU
compute fallback: without the CUDA library, or with CPU inputs and no GPU, the
model raises.

Schedule (eval): fp32 grounding chain (TextBert -> obj/OCR encoders -> QTV ->
grounding) -> bf16 answer transformer with the reference's 36 full passes
collapsed to the arithmetic they actually need: per variant (ref/pos/neg) one
encoder pass over the 1044 encoder rows whose K/V are kept, the `pos` variant
decoded greedily one decoder row at a time, then `ref` and `neg` decoder rows in
one 12-row pass with the final prev_inds (encoder rows never see decoder rows and
decoder rows are causal: reference t2s.py:574-579,609-615).
"""
import math
import os

import torch
from torch import nn

from . import lib as _lib
from . import losses as _losses  # noqa: F401  (registers pos_bce_loss / InfoNCE)
from .pythia_api import BaseModel, registry

LN_EPS_BERT = 1e-12    # config.layer_norm_eps of BertConfig (BERT-internal + PrevPred LayerNorms)
LN_EPS_EMBED = 1e-5    # BertLayerNorm(hidden) class default == nn.LayerNorm default


# =============================================================================== parameter holders
class _SelfAttn(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.query, self.key, self.value = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)


class _DenseLN(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.dense = nn.Linear(i, o)
        self.LayerNorm = nn.LayerNorm(o, eps=LN_EPS_BERT)


class _Attention(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.self = _SelfAttn(h)
        self.output = _DenseLN(h, h)


class _Intermediate(nn.Module):
    def __init__(self, h, i):
        super().__init__()
        self.dense = nn.Linear(h, i)


class _BertLayer(nn.Module):
    def __init__(self, h, i):
        super().__init__()
        self.attention = _Attention(h)
        self.intermediate = _Intermediate(h, i)
        self.output = _DenseLN(i, h)


class _BertEncoder(nn.Module):
    def __init__(self, h, n_layers):
        super().__init__()
        self.layer = nn.ModuleList([_BertLayer(h, 4 * h) for _ in range(n_layers)])


class _BertEmbeddings(nn.Module):
    def __init__(self, h, vocab=30522):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, h, padding_idx=0)
        self.position_embeddings = nn.Embedding(512, h)
        self.token_type_embeddings = nn.Embedding(2, h)
        self.LayerNorm = nn.LayerNorm(h, eps=LN_EPS_BERT)


def _bert_init(module, std=0.02):
    """BertPreTrainedModel.init_weights (pytorch_transformers 1.2.0)."""
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(mean=0.0, std=std)
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)
        if isinstance(m, nn.Linear) and m.bias is not None:
            m.bias.data.zero_()


class TextBert(nn.Module):       # reference t2s.py:521-527
    def __init__(self, h, n_layers):
        super().__init__()
        self.embeddings = _BertEmbeddings(h)
        self.encoder = _BertEncoder(h, n_layers)
        _bert_init(self)


class QTV(nn.Module):            # reference t2s.py:378-382
    def __init__(self, h, n_layers):
        super().__init__()
        self.encoder = _BertEncoder(h, n_layers)
        _bert_init(self)


class _AttentionScore(nn.Module):    # reference stg.py:6-13 (linear_q / linear_k are never applied, Q5)
    def __init__(self, h):
        super().__init__()
        self.linear_q, self.linear_k = nn.Linear(h, h), nn.Linear(h, h)


class _TemporalIndicator(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.frame_pos_att, self.frame_neg_att = _AttentionScore(h), _AttentionScore(h)


class _SpatialIndicator(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.ocr_pos_att, self.ocr_neg_att = _AttentionScore(h), _AttentionScore(h)


class GroundingModule(nn.Module):    # reference t2s.py:434-451 (frame_attn / encoder are dead weights, Q18)
    def __init__(self, h, n_enc_layers):
        super().__init__()
        self.q_linear = nn.Linear(h, h)
        self.frame_attn = nn.Linear(2 * h, 1)
        self.self_attn = nn.Linear(h, 1)
        self.frame_grounding_indicator = _TemporalIndicator(h)
        self.ocr_grounding_indicator = _SpatialIndicator(h)
        self.encoder = _BertEncoder(h, n_enc_layers)


class PostHocAttention(nn.Module):   # reference m4c.py:334-346; t5vitevqa.py:334-347 adds the (unused) frame_att
    def __init__(self, h, frame_att=False):
        super().__init__()
        self.q_linear = nn.Linear(h, h)
        self.self_attn = nn.Linear(h, 1)
        if frame_att:
            self.frame_att = _AttentionScore(h)
        self.ocr_att = _AttentionScore(h)


class _PrevPredEmbeddings(nn.Module):    # reference t2s.py:673-688
    def __init__(self, h):
        super().__init__()
        self.position_embeddings = nn.Embedding(100, h)
        self.token_type_embeddings = nn.Embedding(5, h)
        self.ans_layer_norm = nn.LayerNorm(h, eps=LN_EPS_BERT)
        self.ocr_layer_norm = nn.LayerNorm(h, eps=LN_EPS_BERT)
        self.emb_layer_norm = nn.LayerNorm(h, eps=LN_EPS_BERT)


class MMT(nn.Module):            # reference t2s.py:548-554
    def __init__(self, h, n_layers):
        super().__init__()
        self.prev_pred_embeddings = _PrevPredEmbeddings(h)
        self.encoder = _BertEncoder(h, n_layers)
        _bert_init(self)


class OcrPtrNet(nn.Module):      # reference t2s.py:636-646
    def __init__(self, hidden_size, query_key_size=None):
        super().__init__()
        query_key_size = hidden_size if query_key_size is None else query_key_size
        self.query = nn.Linear(hidden_size, query_key_size)
        self.key = nn.Linear(hidden_size, query_key_size)


class ClassifierLayer(nn.Module):    # reference modules/layers.py:91-107, "linear" only
    def __init__(self, classifier_type, in_dim, out_dim, **kwargs):
        super().__init__()
        if classifier_type != "linear":
            raise NotImplementedError("Unknown classifier type: %s" % classifier_type)
        self.module = nn.Linear(in_dim, out_dim)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _dev_scalar(value, dev):
    """0-d int64 device tensor like the reference's `torch.tensor(k)` result entries (t2s.py:173-174), made by a fill
    kernel: torch.tensor(k, device=dev) is a pageable H2D copy that blocks the host until the stream drains, which
    kept the host from enqueueing the next forward while this one was still running."""
    return torch.full((), int(value), dtype=torch.int64, device=dev)


def _round_up(x, m):
    return (x + m - 1) // m * m


# =============================================================================== shared engine
class _FusionModelBase(BaseModel):
    """Everything T2S and M4C share: construction, weight packing, the kernel schedule."""

    MODEL = "t2s"

    def __init__(self, config):
        super().__init__(config)
        self._datasets = registry.get("config").datasets.split(",")
        self.hidden = int(self.config.mmt.hidden_size)
        self._packed = None
        self._packed_key = None
        self._ws = {}
        self.parity_hooks = {}     # test-only overrides, e.g. {"neg_frame_topk": tensor [B,F]}
        self.last_debug = {}
        # arithmetic of the grounding chain (TextBert, obj/OCR encoders, QTV): "bf16x3" = fp32 operands split as
        # bf16 hi|lo, three tcgen05 products per contraction (fp32-class, default); "fp32" = FMA-pipe GEMMs
        self.grounding_precision = os.environ.get(
            "T2S_B200_GROUNDING", str(self.config.get("b200_grounding_precision", "bf16x3")))
        # attention of the encoder rows: "tc" = tcgen05/TMEM kernel (default), "mma" = mma.sync kernels
        self.attn_impl = os.environ.get("T2S_B200_ATTN", str(self.config.get("b200_attention", "tc")))
        # eval only: run the latency-bound greedy decode of the `pos` variant on a second (high-priority) stream
        # while the throughput-bound encoder passes of `ref` / `neg` run on the caller's stream with their
        # persistent GEMM grids capped at (SMs - overlap_sms) CTAs.  0 disables the overlap.
        # Measured with the pair-form GEMMs (tools/overlap_sweep.sh, one box): 24 -> 21.8 ms per step, 48 -> 21.5, 64 -> 22.4;
        # the pipelined serving path (submit), whose side stream also competes with the NEXT batch's front, prefers 24
        # (3.0 k against 2.8 k samples/s), hence its own knob.
        self.overlap_sms = int(os.environ.get("T2S_B200_OVERLAP_SMS", str(self.config.get("b200_overlap_sms", 48))))
        self.submit_overlap_sms = int(os.environ.get("T2S_B200_SUBMIT_OVERLAP_SMS",
                                                     str(self.config.get("b200_submit_overlap_sms", 24))))
        self._side_streams = {}
        # pipelined eval (submit / PendingForward.result): two workspace sets used alternately; the event that ends
        # the decode tail of the forward that last used a set gates its reuse
        self._pipe_slot = 0
        self._slot_done = [None, None]
        # eval only: the 12 greedy steps are ~310 launches of 3-10 us kernels; enqueued one by one through ctypes
        # (~15 us of host time each, tensor-map encodes included) the chain is HOST-bound (4.7 ms alone, 8.7 ms next to
        # the encoder launches).  So the chain is captured once per (workspace, weights) into a CUDA graph and replayed.
        self.greedy_graph = os.environ.get("T2S_B200_GREEDY_GRAPH", str(self.config.get("b200_greedy_graph", 1))) not in ("0", "False")
        self._greedy_graphs = {}
        self._greedy_warm = set()
        self._pack_gen = 0
        self._phases_on = os.environ.get("T2S_B200_PHASES", "0") == "1"
        self._phase_events = []
        if self.attn_impl not in ("tc", "mma"):
            raise ValueError("b200_attention must be 'tc' or 'mma'")
        if self.grounding_precision not in ("bf16x3", "fp32"):
            raise ValueError("b200_grounding_precision must be 'bf16x3' or 'fp32'")

    # ---------------------------------------------------------------- build (reference t2s.py:31-151)
    def build(self):
        cfg, h = self.config, self.hidden
        if h != 768:
            raise NotImplementedError("the B200 path is built for hidden_size 768 (12 heads x 64)")
        self.finetune_modules = []
        self.text_bert = TextBert(h, int(cfg.text_bert.num_hidden_layers))
        if cfg.text_bert_init_from_bert_base:
            self._load_bert_base()
            self.finetune_modules.append({"module": self.text_bert, "lr_scale": cfg.lr_scale_text_bert})
        else:
            self.writer.write("NOT initializing text_bert from BERT_BASE")
        self.text_bert_out_linear = nn.Identity()
        self.frame_embeddings = nn.Embedding(4000, 50)
        self.linear_obj_feat_to_mmt_in = nn.Linear(int(cfg.obj.mmt_in_dim), h)
        self.obj_feat_layer_norm = nn.LayerNorm(h, eps=LN_EPS_EMBED)
        self.obj_frame_layer_norm = nn.LayerNorm(h, eps=LN_EPS_EMBED)
        self.linear_obj_frame_to_mmt_in = nn.Linear(50, h)
        self.linear_ocr_feat_to_mmt_in = nn.Linear(int(cfg.ocr.mmt_in_dim), h)
        self.linear_ocr_bbox_to_mmt_in = nn.Linear(4, h)
        self.temporal_position_embeddings = nn.Embedding(4000, 50)
        self.track_position_embeddings = nn.Embedding(4000, 50)
        self.ocr_feat_layer_norm = nn.LayerNorm(h, eps=LN_EPS_EMBED)
        self.ocr_bbox_layer_norm = nn.LayerNorm(h, eps=LN_EPS_EMBED)
        self._build_grounding()
        self.mmt = MMT(h, int(cfg.mmt.num_hidden_layers))
        self.finetune_modules.append({"module": self.mmt, "lr_scale": cfg.lr_scale_mmt})
        self.ocr_ptr_net = OcrPtrNet(**cfg.classifier.ocr_ptr_net)
        num_choices = registry.get(self._datasets[0] + "_num_final_outputs")
        num_choices -= int(cfg.classifier.ocr_max_num)
        self.classifier = ClassifierLayer(cfg["classifier"]["type"], in_dim=h, out_dim=num_choices,
                                          **cfg["classifier"]["params"])
        self.answer_processor = registry.get(self._datasets[0] + "_answer_processor")
        g = cfg.grounding
        self.frame_topk, self.ocr_topk = int(g.frame_topk), int(g.ocr_topk)
        self.frame_num, self.ocr_frame_num = int(g.frame_num), int(g.ocr_frame_num)

    def _load_bert_base(self):
        """text_bert_init_from_bert_base (reference t2s.py:47-56) reads a HuggingFace bert-base-uncased
        checkpoint from '../../huggingface/bert-base-uncased'; load it by key name when it is there."""
        import os
        path = os.path.join("..", "..", "huggingface", "bert-base-uncased", "pytorch_model.bin")
        if not os.path.exists(path):
            raise FileNotFoundError(
                "text_bert_init_from_bert_base=true needs %s (as the reference does); "
                "set text_bert_init_from_bert_base=false for random init" % path)
        sd = torch.load(path, map_location="cpu")
        own = self.text_bert.state_dict()
        for k in own:
            for cand in ("bert." + k, k):
                if cand in sd:
                    own[k].copy_(sd[cand])
                    break

    def _build_grounding(self):
        raise NotImplementedError

    # ---------------------------------------------------------------- optimizer groups (reference t2s.py:356-376)
    def get_optimizer_parameters(self, config):
        groups = []
        base_lr = config.optimizer_attributes.params.lr
        finetune = set()
        for m in self.finetune_modules:
            groups.append({"params": list(m["module"].parameters()), "lr": base_lr * m["lr_scale"]})
            finetune.update(list(m["module"].parameters()))
        groups.insert(0, {"params": [p for p in self.parameters() if p not in finetune]})
        return groups

    # ---------------------------------------------------------------- weight packing
    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    @staticmethod
    def _split_w(w, k_pad=None):
        """fp32 [N,K] -> bf16 hi|lo [N, 2*k_pad] (operand format of t2s_gemm_bf16x3), zero padded along K."""
        w = w.detach().float()
        n, k = w.shape
        kp = k if k_pad is None else k_pad
        hi = w.to(torch.bfloat16)
        lo = (w - hi.float()).to(torch.bfloat16)
        out = torch.zeros(n, 2 * kp, device=w.device, dtype=torch.bfloat16)
        out[:, :k], out[:, kp:kp + k] = hi, lo
        return out

    def _pack_layer(self, layer, mode):
        """mode: "bf16" (answer transformer), "bf16x3" (split operands) or "fp32"."""
        a = layer.attention
        wqkv = torch.cat([a.self.query.weight, a.self.key.weight, a.self.value.weight], 0).detach()
        bqkv = torch.cat([a.self.query.bias, a.self.key.bias, a.self.value.bias], 0).detach().float().contiguous()
        cast = {"bf16": lambda w: w.detach().to(torch.bfloat16).contiguous(), "bf16x3": self._split_w,
                "fp32": lambda w: w.detach().float().contiguous()}[mode]
        f32 = lambda t: t.detach().float().contiguous()
        return dict(
            wqkv=cast(wqkv), bqkv=bqkv,
            wo=cast(a.output.dense.weight), bo=f32(a.output.dense.bias),
            ln1g=f32(a.output.LayerNorm.weight), ln1b=f32(a.output.LayerNorm.bias),
            wi=cast(layer.intermediate.dense.weight), bi=f32(layer.intermediate.dense.bias),
            wo2=cast(layer.output.dense.weight), bo2=f32(layer.output.dense.bias),
            ln2g=f32(layer.output.LayerNorm.weight), ln2b=f32(layer.output.LayerNorm.bias))

    def _pack(self, device):
        key = (str(device), self.grounding_precision) + self._weights_key()
        if self._packed is not None and self._packed_key == key:
            return self._packed
        f32 = lambda t: t.detach().float().contiguous()

        def padk(w, kp):       # [N, K] -> [N, kp] zero padded, fp32
            out = torch.zeros(w.shape[0], kp, device=w.device, dtype=torch.float32)
            out[:, :w.shape[1]] = w.detach().float()
            return out

        P = {}
        gp = self.grounding_precision
        x3 = gp == "bf16x3"
        P["text"] = [self._pack_layer(l, gp) for l in self.text_bert.encoder.layer]
        if hasattr(self, "TransLayer"):
            P["qtv"] = [self._pack_layer(l, gp) for l in self.TransLayer.encoder.layer]
        P["mmt"] = [self._pack_layer(l, "bf16") for l in self.mmt.encoder.layer]
        k_obj = self.linear_obj_feat_to_mmt_in.weight.shape[1]
        k_ocr = self.linear_ocr_feat_to_mmt_in.weight.shape[1]
        P["k_obj"], P["k_ocr"] = k_obj, k_ocr
        P["k_obj_pad"], P["k_ocr_pad"] = _round_up(k_obj, 64 if x3 else 16), _round_up(k_ocr, 64 if x3 else 16)
        if x3:
            P["w_obj"] = self._split_w(self.linear_obj_feat_to_mmt_in.weight, P["k_obj_pad"])
            P["w_ocr"] = self._split_w(self.linear_ocr_feat_to_mmt_in.weight, P["k_ocr_pad"])
        else:
            P["w_obj"] = padk(self.linear_obj_feat_to_mmt_in.weight, P["k_obj_pad"])
            P["w_ocr"] = padk(self.linear_ocr_feat_to_mmt_in.weight, P["k_ocr_pad"])
        P["w_cls"] = self.classifier.module.weight.detach().to(torch.bfloat16).contiguous()
        P["w_ptr_q"] = self.ocr_ptr_net.query.weight.detach().to(torch.bfloat16).contiguous()
        P["w_ptr_k"] = self.ocr_ptr_net.key.weight.detach().to(torch.bfloat16).contiguous()
        P["f32"] = {n: f32(p) for n, p in self.named_parameters()}
        self._packed, self._packed_key = P, key
        # captured greedy-decode graphs hold raw pointers / tensor maps of the PREVIOUS packed weights: drop them, and key
        # new captures on a generation counter (id() of a freed dict can be handed to the next one)
        self._pack_gen += 1
        self._greedy_graphs.clear()
        self._greedy_warm.clear()
        return P

    # ---------------------------------------------------------------- workspaces
    def _workspace(self, B, device, dims, slot=0):
        key = (B, str(device), self.grounding_precision, slot) + tuple(sorted(dims.items()))
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        H, Lt, F, O, T, V = 768, dims["Lt"], dims["F"], dims["O"], dims["T"], dims["V"]
        Le = Lt + F + O
        Me, Md = B * Le, B * T
        f32 = dict(device=device, dtype=torch.float32)
        b16 = dict(device=device, dtype=torch.bfloat16)
        i32 = dict(device=device, dtype=torch.int32)
        n_mmt = len(self.mmt.encoder.layer)
        variants = dims["variants"]
        ws = dict(
            # fp32 grounding chain
            xt=torch.empty(B * Lt, H, **f32), xt2=torch.empty(B * Lt, H, **f32),
            h_obj=torch.empty(B * F, H, **f32), h_ocr=torch.empty(B * O, H, **f32),
            J0=torch.empty(Me, H, **f32), J1=torch.empty(Me, H, **f32),
            fx=torch.empty(Me, H, **f32), fqkv=torch.empty(Me, 3 * H, **f32), fctx=torch.empty(Me, H, **f32),
            fh=torch.empty(Me, H, **f32), fx1=torch.empty(Me, H, **f32),
            jm_ref=torch.empty(B, Le, **f32), jm_pos=torch.zeros(B, Le, **f32), jm_neg=torch.zeros(B, Le, **f32),
            jm_txt=torch.empty(B, Lt, **f32),
            keys_txt=torch.empty(B, Lt, **i32), nk_txt=torch.empty(B, **i32),
            keys={v: torch.empty(B, Le, **i32) for v in variants},
            nk={v: torch.empty(B, **i32) for v in variants},
            qp=torch.empty(B * Lt, H, **f32), gq=torch.empty(B, H, **f32), sim=torch.empty(B, F + O, **f32),
            slot=torch.empty(B, O, **f32),
            # bf16 answer transformer
            X16=torch.empty(Me, H, **b16),
            qkv0=torch.empty(Me, 3 * H, **b16),
            qkv={v: [None] + [torch.empty(Me, 3 * H, **b16) for _ in range(n_mmt - 1)] for v in variants},
            ctx=torch.empty(Me, H, **b16), hb=torch.empty(Me, H, **b16), x1=torch.empty(Me, H, **b16),
            xa=torch.empty(Me, H, **b16), xb=torch.empty(Me, H, **b16), inter=torch.empty(Me, 4 * H, **b16),
            keyp={v: torch.empty(Me, H, **b16) for v in variants},
            # decoder rows
            xd=torch.empty(Md, H, **b16), xd1=torch.empty(Md, H, **b16), xd2=torch.empty(Md, H, **b16),
            hd=torch.empty(Md, H, **b16), ctxd=torch.empty(Md, H, **b16), interd=torch.empty(Md, 4 * H, **b16),
            qd=torch.empty(Md, H, **b16), x1d=torch.empty(Md, H, **b16),
            qkvd={v: [torch.empty(Md, 3 * H, **b16) for _ in range(n_mmt)] for v in variants},
            # teacher-forced decoder rows of several variants in one pass (_decode_rows_multi): [variant][b][t] rows
            md={k: torch.empty(len(variants) * Md, w * H, **b16)
                for k, w in (("x", 1), ("x1", 1), ("x2", 1), ("h", 1), ("ctx", 1), ("inter", 4), ("q", 1), ("xa", 1))},
            md_qkv=[torch.empty(len(variants) * Md, 3 * H, **b16) for _ in range(n_mmt)],
            prev=torch.zeros(B, T, device=device, dtype=torch.int64),
            loss_ws=torch.empty(int(_lib.get_lib().loss_workspace_bytes(B, T)), device=device, dtype=torch.uint8),
        )
        if self.grounding_precision == "bf16x3":    # bf16 hi|lo operand buffers of t2s_gemm_bf16x3
            ws.update(xs=torch.empty(Me, 2 * H, **b16), ctxs=torch.empty(Me, 2 * H, **b16),
                      qkvs=torch.empty(Me, 6 * H, **b16),
                      x1s=torch.empty(Me, 2 * H, **b16), inters=torch.empty(Me, 8 * H, **b16),
                      a_obj_s=torch.empty(B * F, 2 * dims["k_obj_pad"], **b16),
                      a_ocr_s=torch.empty(B * O, 2 * dims["k_ocr_pad"], **b16))
        else:
            ws.update(fint=torch.empty(Me, 4 * H, **f32), a_obj=torch.empty(B * F, dims["k_obj_pad"], **f32),
                      a_ocr=torch.empty(B * O, dims["k_ocr_pad"], **f32))
        self._ws[key] = ws
        return ws

    # ---------------------------------------------------------------- kernel-level building blocks
    def _layer_f32(self, L, lw, x, M, rows_L, keys, nk, key_stride, ws, st, out, tanh_base=None, out16=None,
                   remap=(0, 0, 0), first=True, feeds_next=False):
        """One post-LN BERT layer of the grounding chain over `M` rows grouped in samples of `rows_L`; fp32
        activations in and out (final LN -> `out`).  `first`: the layer input has no bf16 hi|lo copy in ws["xs"]
        yet; `feeds_next`: the final LN also writes the hi|lo copy the next layer contracts with."""
        H = 768
        B = M // rows_L
        qkv, h, x1 = ws["fqkv"], ws["fh"], ws["fx1"]
        F32, RES, GELU, SPLIT = _lib.GEMM_OUT_F32, _lib.GEMM_RES_F32, _lib.GEMM_GELU, _lib.GEMM_OUT_SPLIT
        if self.grounding_precision == "bf16x3":
            xs, ctxs, x1s, inters = ws["xs"], ws["ctxs"], ws["x1s"], ws["inters"]
            if first:
                L.split_bf16(_ptr(x), H, M, H, H, _ptr(xs), 2 * H, 0, 0, 0, st)
            qkvs = ws["qkvs"]      # q|k|v as bf16 hi|lo: [M, 2 * 3H]
            L.gemm_bf16x3(_ptr(xs), 2 * H, _ptr(lw["wqkv"]), 2 * H, _ptr(lw["bqkv"]), None, 0, _ptr(qkvs), 6 * H,
                          M, 3 * H, H, SPLIT, 0, st)
            attn = L.attn_tc if self.attn_impl == "tc" else L.attn_x3
            attn(_ptr(qkvs), 6 * H, 3 * H, B, rows_L, H, 12, _ptr(keys), _ptr(nk), key_stride, _ptr(ctxs), 2 * H, st)
            L.gemm_bf16x3(_ptr(ctxs), 2 * H, _ptr(lw["wo"]), 2 * H, _ptr(lw["bo"]), _ptr(x), H, _ptr(h), H,
                          M, H, H, F32 | RES, 0, st)
            L.add_ln_split(_ptr(h), 0, H, None, 0, 0, _ptr(lw["ln1g"]), _ptr(lw["ln1b"]), LN_EPS_BERT, M, H, None, 0,
                           _ptr(x1), H, _ptr(x1s), 2 * H, 0, 0, 0, st)
            L.gemm_bf16x3(_ptr(x1s), 2 * H, _ptr(lw["wi"]), 2 * H, _ptr(lw["bi"]), None, 0, _ptr(inters), 8 * H,
                          M, 4 * H, H, GELU | SPLIT, 0, st)
            L.gemm_bf16x3(_ptr(inters), 8 * H, _ptr(lw["wo2"]), 8 * H, _ptr(lw["bo2"]), _ptr(x1), H, _ptr(h), H,
                          M, H, 4 * H, F32 | RES, 0, st)
            if feeds_next:
                assert out16 is None and tanh_base is None and remap == (0, 0, 0)
                L.add_ln_split(_ptr(h), 0, H, None, 0, 0, _ptr(lw["ln2g"]), _ptr(lw["ln2b"]), LN_EPS_BERT, M, H,
                               None, 0, _ptr(out), H, _ptr(xs), 2 * H, 0, 0, 0, st)
                return
        else:
            ctx, inter = ws["fctx"], ws["fint"]
            L.gemm_f32(_ptr(x), H, _ptr(lw["wqkv"]), H, _ptr(lw["bqkv"]), None, 0, _ptr(qkv), 3 * H, M, 3 * H, H, 0,
                       0, 0, 0, st)
            L.attn_f32(_ptr(qkv), 3 * H, B, rows_L, H, 12, _ptr(keys), _ptr(nk), key_stride, _ptr(ctx), H, None, 0, st)
            L.gemm_f32(_ptr(ctx), H, _ptr(lw["wo"]), H, _ptr(lw["bo"]), _ptr(x), H, _ptr(h), H, M, H, H, 0, 0, 0, 0, st)
            L.add_ln(_ptr(h), 0, H, None, 0, 0, _ptr(lw["ln1g"]), _ptr(lw["ln1b"]), LN_EPS_BERT, M, H, None, 0,
                     _ptr(x1), H, None, 0, 0, 0, 0, st)
            L.gemm_f32(_ptr(x1), H, _ptr(lw["wi"]), H, _ptr(lw["bi"]), None, 0, _ptr(inter), 4 * H, M, 4 * H, H,
                       GELU, 0, 0, 0, st)
            L.gemm_f32(_ptr(inter), 4 * H, _ptr(lw["wo2"]), 4 * H, _ptr(lw["bo2"]), _ptr(x1), H, _ptr(h), H, M, H,
                       4 * H, 0, 0, 0, 0, st)
        L.add_ln(_ptr(h), 0, H, None, 0, 0, _ptr(lw["ln2g"]), _ptr(lw["ln2b"]), LN_EPS_BERT, M, H,
                 _ptr(tanh_base), H, _ptr(out), H, _ptr(out16), H, remap[0], remap[1], remap[2], st)

    def _concat_linear(self, L, P, ws, which, rows, concat_args, bias, out, st):
        """normalize + id-embedding + concat (t2s.py:195-207 / 223-244) feeding the obj / OCR input projection
        (K = 1074 / 1004 zero padded) in the grounding-chain arithmetic."""
        H = 768
        kp, w = P["k_%s_pad" % which], P["w_" + which]
        if self.grounding_precision == "bf16x3":
            a_s = ws["a_%s_s" % which]      # concat written straight as bf16 hi|lo
            L.feat_concat(*concat_args, rows, None, 0, kp, _ptr(a_s), 2 * kp, st)
            L.gemm_bf16x3(_ptr(a_s), 2 * kp, _ptr(w), 2 * kp, _ptr(bias), None, 0, _ptr(out), H, rows, H, kp,
                          _lib.GEMM_OUT_F32, 0, st)
        else:
            a = ws["a_" + which]
            L.feat_concat(*concat_args, rows, _ptr(a), kp, kp, None, 0, st)
            L.gemm_f32(_ptr(a), kp, _ptr(w), kp, _ptr(bias), None, 0, _ptr(out), H, rows, H, kp, 0, 0, 0, 0, st)

    def _q_linear(self, L, P, ws, w, bias, name, B, Lt, Le, st):
        """q_proj = q_linear(txt) over the question rows of the joint buffer J1 (reference t2s.py:472, m4c.py:365)."""
        H = 768
        if self.grounding_precision == "bf16x3":
            if name not in P:
                P[name] = self._split_w(w)
            L.split_bf16(_ptr(ws["J1"]), H, B * Lt, H, H, _ptr(ws["xs"]), 2 * H, Lt, Le, 0, st)    # gathers the txt rows
            L.gemm_bf16x3(_ptr(ws["xs"]), 2 * H, _ptr(P[name]), 2 * H, _ptr(bias), None, 0, _ptr(ws["qp"]), H,
                          B * Lt, H, H, _lib.GEMM_OUT_F32, 0, st)
        else:
            L.gemm_f32(_ptr(ws["J1"]), H, _ptr(w), H, _ptr(bias), None, 0, _ptr(ws["qp"]), H, B * Lt, H, H, 0,
                       Lt, Le, 0, st)

    def _text_bert(self, L, P, ws, inp, B, Lt, Le, st):
        """TextBert (reference t2s.py:529-545); last layer's LN lands in rows [b*Le, b*Le+Lt) of J0."""
        H, f = 768, P["f32"]
        e = "text_bert.embeddings."
        L.bert_embed_ln(_ptr(inp["text"]), B * Lt, Lt, H, _ptr(f[e + "word_embeddings.weight"]),
                        _ptr(f[e + "position_embeddings.weight"]), _ptr(f[e + "token_type_embeddings.weight"]),
                        _ptr(f[e + "LayerNorm.weight"]), _ptr(f[e + "LayerNorm.bias"]), LN_EPS_BERT,
                        _ptr(ws["xt"]), H, st)
        x, y = ws["xt"], ws["xt2"]
        n = len(P["text"])
        for i, lw in enumerate(P["text"]):
            last = i == n - 1
            self._layer_f32(L, lw, x, B * Lt, Lt, ws["keys_txt"], ws["nk_txt"], Lt, ws, st,
                            out=ws["J0"] if last else y, remap=(Lt, Le, 0) if last else (0, 0, 0),
                            first=(i == 0), feeds_next=not last)
            x, y = y, x

    def _encode_obj_ocr(self, L, P, ws, inp, B, Lt, F, O, Le, st, m4c=False):
        """obj / OCR encoders (reference t2s.py:192-258; m4c.py:186-250) into rows of J0."""
        H, f = 768, P["f32"]
        n_obj = 1 if m4c else F
        vit = inp["mid_img_feat"] if m4c else inp["video_feat"]
        self._concat_linear(L, P, ws, "obj", B * n_obj,
                            (_ptr(vit), vit.shape[-1], None, 0, None if m4c else _ptr(inp["frame_id"]),
                             None if m4c else _ptr(f["frame_embeddings.weight"]), None, None, 50),
                            f["linear_obj_feat_to_mmt_in.bias"], ws["h_obj"], st)
        L.add_ln(_ptr(ws["h_obj"]), 0, H, None, 0, 0, _ptr(f["obj_feat_layer_norm.weight"]),
                 _ptr(f["obj_feat_layer_norm.bias"]), LN_EPS_EMBED, B * n_obj, H, None, 0, _ptr(ws["J0"]), H, None, 0,
                 n_obj, Le, Lt, st)
        c0 = inp["context_feature_0"]
        if "context_feature_1" in inp:
            c1 = inp["context_feature_1"]
        elif "ocr_token_bytes" in inp:
            # the OCR token TEXT came instead of its PHOC rows (vitxt_gqa_b200/featurize.py pack_tokens_fixed): build the
            # rows here, in front of the OCR encoder -- what the reference's DataLoader workers do on the CPU
            # (processors.py:904-928 over utils/phoc/src/cphoc.c), bit-identical
            rec = inp["ocr_token_bytes"]
            c1 = ws.get("phoc")
            if c1 is None or c1.shape[0] != B * O:
                c1 = ws["phoc"] = torch.empty(B * O, 604, device=rec.device, dtype=torch.float32)
            L.phoc_build_fixed(_ptr(rec), rec.shape[-1], B * O, _ptr(c1), 604, st)
        else:
            raise KeyError("the sample list carries neither `context_feature_1` nor `ocr_token_bytes`")
        self._concat_linear(L, P, ws, "ocr", B * O,
                            (_ptr(c0), c0.shape[-1], _ptr(c1), c1.shape[-1],
                             None if m4c else _ptr(inp["temporal_id"]),
                             None if m4c else _ptr(f["temporal_position_embeddings.weight"]),
                             None if m4c else _ptr(inp["track_id"]),
                             None if m4c else _ptr(f["track_position_embeddings.weight"]), 50),
                            f["linear_ocr_feat_to_mmt_in.bias"], ws["h_ocr"], st)
        L.ocr_finish(_ptr(ws["h_ocr"]), H, _ptr(inp["ocr_bbox_coordinates"]), _ptr(f["linear_ocr_bbox_to_mmt_in.weight"]),
                     _ptr(f["linear_ocr_bbox_to_mmt_in.bias"]), _ptr(f["ocr_feat_layer_norm.weight"]),
                     _ptr(f["ocr_feat_layer_norm.bias"]), _ptr(f["ocr_bbox_layer_norm.weight"]),
                     _ptr(f["ocr_bbox_layer_norm.bias"]), LN_EPS_EMBED, B * O, H, _ptr(ws["J0"]), H, O, Le,
                     Lt + n_obj, st)

    def _mmt_encoder(self, L, P, ws, variants, B, Le, st, qkv0=True, sm_cap=0):
        """Encoder rows of the answer transformer for each variant; keeps per-layer q|k|v and the
        pointer-net key projection of the last layer (reference t2s.py:622-631, 659).  `sm_cap` > 0 caps the
        persistent GEMM grids (T2S_GEMM_SM_CAP) while the decode chain runs on another stream."""
        H, M = 768, B * Le
        layers = P["mmt"]
        f = P["f32"]
        lw0 = layers[0]
        CAP = (sm_cap & 0xff) << _lib.GEMM_SM_CAP_SHIFT
        if qkv0:
            # layer-0 q|k|v only depends on the (variant-independent) input rows: compute once
            L.gemm_bf16(_ptr(ws["X16"]), H, _ptr(lw0["wqkv"]), H, _ptr(lw0["bqkv"]), None, 0, _ptr(ws["qkv0"]), 3 * H,
                        M, 3 * H, H, CAP, 0, st)
        for v in variants:
            x = ws["X16"]
            ping = [ws["xa"], ws["xb"]]
            for li, lw in enumerate(layers):
                if li == 0:
                    qkv = ws["qkv0"]
                else:
                    qkv = ws["qkv"][v][li]
                    L.gemm_bf16(_ptr(x), H, _ptr(lw["wqkv"]), H, _ptr(lw["bqkv"]), None, 0, _ptr(qkv), 3 * H,
                                M, 3 * H, H, CAP, 0, st)
                if self.attn_impl == "tc":
                    L.attn_tc(_ptr(qkv), 3 * H, 0, B, Le, H, 12, _ptr(ws["keys"][v]), _ptr(ws["nk"][v]), Le,
                              _ptr(ws["ctx"]), H, st)
                else:
                    L.attn_bf16(_ptr(qkv), 3 * H, B, Le, H, 12, _ptr(ws["keys"][v]), _ptr(ws["nk"][v]), Le,
                                _ptr(ws["ctx"]), H, st)
                L.gemm_bf16(_ptr(ws["ctx"]), H, _ptr(lw["wo"]), H, _ptr(lw["bo"]), _ptr(x), H, _ptr(ws["hb"]), H,
                            M, H, H, CAP, 0, st)
                L.add_ln(_ptr(ws["hb"]), 1, H, None, 0, 0, _ptr(lw["ln1g"]), _ptr(lw["ln1b"]), LN_EPS_BERT, M, H,
                         None, 0, None, 0, _ptr(ws["x1"]), H, 0, 0, 0, st)
                L.gemm_bf16(_ptr(ws["x1"]), H, _ptr(lw["wi"]), H, _ptr(lw["bi"]), None, 0, _ptr(ws["inter"]),
                            4 * H, M, 4 * H, H, _lib.GEMM_GELU | CAP, 0, st)
                L.gemm_bf16(_ptr(ws["inter"]), 4 * H, _ptr(lw["wo2"]), 4 * H, _ptr(lw["bo2"]), _ptr(ws["x1"]), H,
                            _ptr(ws["hb"]), H, M, H, 4 * H, CAP, 0, st)
                out = ping[li & 1]
                L.add_ln(_ptr(ws["hb"]), 1, H, None, 0, 0, _ptr(lw["ln2g"]), _ptr(lw["ln2b"]), LN_EPS_BERT, M, H,
                         None, 0, None, 0, _ptr(out), H, 0, 0, 0, st)
                x = out
            L.gemm_bf16(_ptr(x), H, _ptr(P["w_ptr_k"]), H, _ptr(f["ocr_ptr_net.key.bias"]), None, 0,
                        _ptr(ws["keyp"][v]), H, M, H, H, CAP, 0, st)

    def _decode_rows(self, L, P, ws, v, jm, scores, B, Le, T, V, O, n_obj, Lt, t0, nq, st):
        """Decoder rows t0..t0+nq-1 of variant `v` through all layers, then both score heads
        (reference t2s.py:566-568, 622-631, 279-286)."""
        H, f = 768, P["f32"]
        pp = "mmt.prev_pred_embeddings."
        ocr_row0 = Lt + n_obj
        L.prev_embed(_ptr(ws["prev"]), T, B, t0, nq, T, V, H, _ptr(f["classifier.module.weight"]),
                     ws["J1"].data_ptr() + ocr_row0 * H * 4, Le * H, H,
                     _ptr(f[pp + "position_embeddings.weight"]), _ptr(f[pp + "token_type_embeddings.weight"]),
                     _ptr(f[pp + "ans_layer_norm.weight"]), _ptr(f[pp + "ans_layer_norm.bias"]),
                     _ptr(f[pp + "ocr_layer_norm.weight"]), _ptr(f[pp + "ocr_layer_norm.bias"]),
                     _ptr(f[pp + "emb_layer_norm.weight"]), _ptr(f[pp + "emb_layer_norm.bias"]), LN_EPS_BERT,
                     _ptr(ws["xd"]), None, H, O, st)
        if nq == T:
            M, rs, off = B * T, 1, 0          # all decoder rows, contiguous
        else:
            assert nq == 1
            M, rs, off = B, T, t0             # one row per sample, T rows apart

        def at(t, width, esize=2):            # pointer to row t0 of a [B*T, width] buffer
            return t.data_ptr() + off * width * esize

        gemm = L.gemm_bf16

        x = ws["xd"]
        ping = [ws["xd1"], ws["xd2"]]
        for li, lw in enumerate(P["mmt"]):
            qkvd = ws["qkvd"][v][li]
            qkve = ws["qkv0"] if li == 0 else ws["qkv"][v][li]
            gemm(at(x, H), rs * H, _ptr(lw["wqkv"]), H, _ptr(lw["bqkv"]), None, 0, at(qkvd, 3 * H),
                        rs * 3 * H, M, 3 * H, H, 0, 0, st)
            L.attn_dec(_ptr(qkve), 3 * H, Le, _ptr(qkvd), 3 * H, T, B, H, 12, _ptr(ws["keys"][v]), _ptr(ws["nk"][v]),
                       Le, t0, nq, _ptr(ws["ctxd"]), H, st)
            gemm(at(ws["ctxd"], H), rs * H, _ptr(lw["wo"]), H, _ptr(lw["bo"]), at(x, H), rs * H,
                        at(ws["hd"], H), rs * H, M, H, H, 0, 0, st)
            L.add_ln(at(ws["hd"], H), 1, rs * H, None, 0, 0, _ptr(lw["ln1g"]), _ptr(lw["ln1b"]), LN_EPS_BERT, M, H,
                     None, 0, None, 0, at(ws["x1d"], H), rs * H, 0, 0, 0, st)
            gemm(at(ws["x1d"], H), rs * H, _ptr(lw["wi"]), H, _ptr(lw["bi"]), None, 0, at(ws["interd"], 4 * H),
                        rs * 4 * H, M, 4 * H, H, _lib.GEMM_GELU, 0, st)
            gemm(at(ws["interd"], 4 * H), rs * 4 * H, _ptr(lw["wo2"]), 4 * H, _ptr(lw["bo2"]), at(ws["x1d"], H),
                        rs * H, at(ws["hd"], H), rs * H, M, H, 4 * H, 0, 0, st)
            out = ping[li & 1]
            L.add_ln(at(ws["hd"], H), 1, rs * H, None, 0, 0, _ptr(lw["ln2g"]), _ptr(lw["ln2b"]), LN_EPS_BERT, M, H,
                     None, 0, None, 0, at(out, H), rs * H, 0, 0, 0, st)
            x = out
        N = V + O
        gemm(at(x, H), rs * H, _ptr(P["w_cls"]), H, _ptr(f["classifier.module.bias"]), None, 0,
                    scores.data_ptr() + off * N * 4, rs * N, M, V, H, _lib.GEMM_OUT_F32, 0, st)
        gemm(at(x, H), rs * H, _ptr(P["w_ptr_q"]), H, _ptr(f["ocr_ptr_net.query.bias"]), None, 0,
                    at(ws["qd"], H), rs * H, M, H, H, 0, 0, st)
        L.ptr_score(_ptr(ws["qd"]), H, B, T, t0, nq, ws["keyp"][v].data_ptr() + ocr_row0 * H * 2, Le * H, H, O, H,
                    jm.data_ptr() + ocr_row0 * 4, Le, _ptr(scores), N, V, st)

    # ---------------------------------------------------------------- training step (vitxt_gqa_b200/train.py)
    TRAIN_VARIANT = None        # single-variant models: name of the one answer-transformer variant
    TRAIN_DEAD = ("Grounding_Module.", "linear_obj_frame_to_mmt_in.", "obj_frame_layer_norm.")   # never get a gradient

    def train_engine(self):
        """The flat-buffer training engine of this model (vitxt_gqa_b200/train.py), created on first use."""
        eng = getattr(self, "_train_engine", None)
        if eng is None or eng.dev != self._device():
            from .train import TrainEngine
            eng = TrainEngine(self)
            object.__setattr__(self, "_train_engine", eng)
        return eng

    def _train_forward_single(self, inp, dev):
        """Training-mode forward of a single-variant model behind the engine's autograd Function."""
        from . import train as _train
        eng = self.train_engine()
        (scores,) = _train._T2STrainFn.apply(eng, inp, *eng.live_params)
        ground_frame, ground_box = eng.ground
        return scores, ground_frame, ground_box

    def _run_greedy(self, greedy, use_graph, gkey, pos_buf, pos_out, stream, dev):
        """Run the greedy-decode launch chain `greedy(stream_handle)` on torch's current stream `stream`: eagerly, or --
        from the third forward with the same workspace and packed weights on -- as one CUDA-graph replay (the chain
        only touches workspace / weight buffers, whose addresses are fixed; its scores land in `pos_buf` and are
        copied to this forward's `pos_out`)."""
        if not use_graph:
            greedy(stream.cuda_stream)
            return
        lib = _lib.get_lib()
        entry = self._greedy_graphs.get(gkey)
        if entry is None:
            if gkey not in self._greedy_warm:       # first forward: eager (one-time attribute calls, lazy init)
                self._greedy_warm.add(gkey)
                greedy(stream.cuda_stream)
                pos_out.copy_(pos_buf)
                return
            n0 = lib.launches
            try:
                graph = torch.cuda.CUDAGraph()
                # thread_local: other host threads (NCCL watchdog, pin-memory workers) keep using the CUDA API
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    greedy(torch.cuda.current_stream(dev).cuda_stream)
            except RuntimeError as e:       # e.g. another thread touched the CUDA API during the capture
                lib.launches = n0
                self.greedy_graph = False   # same kernels, launched one by one from now on
                self.writer.write("greedy-decode graph capture failed (%s): eager launches" % e, "warning")
                greedy(stream.cuda_stream)
                pos_out.copy_(pos_buf)
                return
            entry = self._greedy_graphs[gkey] = (graph, lib.launches - n0)
            lib.launches = n0               # recorded, not run: the replay below is what executes (and counts) them
        entry[0].replay()
        lib.launches += entry[1]            # kernels of ours inside the replayed graph (bench.py's gpu_launches)
        pos_out.copy_(pos_buf)

    def _decode_rows_multi(self, L, P, ws, vs, jm, scores, B, Le, T, V, O, n_obj, Lt, st):
        """All T teacher-forced decoder rows of the variants `vs` in ONE pass: the decoder rows share every weight,
        so the dense layers, LayerNorms and both score-head projections run over len(vs)*B*T rows at once; only the
        attention (per-variant encoder K|V and key list) and the pointer scores (per-variant OCR keys / mask) are
        issued per variant.  `scores`: one fp32 [len(vs), B, T, V+O] buffer (the per-variant results are its slices).
        Same arithmetic per row as `_decode_rows(..., 0, T)` -- the tcgen05 GEMM is row-independent."""
        H, f = 768, P["f32"]
        pp = "mmt.prev_pred_embeddings."
        ocr_row0 = Lt + n_obj
        nv, Md = len(vs), B * T
        M = nv * Md
        md, N = ws["md"], V + O
        assert scores.shape == (nv, B, T, N) and scores.is_contiguous()

        def vrow(t, i, width, esize=2):       # pointer to the first row of variant i in a [nv*Md, width] buffer
            return t.data_ptr() + i * Md * width * esize

        for i in range(nv):                   # identical input rows per variant (prev_inds are shared)
            L.prev_embed(_ptr(ws["prev"]), T, B, 0, T, T, V, H, _ptr(f["classifier.module.weight"]),
                         ws["J1"].data_ptr() + ocr_row0 * H * 4, Le * H, H,
                         _ptr(f[pp + "position_embeddings.weight"]), _ptr(f[pp + "token_type_embeddings.weight"]),
                         _ptr(f[pp + "ans_layer_norm.weight"]), _ptr(f[pp + "ans_layer_norm.bias"]),
                         _ptr(f[pp + "ocr_layer_norm.weight"]), _ptr(f[pp + "ocr_layer_norm.bias"]),
                         _ptr(f[pp + "emb_layer_norm.weight"]), _ptr(f[pp + "emb_layer_norm.bias"]), LN_EPS_BERT,
                         vrow(md["x"], i, H), None, H, O, st)
        x = md["x"]
        ping = [md["x1"], md["x2"]]
        for li, lw in enumerate(P["mmt"]):
            qkvd = ws["md_qkv"][li]
            L.gemm_bf16(_ptr(x), H, _ptr(lw["wqkv"]), H, _ptr(lw["bqkv"]), None, 0, _ptr(qkvd), 3 * H, M, 3 * H, H,
                        0, 0, st)
            for i, v in enumerate(vs):
                qkve = ws["qkv0"] if li == 0 else ws["qkv"][v][li]
                L.attn_dec(_ptr(qkve), 3 * H, Le, vrow(qkvd, i, 3 * H), 3 * H, T, B, H, 12, _ptr(ws["keys"][v]),
                           _ptr(ws["nk"][v]), Le, 0, T, vrow(md["ctx"], i, H), H, st)
            L.gemm_bf16(_ptr(md["ctx"]), H, _ptr(lw["wo"]), H, _ptr(lw["bo"]), _ptr(x), H, _ptr(md["h"]), H, M, H, H,
                        0, 0, st)
            L.add_ln(_ptr(md["h"]), 1, H, None, 0, 0, _ptr(lw["ln1g"]), _ptr(lw["ln1b"]), LN_EPS_BERT, M, H,
                     None, 0, None, 0, _ptr(md["xa"]), H, 0, 0, 0, st)
            L.gemm_bf16(_ptr(md["xa"]), H, _ptr(lw["wi"]), H, _ptr(lw["bi"]), None, 0, _ptr(md["inter"]), 4 * H,
                        M, 4 * H, H, _lib.GEMM_GELU, 0, st)
            L.gemm_bf16(_ptr(md["inter"]), 4 * H, _ptr(lw["wo2"]), 4 * H, _ptr(lw["bo2"]), _ptr(md["xa"]), H,
                        _ptr(md["h"]), H, M, H, 4 * H, 0, 0, st)
            out = ping[li & 1]
            L.add_ln(_ptr(md["h"]), 1, H, None, 0, 0, _ptr(lw["ln2g"]), _ptr(lw["ln2b"]), LN_EPS_BERT, M, H,
                     None, 0, None, 0, _ptr(out), H, 0, 0, 0, st)
            x = out
        L.gemm_bf16(_ptr(x), H, _ptr(P["w_cls"]), H, _ptr(f["classifier.module.bias"]), None, 0, _ptr(scores), N,
                    M, V, H, _lib.GEMM_OUT_F32, 0, st)
        L.gemm_bf16(_ptr(x), H, _ptr(P["w_ptr_q"]), H, _ptr(f["ocr_ptr_net.query.bias"]), None, 0, _ptr(md["q"]), H,
                    M, H, H, 0, 0, st)
        for i, v in enumerate(vs):
            L.ptr_score(vrow(md["q"], i, H), H, B, T, 0, T, ws["keyp"][v].data_ptr() + ocr_row0 * H * 2, Le * H, H, O, H,
                        jm[v].data_ptr() + ocr_row0 * 4, Le, scores[i].data_ptr(), N, V, st)

    # ---------------------------------------------------------------- input plumbing
    _I64 = ("text", "text_len", "frame_id", "frame_mask", "temporal_id", "track_id", "ocr_mask", "train_prev_inds",
            "middel_frame_id", "middel_frame_idx",
            "ocr_temporal_id", "ocr_track_id", "ocr_mask_embedding", "frame_mask_embedding", "frame_list")
    _F32 = ("video_feat", "context_feature_0", "context_feature_1", "ocr_bbox_coordinates", "mid_img_feat",
            "gumbel_frame", "gumbel_ocr", "ocr_bbox_list")

    def _device(self):
        return next(self.parameters()).device

    # measurement aid (T2S_B200_PHASES=1): CUDA events at the phase boundaries of forward, on the stream that runs
    # the phase; `phase_report()` synchronises and returns {phase: ms since the start of that forward}
    def _mark(self, name, stream=None):
        if not self._phases_on:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream if stream is not None else torch.cuda.current_stream())
        self._phase_events.append((name, ev))

    def phase_report(self):
        torch.cuda.synchronize()
        evs, self._phase_events = self._phase_events, []
        if not evs:
            return {}
        t0 = evs[0][1]
        return {n: t0.elapsed_time(e) for n, e in evs}

    _U8 = ("ocr_token_bytes",)

    def _gather_inputs(self, sample_list, names):
        dev = self._device()
        if dev.type != "cuda":
            raise _lib.T2SLibraryError(
                "the %s forward path runs only on a CUDA device (sm_100a); move the model with .to('cuda'). "
                "There is no CPU fallback." % self.MODEL)
        out = {}
        for n in tuple(names) + self._U8:
            if n not in sample_list:
                continue
            t = sample_list[n]
            want = torch.int64 if n in self._I64 else (torch.uint8 if n in self._U8 else torch.float32)
            out[n] = t.to(device=dev, dtype=want, non_blocking=True).contiguous()
        return out

    def forward(self, sample_list):
        raise NotImplementedError


class PendingForward:
    """Handle of a batch submitted with `model.submit(sample_list)`."""

    def __init__(self, out, done, device):
        self._out, self._done, self._device = out, done, device

    def ready(self):
        """True once the device has finished this batch (does not block)."""
        return self._done.query()

    def result(self, stream=None):
        """Make `stream` (default: the current stream) wait for this batch and return its report dict."""
        (stream if stream is not None else torch.cuda.current_stream(self._device)).wait_event(self._done)
        return self._out


# =============================================================================== T2S
@registry.register_model("t2s")
class T2S(_FusionModelBase):
    MODEL = "t2s"
    ABLATION = ""       # "" | "wo_sg" | "wo_tg": the Grounding_Module wiring of the reference's ablation models

    def _build_grounding(self):
        cfg, h = self.config, self.hidden
        self.TransLayer = QTV(h, int(cfg.translayers.num_hidden_layers))
        self.Grounding_Module = GroundingModule(h, int(cfg.encoder.num_hidden_layers))

    def _grounding(self, L, P, ws, inp, B, Lt, F, O, Of, Le, dev, st):
        """Grounding_Module.forward (reference t2s.py:461-518) on the joint features in ws["J1"]: question pooling,
        similarities, temporal / spatial selection, pos / neg joint masks and their key lists."""
        H, f = 768, P["f32"]
        g = "Grounding_Module."
        self._q_linear(L, P, ws, f[g + "q_linear.weight"], f[g + "q_linear.bias"], "q_linear", B, Lt, Le, st)
        L.question_pool(_ptr(ws["qp"]), B, Lt, H, _ptr(f[g + "self_attn.weight"]), _ptr(f[g + "self_attn.bias"]),
                        _ptr(ws["jm_ref"]), Le, _ptr(ws["gq"]), st)
        L.sim_scores(_ptr(ws["gq"]), _ptr(ws["J1"]), Le * H, H, Lt, F + O, H, B, _ptr(ws["sim"]), st)
        if "gumbel_frame" in inp:
            gf, go = inp["gumbel_frame"], inp["gumbel_ocr"]
        else:   # F.gumbel_softmax draws -log(Exp(1)) noise (reference stg.py:41,89); torch RNG is plumbing here
            gf = -torch.empty(B, 2, F, device=dev).exponential_().log()
            go = -torch.empty(B, 2, O, device=dev).exponential_().log()
        # test-only overrides; the device copies must stay referenced until the launch is enqueued
        pos_ovr = self.parity_hooks.get("pos_frame_topk")
        neg_ovr = self.parity_hooks.get("neg_frame_topk")
        pos_ovr = pos_ovr.to(dev).float().contiguous() if pos_ovr is not None else None
        neg_ovr = neg_ovr.to(dev).float().contiguous() if neg_ovr is not None else None
        dbg = self.parity_hooks.get("debug", False)
        dbg_f = torch.empty(B, F, device=dev) if dbg else None
        dbg_o = torch.empty(B, O, device=dev) if dbg else None
        boxes = inp["ocr_bbox_coordinates"]
        if self.ABLATION == "wo_tg":
            # reference models/t2s_wo_tg.py:483-535: no temporal stage -- the slots of every frame id compete, the
            # spatial stage keeps frame_topk * ocr_topk tokens per frame, and the frame masks follow from the OCR masks
            k_o = self.frame_topk * self.ocr_topk
            ground_frame = torch.empty(B, 5, device=dev, dtype=torch.int64)          # the 5 is literal in the reference
            ground_box = torch.empty(B, F * min(k_o, Of), 4, device=dev, dtype=torch.float32)
            L.frame_slots(_ptr(inp["frame_id"]), F, _ptr(inp["temporal_id"]), B, O, _ptr(ws["slot"]), st)
            L.spatial_select(_ptr(ws["sim"]), F + O, F, _ptr(ws["slot"]), _ptr(ws["jm_ref"]), B, Le, Lt + F, F, Of,
                             _ptr(go), _ptr(boxes), k_o, 3, _ptr(ground_box), _ptr(ws["jm_pos"]), _ptr(ws["jm_neg"]),
                             _ptr(dbg_o), st)
            L.frames_from_ocr(_ptr(ws["jm_ref"]), _ptr(ws["jm_pos"]), _ptr(ws["jm_neg"]), B, Lt, F, Of, 5,
                              _ptr(ground_frame), st)
        else:
            ground_frame = torch.empty(B, self.frame_topk, device=dev, dtype=torch.int64)
            L.temporal_select(_ptr(ws["sim"]), F + O, _ptr(ws["jm_ref"]), B, Lt, F, Of, _ptr(gf), _ptr(inp["frame_id"]),
                              _ptr(inp["temporal_id"]), self.frame_topk,
                              _ptr(pos_ovr), _ptr(neg_ovr),
                              _ptr(ground_frame), _ptr(ws["jm_pos"]), _ptr(ws["jm_neg"]), _ptr(ws["slot"]), _ptr(dbg_f), st)
            if self.ABLATION == "wo_sg":
                # reference models/t2s_wo_sg.py:496-506: all slots of the grounded frames / all the other slots
                ground_box = torch.zeros(B, self.frame_topk * Of, 4, device=dev, dtype=torch.float32)
                L.spatial_select(_ptr(ws["sim"]), F + O, F, _ptr(ws["slot"]), _ptr(ws["jm_ref"]), B, Le, Lt + F, F, Of,
                                 _ptr(go), _ptr(boxes), self.frame_topk, 2, _ptr(ground_box), _ptr(ws["jm_pos"]),
                                 _ptr(ws["jm_neg"]), _ptr(dbg_o), st)
            else:
                ground_box = torch.empty(B, F * min(self.ocr_topk, Of), 4, device=dev, dtype=torch.float32)
                L.spatial_select(_ptr(ws["sim"]), F + O, F, _ptr(ws["slot"]), _ptr(ws["jm_ref"]), B, Le, Lt + F, F, Of,
                                 _ptr(go), _ptr(boxes), self.ocr_topk, 0, _ptr(ground_box), _ptr(ws["jm_pos"]),
                                 _ptr(ws["jm_neg"]), _ptr(dbg_o), st)
        L.build_keys(_ptr(ws["jm_pos"]), B, Le, _ptr(ws["keys"]["pos"]), _ptr(ws["nk"]["pos"]), Le, st)
        L.build_keys(_ptr(ws["jm_neg"]), B, Le, _ptr(ws["keys"]["neg"]), _ptr(ws["nk"]["neg"]), Le, st)

        return ground_frame, ground_box, dbg, dbg_f, dbg_o

    def submit(self, sample_list):
        """Pipelined eval forward (serving API): enqueue this batch and return a `PendingForward` at once.

        Same kernels and results as `model(sample_list)`; the difference is the tail.  The greedy decode of `pos`, the
        decoder rows of `ref` / `neg` and the losses run on the side stream and are NOT joined into the caller's
        stream, so the front of the next submitted batch (caller's stream, GEMM grids capped) overlaps them.
        `PendingForward.result()` makes the current stream wait for the tail and returns the report dict of
        `BaseModel.__call__` (scores, grounding, "losses", "metrics").  At most two batches are in flight: the third
        submit waits (on the device) for the first one's tail before it reuses that workspace set."""
        if self.training or torch.is_grad_enabled():
            raise RuntimeError("submit() is the eval path: call model.eval() and use torch.no_grad()")
        if not 0 < self.overlap_sms:
            raise RuntimeError("submit() needs the decode side stream (b200_overlap_sms > 0)")
        return self.forward(sample_list, _pipelined=True)

    def forward(self, sample_list, _pipelined=False):
        L = _lib.get_lib()
        inp = self._gather_inputs(sample_list, self._I64 + self._F32)
        dev = self._device()
        B, Lt = inp["text"].shape
        F = inp["video_feat"].shape[1]
        O = inp["ocr_mask"].shape[1]
        T = inp["train_prev_inds"].shape[1]
        V = self.classifier.module.weight.shape[0]
        Of = O // F
        if F != self.frame_num or Of != self.ocr_frame_num or O != F * Of:
            raise ValueError("inputs have %d frames x %d OCR slots but the config says %d x %d"
                             % (F, Of, self.frame_num, self.ocr_frame_num))
        Le, H = Lt + F + O, 768
        if self.training and torch.is_grad_enabled():
            # training step: teacher-forced forward that keeps what the backward needs, behind a
            # torch.autograd.Function so that the reference trainer's loss.backward() drives the backward kernels
            from . import train as _train
            eng = self.train_engine()
            ref, pos, neg = _train._T2STrainFn.apply(eng, inp, *eng.live_params)
            ground_frame, ground_box = eng.ground
            return {
                "ref_scores": ref, "pos_scores": pos, "neg_scores": neg, "ground_box": ground_box,
                "ground_frame": ground_frame, "frame_topk": _dev_scalar(self.frame_topk, dev),
                "ocr_topk": _dev_scalar(self.ocr_topk, dev),
            }
        P = self._pack(dev)
        variants = ("pos", "ref", "neg")
        n_sms = torch.cuda.get_device_properties(dev).multi_processor_count
        slot = 0
        if _pipelined:
            slot, self._pipe_slot = self._pipe_slot, self._pipe_slot ^ 1
            if self._slot_done[slot] is not None:       # the forward that used this set two submits ago
                torch.cuda.current_stream(dev).wait_event(self._slot_done[slot])
            L.gemm_cap = max(1, n_sms - self.submit_overlap_sms)
        else:
            for i_, ev_ in enumerate(self._slot_done):     # a plain call after submits: their tails still own set 0 / 1
                if ev_ is not None:
                    torch.cuda.current_stream(dev).wait_event(ev_)
                    self._slot_done[i_] = None
        ws = self._workspace(B, dev, dict(Lt=Lt, F=F, O=O, T=T, V=V, k_obj_pad=P["k_obj_pad"],
                                          k_ocr_pad=P["k_ocr_pad"], variants=variants), slot=slot)
        st = torch.cuda.current_stream(dev).cuda_stream
        f = P["f32"]
        try:
            return self._eval_or_teacher_forced(L, P, ws, inp, sample_list, dev, st, B, Lt, F, O, Of, T, V, Le, H,
                                                variants, n_sms, slot, _pipelined)
        finally:
            L.gemm_cap = 0

    def _eval_or_teacher_forced(self, L, P, ws, inp, sample_list, dev, st, B, Lt, F, O, Of, T, V, Le, H, variants,
                                n_sms, slot, _pipelined):
        f = P["f32"]

        self._phase_events = []
        self._mark("start")
        # ---- masks and key lists
        L.mask_prep(_ptr(inp["text_len"]), _ptr(inp["frame_mask"]), _ptr(inp["ocr_mask"]), B, Lt, F, O,
                    _ptr(ws["jm_ref"]), st)
        L.mask_prep(_ptr(inp["text_len"]), None, None, B, Lt, 0, 0, _ptr(ws["jm_txt"]), st)
        L.build_keys(_ptr(ws["jm_txt"]), B, Lt, _ptr(ws["keys_txt"]), _ptr(ws["nk_txt"]), Lt, st)
        L.build_keys(_ptr(ws["jm_ref"]), B, Le, _ptr(ws["keys"]["ref"]), _ptr(ws["nk"]["ref"]), Le, st)

        # ---- fp32 grounding chain
        self._text_bert(L, P, ws, inp, B, Lt, Le, st)
        self._encode_obj_ocr(L, P, ws, inp, B, Lt, F, O, Le, st)
        x, n = ws["J0"], len(P["qtv"])
        if n > 2 and "fx2" not in ws:
            ws["fx2"] = torch.empty_like(ws["fx"])
        for i, lw in enumerate(P["qtv"]):
            last = i == n - 1
            out = ws["J1"] if last else (ws["fx"] if i % 2 == 0 else ws["fx2"])
            self._layer_f32(L, lw, x, B * Le, Le, ws["keys"]["ref"], ws["nk"]["ref"], Le, ws, st, out=out,
                            tanh_base=ws["J0"] if last else None, out16=ws["X16"] if last else None,
                            first=(i == 0), feeds_next=not last)
            x = out

        self._mark("front")
        ground_frame, ground_box, dbg, dbg_f, dbg_o = self._grounding(L, P, ws, inp, B, Lt, F, O, Of, Le, dev, st)
        self._mark("grounding")

        # ---- bf16 answer transformer
        jm = {"ref": ws["jm_ref"], "pos": ws["jm_pos"], "neg": ws["jm_neg"]}
        N = V + O
        if self.training:
            scores_all = torch.empty(3, B, T, N, device=dev, dtype=torch.float32)
            scores = {v: scores_all[i] for i, v in enumerate(variants)}
            self._mmt_encoder(L, P, ws, variants, B, Le, st)
            ws["prev"].copy_(inp["train_prev_inds"])
            self._decode_rows_multi(L, P, ws, variants, jm, scores_all, B, Le, T, V, O, F, Lt, st)
        else:
            scores_rn = torch.empty(2, B, T, N, device=dev, dtype=torch.float32)
            forced = self.parity_hooks.get("force_prev_inds")     # test-only teacher forcing of the feedback
            forced = forced.to(dev) if forced is not None else None
            # graph replay needs fixed addresses: the chain writes a workspace buffer, the result is a copy of it
            use_graph = self.greedy_graph and forced is None and L.timing is None
            pos_out = torch.empty(B, T, N, device=dev, dtype=torch.float32)
            if use_graph:
                pos_buf = ws.get("scores_pos")
                if pos_buf is None or pos_buf.shape != (B, T, N):
                    pos_buf = ws["scores_pos"] = torch.empty(B, T, N, device=dev, dtype=torch.float32)
            else:
                pos_buf = pos_out
            scores = {"pos": pos_out, "ref": scores_rn[0], "neg": scores_rn[1]}
            ws["prev"].zero_()
            ws["prev"][:, 0] = int(self.answer_processor.BOS_IDX)
            self._mmt_encoder(L, P, ws, ("pos",), B, Le, st)
            self._mark("enc_pos")

            def greedy(stream_handle):   # drives only the `pos` variant (reference t2s.py:353, Q15)
                for t in range(T):
                    self._decode_rows(L, P, ws, "pos", jm["pos"], pos_buf, B, Le, T, V, O, F, Lt, t, 1,
                                      stream_handle)
                    L.argmax_feedback(_ptr(pos_buf), N, B, T, t, 1, N, _ptr(ws["prev"]), T, None, stream_handle)
                    if forced is not None and t + 1 < T:
                        ws["prev"][:, t + 1].copy_(forced[:, t + 1])

            def run_greedy(stream):      # `stream` is torch's current stream here
                self._run_greedy(greedy, use_graph, (id(ws), self._pack_gen, B, T, V, O), pos_buf, pos_out, stream, dev)

            if _pipelined:
                main = torch.cuda.current_stream(dev)
                side = self._side_streams.get(dev.index)
                if side is None:
                    side = self._side_streams[dev.index] = torch.cuda.Stream(device=dev, priority=-1)
                pos_ready = torch.cuda.Event()
                pos_ready.record(main)
                self._mmt_encoder(L, P, ws, ("ref", "neg"), B, Le, st, qkv0=False)     # capped through L.gemm_cap
                self._mark("enc_ref_neg")
                enc_done = torch.cuda.Event()
                enc_done.record(main)
                side.wait_event(pos_ready)
                with torch.cuda.stream(side):
                    run_greedy(side)
                    self._mark("greedy_side", side)
                    side.wait_event(enc_done)
                    self._decode_rows_multi(L, P, ws, ("ref", "neg"), jm, scores_rn, B, Le, T, V, O, F, Lt,
                                            side.cuda_stream)
                    self._mark("dec_ref_neg", side)
                    out = {
                        "ref_scores": scores["ref"], "pos_scores": scores["pos"], "neg_scores": scores["neg"],
                        "ground_box": ground_box, "ground_frame": ground_frame,
                        "frame_topk": _dev_scalar(self.frame_topk, dev),
                        "ocr_topk": _dev_scalar(self.ocr_topk, dev),
                    }
                    for t_ in (scores["pos"], scores_rn, ground_box, ground_frame):
                        t_.record_stream(side)
                    # what BaseModel.__call__ adds (base_model.py:119-149), on the stream that holds the scores
                    out["losses"] = self.losses(sample_list, out)
                    out["metrics"] = self.metrics(sample_list, out)
                    done = torch.cuda.Event()
                    done.record(side)
                self._slot_done[slot] = done
                return PendingForward(out, done, dev)
            if 0 < self.overlap_sms < n_sms:
                # the 12 greedy steps are ~300 small dependent launches that leave most SMs idle; the ref / neg
                # encoder passes are independent of them and saturate whatever SMs they are given
                main = torch.cuda.current_stream(dev)
                side = self._side_streams.get(dev.index)
                if side is None:
                    side = self._side_streams[dev.index] = torch.cuda.Stream(device=dev, priority=-1)
                # the host enqueues the ~70 encoder launches first: the ~300 decode launches are host-bound
                # (~10 us each through ctypes), and issued first they would keep the caller's stream empty meanwhile
                pos_ready = torch.cuda.Event()
                pos_ready.record(main)
                self._mmt_encoder(L, P, ws, ("ref", "neg"), B, Le, st, qkv0=False, sm_cap=n_sms - self.overlap_sms)
                self._mark("enc_ref_neg")
                side.wait_event(pos_ready)
                with torch.cuda.stream(side):
                    run_greedy(side)
                    self._mark("greedy_side", side)
                main.wait_stream(side)
            else:
                run_greedy(torch.cuda.current_stream(dev))
                self._mark("greedy")
                self._mmt_encoder(L, P, ws, ("ref", "neg"), B, Le, st, qkv0=False)
                self._mark("enc_ref_neg")
            self._decode_rows_multi(L, P, ws, ("ref", "neg"), jm, scores_rn, B, Le, T, V, O, F, Lt, st)
            self._mark("dec_ref_neg")
        if dbg:
            self.last_debug = dict(J0=ws["J0"].view(B, Le, H), J1=ws["J1"].view(B, Le, H), sim=ws["sim"],
                                   gq=ws["gq"], frame_score=dbg_f, ocr_score=dbg_o, jm_pos=ws["jm_pos"],
                                   jm_neg=ws["jm_neg"], jm_ref=ws["jm_ref"], slot=ws["slot"], prev_inds=ws["prev"])
        return {
            "ref_scores": scores["ref"], "pos_scores": scores["pos"], "neg_scores": scores["neg"],
            "ground_box": ground_box, "ground_frame": ground_frame,
            "frame_topk": _dev_scalar(self.frame_topk, dev), "ocr_topk": _dev_scalar(self.ocr_topk, dev),
        }


@registry.register_model("t2s_wo_sg")
class T2SWithoutSG(T2S):
    """Ablation "w/o spatial grounding" (reference pythia/models/t2s_wo_sg.py, registry key `t2s_wo_sg`): identical
    modules and state_dict; the positive / negative OCR sets are all slots of the grounded / the other frames."""
    MODEL = "t2s_wo_sg"
    ABLATION = "wo_sg"


@registry.register_model("t2s_wo_tg")
class T2SWithoutTG(T2S):
    """Ablation "w/o temporal grounding" (reference pythia/models/t2s_wo_tg.py, registry key `t2s_wo_tg`): the spatial
    indicator runs over the slots of every frame and the frame masks are derived from its OCR masks."""
    MODEL = "t2s_wo_tg"
    ABLATION = "wo_tg"


# =============================================================================== M4C
@registry.register_model("m4c")
class M4C(_FusionModelBase):
    MODEL = "m4c"

    def _build_grounding(self):
        self.PostHoc = PostHocAttention(self.hidden)

    # hooks of the training engine (single answer-transformer variant; vitxt_gqa_b200/train.py)
    TRAIN_VARIANT = "pos"
    TRAIN_DEAD = ("PostHoc.", "frame_embeddings.", "temporal_position_embeddings.", "track_position_embeddings.",
                  "linear_obj_frame_to_mmt_in.", "obj_frame_layer_norm.")

    def _sv_dims(self, inp):
        return 1, inp["ocr_mask"].shape[1]

    def _sv_masks(self, L, ws, inp, B, Lt, n_obj, O, st):
        ones = torch.ones(B, 1, device=inp["text"].device, dtype=torch.int64)
        L.mask_prep(_ptr(inp["text_len"]), _ptr(ones), _ptr(inp["ocr_mask"]), B, Lt, 1, O, _ptr(ws["jm_ref"]), st)
        L.mask_prep(_ptr(inp["text_len"]), None, None, B, Lt, 0, 0, _ptr(ws["jm_txt"]), st)
        L.build_keys(_ptr(ws["jm_txt"]), B, Lt, _ptr(ws["keys_txt"]), _ptr(ws["nk_txt"]), Lt, st)

    def _sv_encoder_inputs(self, inp):
        return inp, True

    def _sv_ground(self, L, P, ws, inp, B, Lt, n_obj, O, Le, dev, st):
        """PostHoc_Attention (m4c.py:356-422) on ws["J1"]; leaves the variant's joint mask / key list in ws."""
        H, f, g = 768, P["f32"], "PostHoc."
        F, Of = self.frame_num, self.ocr_frame_num
        self._q_linear(L, P, ws, f[g + "q_linear.weight"], f[g + "q_linear.bias"], "q_linear", B, Lt, Le, st)
        L.question_pool(_ptr(ws["qp"]), B, Lt, H, _ptr(f[g + "self_attn.weight"]), _ptr(f[g + "self_attn.bias"]),
                        _ptr(ws["jm_ref"]), Le, _ptr(ws["gq"]), st)
        L.sim_scores(_ptr(ws["gq"]), _ptr(ws["J1"]), Le * H, H, Lt + 1, O, H, B, _ptr(ws["sim"]), st)
        L.middle_frame_slots(_ptr(inp["middel_frame_id"]), _ptr(inp["temporal_id"]), B, O, _ptr(ws["slot"]), st)
        kk = min(self.ocr_topk, Of)
        ground_box = torch.zeros(B, kk, 4, device=dev, dtype=torch.float32)
        ws["jm_pos"].copy_(ws["jm_ref"])
        L.spatial_select(_ptr(ws["sim"]), O, 0, _ptr(ws["slot"]), _ptr(ws["jm_ref"]), B, Le, Lt + 1, F, Of, None,
                         _ptr(inp["ocr_bbox_coordinates"]), self.ocr_topk, 1, _ptr(ground_box), _ptr(ws["jm_pos"]),
                         None, None, st)
        L.build_keys(_ptr(ws["jm_pos"]), B, Le, _ptr(ws["keys"]["pos"]), _ptr(ws["nk"]["pos"]), Le, st)
        return inp["middel_frame_id"], ground_box

    def forward(self, sample_list):
        L = _lib.get_lib()
        inp = self._gather_inputs(sample_list, self._I64 + self._F32)
        dev = self._device()
        B, Lt = inp["text"].shape
        F, Of = self.frame_num, self.ocr_frame_num
        O = inp["ocr_mask"].shape[1]
        T = inp["train_prev_inds"].shape[1]
        V = self.classifier.module.weight.shape[0]
        if O != F * Of:
            raise ValueError("inputs have %d OCR slots but the config says %d x %d" % (O, F, Of))
        if self.training and torch.is_grad_enabled():
            scores, ground_frame, ground_box = self._train_forward_single(inp, dev)
            return {"pos_scores": scores, "ground_box": ground_box, "ground_frame": ground_frame,
                    "frame_topk": _dev_scalar(self.frame_topk, dev), "ocr_topk": _dev_scalar(self.ocr_topk, dev)}
        Le, H = Lt + 1 + O, 768      # one object token: the middle frame (reference m4c.py:188,420)
        P = self._pack(dev)
        variants = ("pos",)
        ws = self._workspace(B, dev, dict(Lt=Lt, F=1, O=O, T=T, V=V, k_obj_pad=P["k_obj_pad"],
                                          k_ocr_pad=P["k_ocr_pad"], variants=variants))
        st = torch.cuda.current_stream(dev).cuda_stream
        f = P["f32"]
        ones = torch.ones(B, 1, device=dev, dtype=torch.int64)
        L.mask_prep(_ptr(inp["text_len"]), _ptr(ones), _ptr(inp["ocr_mask"]), B, Lt, 1, O, _ptr(ws["jm_ref"]), st)
        L.mask_prep(_ptr(inp["text_len"]), None, None, B, Lt, 0, 0, _ptr(ws["jm_txt"]), st)
        L.build_keys(_ptr(ws["jm_txt"]), B, Lt, _ptr(ws["keys_txt"]), _ptr(ws["nk_txt"]), Lt, st)
        self._text_bert(L, P, ws, inp, B, Lt, Le, st)
        self._encode_obj_ocr(L, P, ws, inp, B, Lt, 1, O, Le, st, m4c=True)
        # no QTV in M4C: the joint buffer feeds the answer transformer directly
        ws["J1"].copy_(ws["J0"])
        L.cast_rows_bf16(_ptr(ws["J0"]), H, B * Le, H, _ptr(ws["X16"]), H, 0, 0, 0, st)
        g = "PostHoc."
        self._q_linear(L, P, ws, f[g + "q_linear.weight"], f[g + "q_linear.bias"], "q_linear", B, Lt, Le, st)
        L.question_pool(_ptr(ws["qp"]), B, Lt, H, _ptr(f[g + "self_attn.weight"]), _ptr(f[g + "self_attn.bias"]),
                        _ptr(ws["jm_ref"]), Le, _ptr(ws["gq"]), st)
        L.sim_scores(_ptr(ws["gq"]), _ptr(ws["J1"]), Le * H, H, Lt + 1, O, H, B, _ptr(ws["sim"]), st)
        L.middle_frame_slots(_ptr(inp["middel_frame_id"]), _ptr(inp["temporal_id"]), B, O, _ptr(ws["slot"]), st)
        kk = min(self.ocr_topk, Of)
        ground_box = torch.zeros(B, kk, 4, device=dev, dtype=torch.float32)
        ws["jm_pos"].copy_(ws["jm_ref"])
        L.spatial_select(_ptr(ws["sim"]), O, 0, _ptr(ws["slot"]), _ptr(ws["jm_ref"]), B, Le, Lt + 1, F, Of, None,
                         _ptr(inp["ocr_bbox_coordinates"]), self.ocr_topk, 1, _ptr(ground_box), _ptr(ws["jm_pos"]),
                         None, None, st)
        L.build_keys(_ptr(ws["jm_pos"]), B, Le, _ptr(ws["keys"]["pos"]), _ptr(ws["nk"]["pos"]), Le, st)
        N = V + O
        scores = torch.empty(B, T, N, device=dev, dtype=torch.float32)
        self._mmt_encoder(L, P, ws, variants, B, Le, st)
        if self.training:
            ws["prev"].copy_(inp["train_prev_inds"])
            self._decode_rows(L, P, ws, "pos", ws["jm_pos"], scores, B, Le, T, V, O, 1, Lt, 0, T, st)
        else:
            ws["prev"].zero_()
            ws["prev"][:, 0] = int(self.answer_processor.BOS_IDX)
            forced = self.parity_hooks.get("force_prev_inds")
            forced = forced.to(dev) if forced is not None else None
            use_graph = self.greedy_graph and forced is None and L.timing is None
            pos_buf = scores
            if use_graph:
                pos_buf = ws.get("scores_pos")
                if pos_buf is None or pos_buf.shape != (B, T, N):
                    pos_buf = ws["scores_pos"] = torch.empty(B, T, N, device=dev, dtype=torch.float32)

            def greedy(stream_handle):
                for t in range(T):
                    self._decode_rows(L, P, ws, "pos", ws["jm_pos"], pos_buf, B, Le, T, V, O, 1, Lt, t, 1, stream_handle)
                    L.argmax_feedback(_ptr(pos_buf), N, B, T, t, 1, N, _ptr(ws["prev"]), T, None, stream_handle)
                    if forced is not None and t + 1 < T:
                        ws["prev"][:, t + 1].copy_(forced[:, t + 1])

            self._run_greedy(greedy, use_graph, (id(ws), self._pack_gen, B, T, V, O), pos_buf, scores,
                             torch.cuda.current_stream(dev), dev)
        return {
            "pos_scores": scores, "ground_box": ground_box, "ground_frame": inp["middel_frame_id"],
            "frame_topk": _dev_scalar(self.frame_topk, dev), "ocr_topk": _dev_scalar(self.ocr_topk, dev),
        }


# =============================================================================== T5-ViteVQA baseline
@registry.register_model("t5vitevqa")
class T5ViteVQA(_FusionModelBase):
    """The T5-ViteVQA baseline of the reference (pythia/models/t5vitevqa.py, registry key `t5vitevqa`,
    configs/t5vitevqa_abinet.yml): M4C's single answer transformer over the question, ALL sampled frames (ViT feature +
    frame-id embedding) and all OCR tokens (FastText + PHOC + temporal / track id embeddings), with a post-hoc attention
    that reports the frame_topk * ocr_topk OCR boxes the pooled question attends to most.  Same kernels as T2S / M4C."""
    MODEL = "t5vitevqa"

    def _build_grounding(self):
        self.PostHoc = PostHocAttention(self.hidden, frame_att=True)

    # hooks of the training engine (vitxt_gqa_b200/train.py)
    TRAIN_VARIANT = "ref"
    TRAIN_DEAD = ("PostHoc.", "linear_obj_frame_to_mmt_in.", "obj_frame_layer_norm.")

    def _sv_dims(self, inp):
        return inp["video_feat"].shape[1], inp["ocr_mask"].shape[1]

    def _sv_masks(self, L, ws, inp, B, Lt, F, O, st):
        L.mask_prep(_ptr(inp["text_len"]), _ptr(inp["frame_mask"]), _ptr(inp["ocr_mask"]), B, Lt, F, O,
                    _ptr(ws["jm_ref"]), st)
        L.mask_prep(_ptr(inp["text_len"]), None, None, B, Lt, 0, 0, _ptr(ws["jm_txt"]), st)
        L.build_keys(_ptr(ws["jm_txt"]), B, Lt, _ptr(ws["keys_txt"]), _ptr(ws["nk_txt"]), Lt, st)
        L.build_keys(_ptr(ws["jm_ref"]), B, Lt + F + O, _ptr(ws["keys"]["ref"]), _ptr(ws["nk"]["ref"]), Lt + F + O, st)

    def _sv_encoder_inputs(self, inp):
        return inp, False

    def _sv_ground(self, L, P, ws, inp, B, Lt, F, O, Le, dev, st):
        """Post-hoc attention of t5vitevqa.py:357-416 on ws["J1"] (reports boxes only; masks stay the dataset's)."""
        H, f, g = 768, P["f32"], "PostHoc."
        self._q_linear(L, P, ws, f[g + "q_linear.weight"], f[g + "q_linear.bias"], "q_linear", B, Lt, Le, st)
        L.question_pool(_ptr(ws["qp"]), B, Lt, H, _ptr(f[g + "self_attn.weight"]), _ptr(f[g + "self_attn.bias"]),
                        _ptr(ws["jm_ref"]), Le, _ptr(ws["gq"]), st)
        L.sim_scores(_ptr(ws["gq"]), _ptr(ws["J1"]), Le * H, H, Lt + F, O, H, B, _ptr(ws["sim"]), st)
        K = min(self.frame_topk * self.ocr_topk, O)
        ground_box = torch.zeros(B, K, 4, device=dev, dtype=torch.float32)
        L.spatial_select(_ptr(ws["sim"]), O, 0, None, _ptr(ws["jm_ref"]), B, Le, Lt + F, F, self.ocr_frame_num, None,
                         _ptr(inp["ocr_bbox_coordinates"]), K, 4, _ptr(ground_box), None, None, None, st)
        return inp["frame_id"], ground_box

    def forward(self, sample_list):
        L = _lib.get_lib()
        inp = self._gather_inputs(sample_list, self._I64 + self._F32)
        dev = self._device()
        B, Lt = inp["text"].shape
        F, Of = self.frame_num, self.ocr_frame_num
        O = inp["ocr_mask"].shape[1]
        T = inp["train_prev_inds"].shape[1]
        V = self.classifier.module.weight.shape[0]
        if inp["video_feat"].shape[1] != F or O != F * Of:
            raise ValueError("inputs have %d frames / %d OCR slots but the config says %d x %d"
                             % (inp["video_feat"].shape[1], O, F, Of))
        if self.training and torch.is_grad_enabled():
            scores, ground_frame, ground_box = self._train_forward_single(inp, dev)
            return {"pos_scores": scores, "ground_box": ground_box, "ground_frame": ground_frame,
                    "frame_topk": _dev_scalar(self.frame_topk, dev), "ocr_topk": _dev_scalar(self.ocr_topk, dev)}
        Le, H = Lt + F + O, 768
        P = self._pack(dev)
        variants = ("ref",)          # one variant, keyed by the dataset masks (t5vitevqa.py:411-415)
        ws = self._workspace(B, dev, dict(Lt=Lt, F=F, O=O, T=T, V=V, k_obj_pad=P["k_obj_pad"],
                                          k_ocr_pad=P["k_ocr_pad"], variants=variants))
        st = torch.cuda.current_stream(dev).cuda_stream
        f = P["f32"]
        L.mask_prep(_ptr(inp["text_len"]), _ptr(inp["frame_mask"]), _ptr(inp["ocr_mask"]), B, Lt, F, O,
                    _ptr(ws["jm_ref"]), st)
        L.mask_prep(_ptr(inp["text_len"]), None, None, B, Lt, 0, 0, _ptr(ws["jm_txt"]), st)
        L.build_keys(_ptr(ws["jm_txt"]), B, Lt, _ptr(ws["keys_txt"]), _ptr(ws["nk_txt"]), Lt, st)
        L.build_keys(_ptr(ws["jm_ref"]), B, Le, _ptr(ws["keys"]["ref"]), _ptr(ws["nk"]["ref"]), Le, st)
        self._text_bert(L, P, ws, inp, B, Lt, Le, st)
        self._encode_obj_ocr(L, P, ws, inp, B, Lt, F, O, Le, st)
        ws["J1"].copy_(ws["J0"])          # no QTV: the joint buffer feeds the answer transformer directly
        L.cast_rows_bf16(_ptr(ws["J0"]), H, B * Le, H, _ptr(ws["X16"]), H, 0, 0, 0, st)
        g = "PostHoc."
        self._q_linear(L, P, ws, f[g + "q_linear.weight"], f[g + "q_linear.bias"], "q_linear", B, Lt, Le, st)
        L.question_pool(_ptr(ws["qp"]), B, Lt, H, _ptr(f[g + "self_attn.weight"]), _ptr(f[g + "self_attn.bias"]),
                        _ptr(ws["jm_ref"]), Le, _ptr(ws["gq"]), st)
        L.sim_scores(_ptr(ws["gq"]), _ptr(ws["J1"]), Le * H, H, Lt + F, O, H, B, _ptr(ws["sim"]), st)
        K = min(self.frame_topk * self.ocr_topk, O)
        ground_box = torch.zeros(B, K, 4, device=dev, dtype=torch.float32)
        L.spatial_select(_ptr(ws["sim"]), O, 0, None, _ptr(ws["jm_ref"]), B, Le, Lt + F, F, Of, None,
                         _ptr(inp["ocr_bbox_coordinates"]), K, 4, _ptr(ground_box), None, None, None, st)
        N = V + O
        scores = torch.empty(B, T, N, device=dev, dtype=torch.float32)
        self._mmt_encoder(L, P, ws, variants, B, Le, st)
        if self.training:
            ws["prev"].copy_(inp["train_prev_inds"])
            self._decode_rows(L, P, ws, "ref", ws["jm_ref"], scores, B, Le, T, V, O, F, Lt, 0, T, st)
        else:
            ws["prev"].zero_()
            ws["prev"][:, 0] = int(self.answer_processor.BOS_IDX)
            forced = self.parity_hooks.get("force_prev_inds")
            forced = forced.to(dev) if forced is not None else None
            use_graph = self.greedy_graph and forced is None and L.timing is None
            pos_buf = scores
            if use_graph:
                pos_buf = ws.get("scores_pos")
                if pos_buf is None or pos_buf.shape != (B, T, N):
                    pos_buf = ws["scores_pos"] = torch.empty(B, T, N, device=dev, dtype=torch.float32)

            def greedy(stream_handle):
                for t in range(T):
                    self._decode_rows(L, P, ws, "ref", ws["jm_ref"], pos_buf, B, Le, T, V, O, F, Lt, t, 1, stream_handle)
                    L.argmax_feedback(_ptr(pos_buf), N, B, T, t, 1, N, _ptr(ws["prev"]), T, None, stream_handle)
                    if forced is not None and t + 1 < T:
                        ws["prev"][:, t + 1].copy_(forced[:, t + 1])

            self._run_greedy(greedy, use_graph, (id(ws), self._pack_gen, B, T, V, O), pos_buf, scores,
                             torch.cuda.current_stream(dev), dev)
        return {
            "pos_scores": scores, "ground_box": ground_box, "ground_frame": inp["frame_id"],
            "frame_topk": _dev_scalar(self.frame_topk, dev), "ocr_topk": _dev_scalar(self.ocr_topk, dev),
        }


# =============================================================================== GT-box upper bound
@registry.register_model("gt_box")
class GTBox(_FusionModelBase):
    """The upper-bound model of the reference that is GIVEN the annotated frames and OCR boxes
    (pythia/models/gt_box.py, registry key `gt_box`, configs/gt_box_clipocr.yml): T2S's text / frame / OCR encoders over
    the annotated OCR fields (`ocr_temporal_id`, `ocr_track_id`, `ocr_bbox_list`), no QTV, no grounding computation
    (outputs = the annotation; `frame_topk` / `ocr_topk` are the literals 64 / 15 of gt_box.py:480-481) and one
    answer-transformer pass masked by `frame_mask_embedding` / `ocr_mask_embedding`.  Every T2S module still exists
    (TransLayer, Grounding_Module) plus a never-called LSTM, so that checkpoints interchange."""
    MODEL = "gt_box"

    def _build_grounding(self):
        cfg, h = self.config, self.hidden
        self.spatial_enhance = nn.LSTM(num_layers=2, input_size=300, hidden_size=300, batch_first=True,
                                       bidirectional=True)        # dead weights (gt_box.py:104)
        self.TransLayer = QTV(h, int(cfg.translayers.num_hidden_layers))
        self.Grounding_Module = GroundingModule(h, int(cfg.encoder.num_hidden_layers))

    # hooks of the training engine (vitxt_gqa_b200/train.py)
    TRAIN_VARIANT = "pos"
    TRAIN_DEAD = ("Grounding_Module.", "TransLayer.", "spatial_enhance.", "linear_obj_frame_to_mmt_in.",
                  "obj_frame_layer_norm.")

    def _sv_dims(self, inp):
        return inp["video_feat"].shape[1], inp["ocr_mask_embedding"].shape[1]

    def _sv_masks(self, L, ws, inp, B, Lt, F, O, st):
        L.mask_prep(_ptr(inp["text_len"]), _ptr(inp["frame_mask_embedding"]), _ptr(inp["ocr_mask_embedding"]), B, Lt, F, O,
                    _ptr(ws["jm_pos"]), st)
        L.mask_prep(_ptr(inp["text_len"]), None, None, B, Lt, 0, 0, _ptr(ws["jm_txt"]), st)
        L.build_keys(_ptr(ws["jm_txt"]), B, Lt, _ptr(ws["keys_txt"]), _ptr(ws["nk_txt"]), Lt, st)
        L.build_keys(_ptr(ws["jm_pos"]), B, Lt + F + O, _ptr(ws["keys"]["pos"]), _ptr(ws["nk"]["pos"]), Lt + F + O, st)

    def _sv_encoder_inputs(self, inp):
        return dict(inp, temporal_id=inp["ocr_temporal_id"], track_id=inp["ocr_track_id"],
                    ocr_bbox_coordinates=inp["ocr_bbox_list"]), False

    def _sv_ground(self, L, P, ws, inp, B, Lt, F, O, Le, dev, st):
        return inp["frame_list"], inp["ocr_bbox_list"]          # the annotation itself (gt_box.py:478-479)

    def forward(self, sample_list):
        L = _lib.get_lib()
        inp = self._gather_inputs(sample_list, self._I64 + self._F32)
        dev = self._device()
        B, Lt = inp["text"].shape
        F = inp["video_feat"].shape[1]
        O = inp["ocr_mask_embedding"].shape[1]
        T = inp["train_prev_inds"].shape[1]
        V = self.classifier.module.weight.shape[0]
        if self.training and torch.is_grad_enabled():
            scores, ground_frame, ground_box = self._train_forward_single(inp, dev)
            return {"pos_scores": scores, "ground_box": ground_box, "ground_frame": ground_frame,
                    "frame_topk": _dev_scalar(64, dev), "ocr_topk": _dev_scalar(15, dev)}
        Le, H = Lt + F + O, 768
        P = self._pack(dev)
        variants = ("pos",)
        ws = self._workspace(B, dev, dict(Lt=Lt, F=F, O=O, T=T, V=V, k_obj_pad=P["k_obj_pad"],
                                          k_ocr_pad=P["k_ocr_pad"], variants=variants))
        st = torch.cuda.current_stream(dev).cuda_stream
        L.mask_prep(_ptr(inp["text_len"]), _ptr(inp["frame_mask_embedding"]), _ptr(inp["ocr_mask_embedding"]), B, Lt, F, O,
                    _ptr(ws["jm_pos"]), st)
        L.mask_prep(_ptr(inp["text_len"]), None, None, B, Lt, 0, 0, _ptr(ws["jm_txt"]), st)
        L.build_keys(_ptr(ws["jm_txt"]), B, Lt, _ptr(ws["keys_txt"]), _ptr(ws["nk_txt"]), Lt, st)
        L.build_keys(_ptr(ws["jm_pos"]), B, Le, _ptr(ws["keys"]["pos"]), _ptr(ws["nk"]["pos"]), Le, st)
        self._text_bert(L, P, ws, inp, B, Lt, Le, st)
        gt_inp = dict(inp, temporal_id=inp["ocr_temporal_id"], track_id=inp["ocr_track_id"],
                      ocr_bbox_coordinates=inp["ocr_bbox_list"])
        self._encode_obj_ocr(L, P, ws, gt_inp, B, Lt, F, O, Le, st)
        ws["J1"].copy_(ws["J0"])          # no QTV (gt_box.py:298-299)
        L.cast_rows_bf16(_ptr(ws["J0"]), H, B * Le, H, _ptr(ws["X16"]), H, 0, 0, 0, st)
        N = V + O
        scores = torch.empty(B, T, N, device=dev, dtype=torch.float32)
        self._mmt_encoder(L, P, ws, variants, B, Le, st)
        if self.training:
            ws["prev"].copy_(inp["train_prev_inds"])
            self._decode_rows(L, P, ws, "pos", ws["jm_pos"], scores, B, Le, T, V, O, F, Lt, 0, T, st)
        else:
            ws["prev"].zero_()
            ws["prev"][:, 0] = int(self.answer_processor.BOS_IDX)
            forced = self.parity_hooks.get("force_prev_inds")
            forced = forced.to(dev) if forced is not None else None
            use_graph = self.greedy_graph and forced is None and L.timing is None
            pos_buf = scores
            if use_graph:
                pos_buf = ws.get("scores_pos")
                if pos_buf is None or pos_buf.shape != (B, T, N):
                    pos_buf = ws["scores_pos"] = torch.empty(B, T, N, device=dev, dtype=torch.float32)

            def greedy(stream_handle):
                for t in range(T):
                    self._decode_rows(L, P, ws, "pos", ws["jm_pos"], pos_buf, B, Le, T, V, O, F, Lt, t, 1, stream_handle)
                    L.argmax_feedback(_ptr(pos_buf), N, B, T, t, 1, N, _ptr(ws["prev"]), T, None, stream_handle)
                    if forced is not None and t + 1 < T:
                        ws["prev"][:, t + 1].copy_(forced[:, t + 1])

            self._run_greedy(greedy, use_graph, (id(ws), self._pack_gen, B, T, V, O), pos_buf, scores,
                             torch.cuda.current_stream(dev), dev)
        return {
            "pos_scores": scores, "ground_box": inp["ocr_bbox_list"], "ground_frame": inp["frame_list"],
            "frame_topk": _dev_scalar(64, dev), "ocr_topk": _dev_scalar(15, dev),
        }
