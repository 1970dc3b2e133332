"""Synthetic t2s_abinet-shaped inputs and seeded random-init weights.

There is no dataset and no checkpoint offline, so every parity / bench input is
generated here from a seeded CPU `torch.Generator` (bit-reproducible on a given
torch build).  Field names, shapes and dtypes follow what the reference dataset
hands the model (reference pythia/datasets/videoqa/vtextgqa/dataset.py:83-287,
consumed at pythia/models/t2s.py:177-258): int64 ids and masks, fp32 features,
`temporal_id` = `frame_id` repeated per OCR slot (t2s.py:486-492 requires it),
pad OCR slots all carry one identical "<pad>" feature pair (dataset.py:140).
"""
import math
from dataclasses import dataclass, field

import torch


@dataclass
class Dims:
    """Shape parameters of one config (names after reference config keys)."""
    txt_len: int = 20          # text_processor.max_length
    frames: int = 64           # grounding.frame_num
    ocr_per_frame: int = 15    # grounding.ocr_frame_num
    dec_steps: int = 12        # answer_processor.max_copy_steps
    vocab: int = 5000          # |fixed answer vocab|
    hidden: int = 768
    vit_dim: int = 1024
    ft_dim: int = 300
    phoc_dim: int = 604
    id_dim: int = 50
    frame_topk: int = 5
    ocr_topk: int = 5
    text_layers: int = 3
    qtv_layers: int = 2
    mmt_layers: int = 3
    ground_enc_layers: int = 2   # Grounding_Module.encoder: dead weights
    word_vocab: int = 30522
    model: str = "t2s"           # "t2s" | "m4c" | "t5vitevqa" (M4C over all frames with T2S's encoders, reference models/t5vitevqa.py) | "gt_box" (annotated OCR as input, reference models/gt_box.py)
    ablation: str = ""           # t2s only: "" | "wo_sg" | "wo_tg" (reference models/t2s_wo_sg.py, t2s_wo_tg.py)

    @property
    def ocr(self):
        return self.frames * self.ocr_per_frame

    @property
    def num_outputs(self):
        return self.vocab + self.ocr


def dims_from_config(model_cfg, vocab=5000, model="t2s"):
    g = model_cfg["grounding"]
    return Dims(
        frames=int(g["frame_num"]), ocr_per_frame=int(g["ocr_frame_num"]),
        vocab=vocab, frame_topk=int(g["frame_topk"]), ocr_topk=int(g["ocr_topk"]),
        text_layers=int(model_cfg["text_bert"]["num_hidden_layers"]),
        qtv_layers=int(model_cfg["translayers"]["num_hidden_layers"]),
        mmt_layers=int(model_cfg["mmt"]["num_hidden_layers"]),
        ground_enc_layers=int(model_cfg["encoder"]["num_hidden_layers"]),
        model=model,
    )


def model_config_for_dims(d: Dims):
    """`model_attributes.<model>` dict for arbitrary dims (stress sweep / small tests)."""
    t2s = d.model == "t2s"
    ids = d.model != "m4c"
    return {
        "lr_scale_frcn": 0.1, "lr_scale_text_bert": 0.1, "lr_scale_mmt": 1.0,
        "text_bert_init_from_bert_base": False,
        "text_bert": {"num_hidden_layers": d.text_layers},
        "obj": {"mmt_in_dim": d.vit_dim + (d.id_dim if ids else 0), "dropout_prob": 0.1},
        "ocr": {"mmt_in_dim": d.ft_dim + d.phoc_dim + (2 * d.id_dim if ids else 0), "dropout_prob": 0.1},
        "translayers": {"hidden_size": d.hidden, "num_hidden_layers": d.qtv_layers},
        "grounding": {"frame_topk": d.frame_topk, "ocr_topk": d.ocr_topk, "max_ocr_num": d.ocr,
                      "frame_num": d.frames, "ocr_frame_num": d.ocr_per_frame, "hidden_size": d.hidden},
        "encoder": {"hidden_size": d.hidden, "num_hidden_layers": d.ground_enc_layers},
        "mmt": {"hidden_size": d.hidden, "num_hidden_layers": d.mmt_layers},
        "classifier": {"type": "linear", "ocr_max_num": d.ocr,
                       "ocr_ptr_net": {"hidden_size": d.hidden, "query_key_size": d.hidden}, "params": {}},
        "metrics": [],
        "losses": ([{"type": "pos_bce_loss", "weight": 1.0, "params": {}},
                    {"type": "InfoNCE", "weight": 1000, "params": {}}] if t2s else
                   [{"type": "pos_bce_loss", "weight": 1.0, "params": {}}]),
    }


# --------------------------------------------------------------------------- inputs
def make_inputs(d: Dims, batch: int, seed: int = 1234, full_frames: bool = False, train: bool = False):
    """Returns an ordered dict of CPU tensors (the SampleList fields) plus the
    injected Gumbel noise `gumbel_frame [B,2,F]`, `gumbel_ocr [B,2,O]`
    (= -log(Exp(1)), the quantity F.gumbel_softmax adds to the logits)."""
    g = torch.Generator().manual_seed(seed)
    B, F, Of, O = batch, d.frames, d.ocr_per_frame, d.ocr

    def randint(lo, hi, shape):
        return torch.randint(lo, hi + 1, shape, generator=g, dtype=torch.int64)

    out = {}
    text_len = randint(5, d.txt_len, (B,))
    text = randint(1000, d.word_vocab - 1, (B, d.txt_len))
    text = text * (torch.arange(d.txt_len)[None, :] < text_len[:, None])
    out["text"], out["text_len"] = text, text_len

    lo_f = min(8, F)
    n_frames = torch.full((B,), F, dtype=torch.int64) if full_frames else randint(lo_f, F, (B,))
    max_step = max(1, min(10, 3999 // max(F, 1)))
    step = randint(1, max_step, (B,))
    pos = torch.arange(F)[None, :]
    fvalid = pos < n_frames[:, None]
    out["frame_id"] = (1 + pos * step[:, None]) * fvalid
    out["frame_mask"] = fvalid.to(torch.int64)
    vf = torch.randn(B, F, d.vit_dim, generator=g)
    out["video_feat"] = vf * fvalid[:, :, None]

    out["temporal_id"] = out["frame_id"].repeat_interleave(Of, dim=1)
    n_ocr = randint(0, Of, (B, F)) * fvalid
    slot = torch.arange(Of)[None, None, :]
    ovalid = (slot < n_ocr[:, :, None]).reshape(B, O)
    out["ocr_mask"] = ovalid.to(torch.int64)
    out["track_id"] = randint(1, 200, (B, O)) * ovalid
    pad_ft = torch.randn(d.ft_dim, generator=g) * 0.3
    pad_phoc = (torch.rand(d.phoc_dim, generator=g) < 0.04).float()
    ft = torch.randn(B, O, d.ft_dim, generator=g) * 0.3
    phoc = (torch.rand(B, O, d.phoc_dim, generator=g) < 0.04).float()
    out["context_feature_0"] = torch.where(ovalid[:, :, None], ft, pad_ft)
    out["context_feature_1"] = torch.where(ovalid[:, :, None], phoc, pad_phoc)
    c = torch.rand(B, O, 2, 2, generator=g).sort(dim=2).values   # [.., (lo,hi), (x,y)]
    box = torch.stack([c[:, :, 0, 0], c[:, :, 0, 1], c[:, :, 1, 0], c[:, :, 1, 1]], -1)
    out["ocr_bbox_coordinates"] = box * ovalid[:, :, None]

    T, n_out = d.dec_steps, d.num_outputs
    if train:
        alen = randint(1, min(6, T - 1), (B,))
        idx = randint(4, n_out - 1, (B, T))
        tpos = torch.arange(T)[None, :]
        prev = torch.where((tpos >= 1) & (tpos <= alen[:, None]), idx, torch.zeros_like(idx))
        prev[:, 0] = 1
        out["train_prev_inds"] = prev
        out["train_loss_mask"] = (tpos <= alen[:, None]).float()
    else:
        out["train_prev_inds"] = torch.zeros(B, T, dtype=torch.int64)
        alen = randint(1, min(6, T - 1), (B,))
        out["train_loss_mask"] = (torch.arange(T)[None, :] <= alen[:, None]).float()
    targets = torch.zeros(B, T, n_out)
    tgt_idx = randint(4, n_out - 1, (B, T, 2))
    targets.scatter_(2, tgt_idx, 1.0)
    out["targets"] = targets * out["train_loss_mask"][:, :, None]

    if d.model == "gt_box":
        # the dataset's human-annotated OCR fields (gt_box.py:268-274,478-486): own ids / boxes / masks, so that reading
        # the detector's fields instead would show
        keep = torch.rand(B, O, generator=g) < 0.6
        gt_valid = ovalid & keep
        out["ocr_mask_embedding"] = gt_valid.to(torch.int64)
        out["ocr_temporal_id"] = out["temporal_id"] * gt_valid
        out["ocr_track_id"] = randint(1, 200, (B, O)) * gt_valid
        c2 = torch.rand(B, O, 2, 2, generator=g).sort(dim=2).values
        box2 = torch.stack([c2[:, :, 0, 0], c2[:, :, 0, 1], c2[:, :, 1, 0], c2[:, :, 1, 1]], -1)
        out["ocr_bbox_list"] = box2 * gt_valid[:, :, None]
        fkeep = torch.rand(B, F, generator=g) < 0.7
        fkeep[:, 0] = True
        out["frame_mask_embedding"] = (fvalid & fkeep).to(torch.int64)
        out["frame_list"] = out["frame_id"] * out["frame_mask_embedding"]
    if d.model == "m4c":
        mid = (n_frames - 1) // 2
        out["middel_frame_idx"] = (mid + 1)[:, None]                      # 1-based position
        out["middel_frame_id"] = out["frame_id"].gather(1, mid[:, None])
        out["mid_img_feat"] = out["video_feat"].gather(1, mid[:, None, None].expand(B, 1, d.vit_dim))

    ef = torch.empty(B, 2, F).exponential_(generator=g)
    eo = torch.empty(B, 2, O).exponential_(generator=g)
    out["gumbel_frame"] = -ef.log()
    out["gumbel_ocr"] = -eo.log()
    return out


SAMPLE_FIELDS = ("text", "text_len", "video_feat", "frame_id", "frame_mask", "context_feature_0",
                 "context_feature_1", "temporal_id", "track_id", "ocr_bbox_coordinates", "ocr_mask",
                 "train_prev_inds", "targets", "train_loss_mask",
                 "mid_img_feat", "middel_frame_id", "middel_frame_idx",
                 "ocr_mask_embedding", "ocr_temporal_id", "ocr_track_id", "ocr_bbox_list", "frame_mask_embedding",
                 "frame_list", "ocr_token_bytes")


def to_sample_list(inputs, sample_list_cls, with_noise=True, dataset_name="vtextgqa", dataset_type="val"):
    sl = sample_list_cls()
    for k in SAMPLE_FIELDS:
        if k in inputs:
            sl.add_field(k, inputs[k])
    if with_noise:
        sl.add_field("gumbel_frame", inputs["gumbel_frame"])
        sl.add_field("gumbel_ocr", inputs["gumbel_ocr"])
    sl.add_field("dataset_name", dataset_name)
    sl.add_field("dataset_type", dataset_type)
    return sl


# --------------------------------------------------------------------------- weights
def _bert_layer_shapes(prefix, hidden, inter):
    s = {}
    for n in ("query", "key", "value"):
        s[f"{prefix}.attention.self.{n}.weight"] = (hidden, hidden)
        s[f"{prefix}.attention.self.{n}.bias"] = (hidden,)
    s[f"{prefix}.attention.output.dense.weight"] = (hidden, hidden)
    s[f"{prefix}.attention.output.dense.bias"] = (hidden,)
    s[f"{prefix}.attention.output.LayerNorm.weight"] = (hidden,)
    s[f"{prefix}.attention.output.LayerNorm.bias"] = (hidden,)
    s[f"{prefix}.intermediate.dense.weight"] = (inter, hidden)
    s[f"{prefix}.intermediate.dense.bias"] = (inter,)
    s[f"{prefix}.output.dense.weight"] = (hidden, inter)
    s[f"{prefix}.output.dense.bias"] = (hidden,)
    s[f"{prefix}.output.LayerNorm.weight"] = (hidden,)
    s[f"{prefix}.output.LayerNorm.bias"] = (hidden,)
    return s


def param_shapes(d: Dims):
    """Ordered {state_dict key: shape}; key names are the reference's
    (SURVEY 8b 'Parameter naming'; reference t2s.py:43-151,378-451,521-527,
    548-554,636-646,673-687; m4c.py equivalents)."""
    H, I = d.hidden, 4 * d.hidden
    t2s = d.model in ("t2s", "gt_box")      # gt_box keeps every T2S module (most of them unused) ...
    ids = d.model != "m4c"             # frame / temporal / track id embeddings are part of the encoder inputs
    s = {}
    s["text_bert.embeddings.word_embeddings.weight"] = (d.word_vocab, H)
    s["text_bert.embeddings.position_embeddings.weight"] = (512, H)
    s["text_bert.embeddings.token_type_embeddings.weight"] = (2, H)
    s["text_bert.embeddings.LayerNorm.weight"] = (H,)
    s["text_bert.embeddings.LayerNorm.bias"] = (H,)
    for i in range(d.text_layers):
        s.update(_bert_layer_shapes(f"text_bert.encoder.layer.{i}", H, I))
    s["frame_embeddings.weight"] = (4000, d.id_dim)
    obj_in = d.vit_dim + (d.id_dim if ids else 0)
    ocr_in = d.ft_dim + d.phoc_dim + (2 * d.id_dim if ids else 0)
    s["linear_obj_feat_to_mmt_in.weight"] = (H, obj_in)
    s["linear_obj_feat_to_mmt_in.bias"] = (H,)
    s["obj_feat_layer_norm.weight"] = (H,)
    s["obj_feat_layer_norm.bias"] = (H,)
    s["obj_frame_layer_norm.weight"] = (H,)
    s["obj_frame_layer_norm.bias"] = (H,)
    s["linear_obj_frame_to_mmt_in.weight"] = (H, d.id_dim)
    s["linear_obj_frame_to_mmt_in.bias"] = (H,)
    s["linear_ocr_feat_to_mmt_in.weight"] = (H, ocr_in)
    s["linear_ocr_feat_to_mmt_in.bias"] = (H,)
    s["linear_ocr_bbox_to_mmt_in.weight"] = (H, 4)
    s["linear_ocr_bbox_to_mmt_in.bias"] = (H,)
    s["temporal_position_embeddings.weight"] = (4000, d.id_dim)
    s["track_position_embeddings.weight"] = (4000, d.id_dim)
    s["ocr_feat_layer_norm.weight"] = (H,)
    s["ocr_feat_layer_norm.bias"] = (H,)
    s["ocr_bbox_layer_norm.weight"] = (H,)
    s["ocr_bbox_layer_norm.bias"] = (H,)
    if d.model == "gt_box":                 # ... plus an LSTM that forward never calls (gt_box.py:104)
        for layer in range(2):
            for sfx in ("", "_reverse"):
                i = 300 if layer == 0 else 600
                s[f"spatial_enhance.weight_ih_l{layer}{sfx}"] = (1200, i)
                s[f"spatial_enhance.weight_hh_l{layer}{sfx}"] = (1200, 300)
                s[f"spatial_enhance.bias_ih_l{layer}{sfx}"] = (1200,)
                s[f"spatial_enhance.bias_hh_l{layer}{sfx}"] = (1200,)
    if t2s:
        for i in range(d.qtv_layers):
            s.update(_bert_layer_shapes(f"TransLayer.encoder.layer.{i}", H, I))
        g = "Grounding_Module"
        s[f"{g}.q_linear.weight"] = (H, H)
        s[f"{g}.q_linear.bias"] = (H,)
        s[f"{g}.frame_attn.weight"] = (1, 2 * H)
        s[f"{g}.frame_attn.bias"] = (1,)
        s[f"{g}.self_attn.weight"] = (1, H)
        s[f"{g}.self_attn.bias"] = (1,)
        for ind, names in (("frame_grounding_indicator", ("frame_pos_att", "frame_neg_att")),
                           ("ocr_grounding_indicator", ("ocr_pos_att", "ocr_neg_att"))):
            for n in names:
                for lin in ("linear_q", "linear_k"):
                    s[f"{g}.{ind}.{n}.{lin}.weight"] = (H, H)
                    s[f"{g}.{ind}.{n}.{lin}.bias"] = (H,)
        for i in range(d.ground_enc_layers):
            s.update(_bert_layer_shapes(f"{g}.encoder.layer.{i}", H, I))
    else:
        g = "PostHoc"
        s[f"{g}.q_linear.weight"] = (H, H)
        s[f"{g}.q_linear.bias"] = (H,)
        s[f"{g}.self_attn.weight"] = (1, H)
        s[f"{g}.self_attn.bias"] = (1,)
        for att in (("frame_att", "ocr_att") if d.model == "t5vitevqa" else ("ocr_att",)):
            for lin in ("linear_q", "linear_k"):
                s[f"{g}.{att}.{lin}.weight"] = (H, H)
                s[f"{g}.{att}.{lin}.bias"] = (H,)
    p = "mmt.prev_pred_embeddings"
    s[f"{p}.position_embeddings.weight"] = (100, H)
    s[f"{p}.token_type_embeddings.weight"] = (5, H)
    for ln in ("ans_layer_norm", "ocr_layer_norm", "emb_layer_norm"):
        s[f"{p}.{ln}.weight"] = (H,)
        s[f"{p}.{ln}.bias"] = (H,)
    for i in range(d.mmt_layers):
        s.update(_bert_layer_shapes(f"mmt.encoder.layer.{i}", H, I))
    s["ocr_ptr_net.query.weight"] = (H, H)
    s["ocr_ptr_net.query.bias"] = (H,)
    s["ocr_ptr_net.key.weight"] = (H, H)
    s["ocr_ptr_net.key.bias"] = (H,)
    s["classifier.module.weight"] = (d.vocab, H)
    s["classifier.module.bias"] = (d.vocab,)
    return s


_BERT_PREFIXES = ("text_bert.", "TransLayer.", "mmt.")


def make_state_dict(d: Dims, seed: int = 0, variant: str = "default"):
    """Seeded random-init weights, by key name, as CPU fp32 tensors.

    Distributions follow the reference's initialisers: N(0, 0.02) for every
    Linear/Embedding inside a BertPreTrainedModel (text_bert / TransLayer / mmt,
    via init_weights), PyTorch defaults elsewhere (Linear U(+-1/sqrt(in)),
    Embedding N(0,1), LayerNorm 1/0).  variant="stress": additionally
    classifier weight x0.05 and bias 0 so the greedy decode is non-degenerate
    (SURVEY hard part 8), LayerNorm affine parameters perturbed and Linear
    biases inside BERT stacks made non-zero, so parity exercises every term.
    """
    assert variant in ("default", "stress")
    g = torch.Generator().manual_seed(seed)
    sd = {}
    shapes = param_shapes(d)
    for k, shape in shapes.items():
        is_ln = "LayerNorm" in k or "layer_norm" in k
        in_bert = k.startswith(_BERT_PREFIXES)
        if k.startswith("spatial_enhance."):        # nn.LSTM default: U(+-1/sqrt(hidden_size)) for every tensor
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(300)
        elif is_ln:
            base = torch.ones(shape) if k.endswith("weight") else torch.zeros(shape)
            if variant == "stress":
                base = base + 0.1 * torch.randn(shape, generator=g)
            t = base
        elif k.endswith("bias"):
            if in_bert:
                t = torch.zeros(shape)
                if variant == "stress":
                    t = 0.02 * torch.randn(shape, generator=g)
            else:
                w_in = shapes[k[:-4] + "weight"][1]
                t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(w_in)
        elif in_bert:
            t = torch.randn(shape, generator=g) * 0.02
        elif "embeddings.weight" in k:          # nn.Embedding default N(0,1)
            t = torch.randn(shape, generator=g)
        else:                                   # nn.Linear default: U(+-1/sqrt(fan_in))
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(shape[1])
        sd[k] = t.contiguous()
    sd["text_bert.embeddings.word_embeddings.weight"][0].zero_()     # padding_idx=0
    if variant == "stress":
        sd["classifier.module.weight"] *= 0.05
        sd["classifier.module.bias"].zero_()
    return sd


# ---------------------------------------------------------------------------------------------
# OCR token strings (input of the PHOC featuriser)
# ---------------------------------------------------------------------------------------------
_OCR_EDGE_TOKENS = [
    "<pad>", "", "a", "ab", "the", "THE", "he's", "St.", "café", "  x ", "İstanbul", "ﬁne", "K", "Straße",
    "12345678901234567890", "antidisestablishmentarianism", "ngngng", "ththth", "-", "?!", "q9", "O'Neil",
    "x" * 200, "th" * 64, "東京tower", "no.7", "A1", "zz9", "\t tab\n",
]


def make_ocr_tokens(n, seed=0, pad_ratio=0.3):
    """n seeded synthetic OCR token strings: scene-text-like words (mixed case, digits, punctuation, a few
    non-ASCII code points), the dataset's literal "<pad>" filler (vtextgqa/dataset.py:140) and the edge cases
    above.  Pure python `random`, so the list is identical everywhere."""
    import random
    rng = random.Random(seed)
    letters = "etaoinshrdlucmfwypvbgkqjxz"
    out = list(_OCR_EDGE_TOKENS[:n])
    while len(out) < n:
        r = rng.random()
        if r < pad_ratio:
            out.append("<pad>")
            continue
        ln = min(40, max(1, int(rng.expovariate(1 / 5.0)) + 1))
        w = "".join(rng.choice(letters if rng.random() < 0.85 else "0123456789") for _ in range(ln))
        if rng.random() < 0.3:
            w = w.capitalize() if rng.random() < 0.5 else w.upper()
        if rng.random() < 0.1:
            p = rng.randrange(len(w) + 1)
            w = w[:p] + rng.choice(["'", "-", ".", ",", " ", "é", "ß", "&"]) + w[p:]
        out.append(w)
    return out


def attach_ocr_tokens(inputs, seed=0, width=64):
    """Replace `context_feature_1` (the PHOC rows, 604 fp32 per OCR slot) of a synthetic batch by `ocr_token_bytes`
    (uint8 [B, O, width]): seeded scene-text-like words on the valid slots, the dataset's literal "<pad>" on the padded
    ones (vtextgqa/dataset.py:140).  The forward then builds the PHOC rows on the device.  Returns the token lists."""
    from .featurize import pack_tokens_fixed
    valid = inputs["ocr_mask"].bool()
    B, O = valid.shape
    words = [w for w in make_ocr_tokens(B * O + 64, seed=seed, pad_ratio=0.0) if w != "<pad>" and len(w.encode()) <= width]
    tokens = [[words[b * O + o] if valid[b, o] else "<pad>" for o in range(O)] for b in range(B)]
    import torch as _torch
    inputs["ocr_token_bytes"] = _torch.stack([pack_tokens_fixed(t, width) for t in tokens])
    inputs.pop("context_feature_1", None)
    return tokens


# ---------------------------------------------------------------------------------------------
# Grounding annotations + predictions (input of the grounding metrics)
# ---------------------------------------------------------------------------------------------
def make_ground_info(n, seed=0, max_spans=3):
    """n synthetic entries shaped like the reference's ground-annotation file (a list of dicts, loaded at
    modules/metrics.py:260 and read at metrics.py:269-272, m4c_evaluators.py:378-393): question_id, fps, width,
    height, spatial_temporal_gt = [{temporal_gt: [t0, t1] seconds, bbox_gt: {"<frame index>": [x1, y1, x2, y2]}}].
    Pure python `random`: identical everywhere."""
    import random
    rng = random.Random(seed)
    out = []
    for i in range(n):
        fps = rng.choice([10, 10, 25, 29.97, 30])
        width, height = rng.choice([(1280, 720), (1920, 1080), (640, 360), (720, 1280)])
        spans = []
        for _ in range(rng.randint(1, max_spans)):
            t0 = round(rng.uniform(0.0, 20.0), 2)
            t1 = round(t0 + rng.uniform(0.1, 6.0), 2)
            st, ed = int(t0 * fps) + 1, int(t1 * fps) + 1
            boxes = {}
            for fr in range(max(0, st - 2), ed + 1):
                if rng.random() < 0.7:
                    x1, y1 = rng.uniform(0, width * 0.8), rng.uniform(0, height * 0.8)
                    bw, bh = rng.uniform(8, width * 0.2), rng.uniform(6, height * 0.2)
                    box = [x1, y1, min(x1 + bw, width - 1.0), min(y1 + bh, height - 1.0)]
                    boxes[str(fr)] = [round(v, 1) for v in box] if rng.random() < 0.5 else [int(v) for v in box]
            spans.append({"temporal_gt": [t0, t1], "bbox_gt": boxes})
        out.append({"question_id": 1000 + i, "fps": fps, "width": width, "height": height,
                    "spatial_temporal_gt": spans})
    return out


def make_ground_predictions(ground_info, frame_topk, ocr_topk, n_boxes, seed=0):
    """Seeded model outputs for the entries of `ground_info`: ground_frame [B, frame_topk] int64 and ground_box
    [B, n_boxes, 4] fp32 normalised boxes (sorted corners, a few all-zero padding slots).  About 60 % of the grounded
    frames are frames whose predecessor carries a labelled box inside an annotated span (what the evaluator looks up,
    m4c_evaluators.py:389-391); the boxes the evaluator pairs with such a frame (slice i*ocr_topk .. of the list, E3
    in csrc/metrics.cu) are then tight, loose or unrelated copies of that label, so IoU lands on both sides of 0.3 and
    0.5, samples score several hits, and the last check of a sample fails after earlier ones passed."""
    import random
    import torch
    rng = random.Random(seed)
    B = len(ground_info)
    frames = torch.zeros(B, frame_topk, dtype=torch.int64)
    boxes = torch.zeros(B, n_boxes, 4, dtype=torch.float32)
    for b, g in enumerate(ground_info):
        W, H = float(g["width"]), float(g["height"])
        for j in range(n_boxes):
            if rng.random() < 0.1:
                continue                                          # a padding slot: all zeros
            x = sorted(rng.random() for _ in range(2))
            y = sorted(rng.random() for _ in range(2))
            boxes[b, j] = torch.tensor([x[0], y[0], x[1], y[1]])
        for i in range(frame_topk):
            sp = rng.choice(g["spatial_temporal_gt"])
            st = int(sp["temporal_gt"][0] * g["fps"]) + 1
            ed = int(sp["temporal_gt"][1] * g["fps"]) + 1
            labelled = [int(k) + 1 for k in sp["bbox_gt"] if st <= int(k) + 1 <= ed]
            r = rng.random()
            if r < 0.6 and labelled:
                fr = rng.choice(labelled)
                gt = sp["bbox_gt"][str(fr - 1)]
                for j in range(i * ocr_topk, min(n_boxes, (i + 1) * ocr_topk)):
                    mode = rng.random()
                    if mode < 0.35:
                        amp = 0.004                               # tight: IoU well above 0.5
                    elif mode < 0.7:
                        amp = 0.03                                # loose: IoU around 0.3 .. 0.6
                    else:
                        continue                                  # keep the unrelated box
                    jit = lambda: rng.uniform(-amp, amp)          # noqa: E731
                    box = [gt[0] / W + jit(), gt[1] / H + jit(), gt[2] / W + jit(), gt[3] / H + jit()]
                    box = [min(max(v, 0.0), 1.0) for v in box]
                    box = [min(box[0], box[2]), min(box[1], box[3]), max(box[0], box[2]), max(box[1], box[3])]
                    boxes[b, j] = torch.tensor(box)
            elif r < 0.8:
                fr = rng.randint(st, ed)                          # inside the span, label or not
            else:
                fr = rng.randint(1, 700)
            frames[b, i] = fr
    return frames, boxes


# ---------------------------------------------------------------------------------------------
# Answer side of the metrics: scores whose argmax spells an answer, OCR token lists, 10 human answers
# ---------------------------------------------------------------------------------------------
_ANSWER_WORDS = (
    "stop exit 7 seven the a an open closed coca cola coca-cola dont don't youre can't cant let's she's its "
    "o'clock oclock none zero ten 10 1,000 1,000,000 3.5 u.s.a. a.m. hello! yes? [sale] {50%} (off) a/b c\\d e_f "
    "g-h >> << @home `q` semi;colon plus+minus equal=s \"quoted\" mc'donalds somebody'd yall'd've wouldn'tve "
    "Id've main st. blvd 42nd new york taxi").split()


def make_answer_vocab(V, seed=0):
    """V answer-vocabulary words: the four specials of the answer processor (`<pad>` 0, `<s>` 1, `</s>` 2, `<unk>` 3;
    reference datasets/processors.py:994-1005) followed by seeded picks that exercise every branch of the EvalAI
    normaliser (articles, number words, contractions, each punctuation mark, commas inside numbers, periods)."""
    import random
    rng = random.Random(seed)
    words = ["<pad>", "<s>", "</s>", "<unk>"] + list(_ANSWER_WORDS)
    while len(words) < V:
        words.append(rng.choice(_ANSWER_WORDS) if rng.random() < 0.2 else "w%d" % len(words))
    return words[:V]


def make_answer_batch(B, T, V, O, seed=0):
    """Seeded inputs of the answer metrics: pos_scores [B, T, V + O] fp32 whose row-wise argmax spells a planted id
    sequence (vocabulary ids, OCR copies, EOS anywhere from step 0 to never, exact ties so that the lowest-index rule
    is visible), `ocr_tokens` (B lists of O strings), `gt_answers` (B lists of 10 strings: a seeded number of them
    agree with the planted answer so that every soft-accuracy level occurs) and the planted ids.  Pure python
    `random` + torch fills: identical everywhere."""
    import random
    import torch
    rng = random.Random(seed)
    vocab = make_answer_vocab(V, seed)
    N = V + O
    g = torch.Generator().manual_seed(seed)
    scores = torch.randn(B, T, N, generator=g)
    planted, ocr_tokens, gt_answers = [], [], []
    for b in range(B):
        toks = make_ocr_tokens(O, seed=seed * 131 + b, pad_ratio=0.2)
        ocr_tokens.append(toks)
        n_words = rng.choice([0, 1, 1, 2, 2, 3, 5, T])
        ids = []
        for t in range(T):
            if t == n_words:
                ids.append(2)                                        # EOS
            elif rng.random() < 0.4:
                ids.append(V + rng.randrange(O))                     # copy an OCR token
            else:
                ids.append(rng.randrange(4, V) if t < n_words else rng.randrange(0, N))
        planted.append(ids)
        for t, i in enumerate(ids):
            scores[b, t, i] = 9.0 + rng.random()
            if rng.random() < 0.3:                                   # an exact tie at a HIGHER index must lose
                j = rng.randrange(N)
                if j > i:
                    scores[b, t, j] = scores[b, t, i]
        words = []
        for i in ids:
            if i >= V:
                words.append(toks[i - V])
            elif i == 2:
                break
            else:
                words.append(vocab[i])
        answer = " ".join(words)
        agree = rng.choice([0, 1, 2, 3, 4, 10]) if answer.strip() else 0     # ANLS of '' against '' divides by zero
        pool = [" ".join(rng.choice(_ANSWER_WORDS) for _ in range(rng.randint(1, 3))) for _ in range(3)]
        gts = [answer if k < agree else rng.choice(pool) for k in range(10)]
        rng.shuffle(gts)
        gt_answers.append(gts)
    return {"pos_scores": scores, "ocr_tokens": ocr_tokens, "gt_answers": gt_answers, "vocab": vocab,
            "planted_ids": torch.tensor(planted, dtype=torch.int64)}


def encode_object(obj, max_size=16384):
    """The dataset's byte-tensor encoding of a python object (reference pythia/utils/objects_to_byte_tensor.py:11-31:
    two size bytes, then the pickle, zero padded to max_size)."""
    import pickle
    import torch
    raw = pickle.dumps(obj)
    if len(raw) > max_size - 2:
        raise ValueError("object too large for the byte tensor: %d bytes" % len(raw))
    out = torch.zeros(max_size, dtype=torch.uint8)
    out[0], out[1] = len(raw) // 256, len(raw) % 256
    out[2:2 + len(raw)] = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    return out


class SynthAnswerProcessor:
    """The members of the reference answer processor the metrics read (datasets/processors.py:994-1063):
    `EOS_IDX`, `BOS_IDX`, `PAD_IDX`, `get_true_vocab_size()`, `answer_vocab.idx2word(i)`."""
    PAD_IDX, BOS_IDX, EOS_IDX = 0, 1, 2

    def __init__(self, words):
        self.words = list(words)
        self.answer_vocab = self

    def idx2word(self, i):
        return self.words[i]

    def get_true_vocab_size(self):
        return len(self.words)


def make_metrics_case(B=6, T=12, V=120, O=24, frame_topk=3, ocr_topk=2, n_boxes=None, seed=0):
    """One seeded evaluation batch: annotation records (B + 3 of them, so lookups skip entries), the batch's
    question ids, the grounding outputs and the answer-side tensors / strings."""
    import random
    info = make_ground_info(B + 3, seed=seed)
    order = list(range(len(info)))
    random.Random(seed + 7).shuffle(order)
    chosen = [info[i] for i in order[:B]]
    n_boxes = frame_topk * ocr_topk + 4 if n_boxes is None else n_boxes
    gf, gb = make_ground_predictions(chosen, frame_topk, ocr_topk, n_boxes, seed=seed)
    case = make_answer_batch(B, T, V, O, seed=seed)
    case.update(records=info, question_id=[c["question_id"] for c in chosen], ground_frame=gf, ground_box=gb,
                frame_topk=frame_topk, ocr_topk=ocr_topk, V=V, O=O)
    return case


def metrics_sample_list(case, sample_list_cls, dataset_type="val", dataset_name="vtextgqa"):
    """(sample_list, model_output) of a `make_metrics_case` batch, with the fields the reference metrics read
    (modules/metrics.py:181-214, 256-275): context_tokens_enc, gt_answers_enc, frame_num, question_id, targets."""
    import torch
    B = len(case["question_id"])
    sl = sample_list_cls()
    sl.add_field("context_tokens_enc", torch.stack([encode_object(t) for t in case["ocr_tokens"]]))
    sl.add_field("gt_answers_enc", torch.stack([encode_object(a) for a in case["gt_answers"]]))
    sl.add_field("frame_num", torch.full((B,), 64, dtype=torch.int64))
    sl.add_field("question_id", torch.tensor(case["question_id"], dtype=torch.int64))
    sl.add_field("targets", torch.zeros(B, 1))
    sl.add_field("dataset_type", dataset_type)
    sl.add_field("dataset_name", dataset_name)
    out = {"pos_scores": case["pos_scores"].clone(), "ground_frame": case["ground_frame"].clone(),
           "ground_box": case["ground_box"].clone(), "frame_topk": torch.tensor(case["frame_topk"]),
           "ocr_topk": torch.tensor(case["ocr_topk"])}
    return sl, out
