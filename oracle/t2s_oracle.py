"""TEST INFRASTRUCTURE -- CPU fp32 restatement of the reference's T2S / M4C
forward + grounding + losses, written functionally over a `state_dict`.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  Nothing in vitxt_gqa_b200/ does.

Pinning: the reference has no tests or golden vectors for this path ("parity
unpinned" at the reference-test level).  This restatement is pinned
instead against outputs of the REAL reference model run in the dev container
(tests/golden/make_golden.py imports /root/reference through oracle/pt_bert.py
and dumps tests/golden/*.npz); tests/test_oracle_cpu.py checks this file
against those fixtures on every CPU run.

Every function cites the reference lines it follows (paths relative to
/root/reference/pythia).  Op order follows the reference so fp32 results agree
to reduction-order noise.  The additive attention mask constant is -10000.0
(models/t2s.py:419,534,618), the embedding LayerNorms use eps `ln_eps_embed`
(class default of BertLayerNorm = nn.LayerNorm, 1e-5: SURVEY Q17), BERT-internal
and PrevPred LayerNorms use 1e-12 (t2s.py:680-687).
"""
import math

import torch
import torch.nn.functional as F

NEG = -10000.0

# Training-step tests at dropout p > 0 inject the masks here: DROPOUT_HOOK(site, tensor) -> tensor stands in for every
# nn.Dropout of the reference's training forward (BertEmbeddings, BertSelfAttention probabilities, BertSelfOutput /
# BertOutput, obj_drop / ocr_drop t2s.py:214,253, PrevPredEmbeddings.emb_dropout t2s.py:720).  `site` = the module's
# state-dict prefix + ".emb" / ".attn" / ".h1" / ".h2" (or "obj" / "ocr" / "prev"), prefixed by DROPOUT_CONTEXT + "|"
# inside the answer transformer (the variant name: ref / pos / neg).  None (default) = dropout off, as in eval.
DROPOUT_HOOK = None
DROPOUT_CONTEXT = ""


def _drop(site, x):
    return x if DROPOUT_HOOK is None else DROPOUT_HOOK(DROPOUT_CONTEXT + "|" + site if DROPOUT_CONTEXT else site, x)


def _ln(x, sd, prefix, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _lin(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def gelu(x):
    # pytorch_transformers 1.2.0 modeling_bert.gelu (exact erf form)
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


# ------------------------------------------------------------------ BERT blocks
def bert_self_attention(sd, prefix, x, ext_mask, heads=12):
    """BertSelfAttention.forward (pytorch_transformers 1.2.0): scores =
    QK^T/sqrt(dh) + mask; softmax; .V; merge heads.  ext_mask broadcasts to
    [B, heads, Lq, Lk]."""
    B, L, H = x.shape
    dh = H // heads

    def split(t):
        return t.view(B, L, heads, dh).permute(0, 2, 1, 3)

    q = split(_lin(x, sd, prefix + ".query"))
    k = split(_lin(x, sd, prefix + ".key"))
    v = split(_lin(x, sd, prefix + ".value"))
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh)
    scores = scores + ext_mask
    probs = _drop(prefix[:-len(".attention.self")] + ".attn", F.softmax(scores, dim=-1))
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous()
    return ctx.view(B, L, H)


def bert_layer(sd, prefix, x, ext_mask, eps=1e-12):
    """BertLayer = BertAttention(self + output: dense, LN(x+res)) ->
    BertIntermediate(dense, gelu) -> BertOutput(dense, LN(x+res))."""
    ctx = bert_self_attention(sd, prefix + ".attention.self", x, ext_mask)
    a = _ln(_drop(prefix + ".h1", _lin(ctx, sd, prefix + ".attention.output.dense")) + x, sd,
            prefix + ".attention.output.LayerNorm", eps)
    inter = gelu(_lin(a, sd, prefix + ".intermediate.dense"))
    return _ln(_drop(prefix + ".h2", _lin(inter, sd, prefix + ".output.dense")) + a, sd, prefix + ".output.LayerNorm", eps)


def bert_encoder(sd, prefix, x, ext_mask, n_layers):
    for i in range(n_layers):
        x = bert_layer(sd, f"{prefix}.layer.{i}", x, ext_mask)
    return x


def get_mask(nums, max_num):
    """models/t2s.py:726-732 `_get_mask`."""
    ar = torch.arange(0, max_num).unsqueeze(0).expand(nums.size(0), -1)
    return ar.lt(nums.unsqueeze(-1)).type(torch.float32)


def text_bert(sd, d, text, txt_mask):
    """TextBert.forward (models/t2s.py:529-545) + BertEmbeddings."""
    p = "text_bert.embeddings"
    L = text.size(1)
    pos = torch.arange(L, dtype=torch.long).unsqueeze(0).expand_as(text)
    e = (F.embedding(text, sd[p + ".word_embeddings.weight"])
         + F.embedding(pos, sd[p + ".position_embeddings.weight"])
         + F.embedding(torch.zeros_like(text), sd[p + ".token_type_embeddings.weight"]))
    e = _drop("text_bert.emb", _ln(e, sd, p + ".LayerNorm", 1e-12))
    ext = (1.0 - txt_mask.unsqueeze(1).unsqueeze(2)) * NEG
    return bert_encoder(sd, "text_bert.encoder", e, ext, d.text_layers)


# ------------------------------------------------------------------ modality encoders
def encode_obj(sd, d, inp, ln_eps):
    """T2S._forward_obj_encoding (models/t2s.py:192-219); M4C variant
    m4c.py:186-211 (mid_img_feat, no frame-id embedding)."""
    if d.model == "m4c":
        x = F.normalize(inp["mid_img_feat"], dim=-1)
    else:
        x = F.normalize(inp["video_feat"], dim=-1)
        x = torch.cat([x, F.embedding(inp["frame_id"], sd["frame_embeddings.weight"])], dim=-1)
    return _drop("obj", _ln(_lin(x, sd, "linear_obj_feat_to_mmt_in"), sd, "obj_feat_layer_norm", ln_eps))


def encode_ocr(sd, d, inp, ln_eps):
    """T2S._forward_ocr_encoding (models/t2s.py:221-258); M4C m4c.py:214-248."""
    ft = F.normalize(inp["context_feature_0"], dim=-1)
    phoc = F.normalize(inp["context_feature_1"], dim=-1)
    parts = [ft, phoc]
    if d.model != "m4c":
        parts.append(F.embedding(inp["temporal_id"], sd["temporal_position_embeddings.weight"]))
        parts.append(F.embedding(inp["track_id"], sd["track_position_embeddings.weight"]))
    feat = torch.cat(parts, dim=-1)
    return _drop("ocr", _ln(_lin(feat, sd, "linear_ocr_feat_to_mmt_in"), sd, "ocr_feat_layer_norm", ln_eps)
                 + _ln(_lin(inp["ocr_bbox_coordinates"], sd, "linear_ocr_bbox_to_mmt_in"), sd,
                       "ocr_bbox_layer_norm", ln_eps))


def qtv(sd, d, txt, txt_mask, obj, obj_mask, ocr, ocr_mask):
    """QTV.forward (models/t2s.py:384-432): joint encoder, x += tanh(enc(x))."""
    x = torch.cat([txt, obj, ocr], dim=1)
    m = torch.cat([txt_mask, obj_mask, ocr_mask], dim=1)      # int64 masks promote to fp32 (Q11)
    L = m.size(1)
    ext = m.unsqueeze(1).unsqueeze(2).repeat(1, 1, L, 1)
    ext = (1.0 - ext) * NEG
    y = bert_encoder(sd, "TransLayer.encoder", x, ext, d.qtv_layers)
    nt, no = txt.size(1), obj.size(1)
    return (txt + torch.tanh(y[:, :nt]), obj + torch.tanh(y[:, nt:nt + no]),
            ocr + torch.tanh(y[:, nt + no:]))


# ------------------------------------------------------------------ grounding
def attention_score(q, k, m):
    """AttentionScore.forward (modules/spatio_temporal_grounding.py:15-23):
    no projections, no scale."""
    a = torch.bmm(q, k.transpose(-2, -1)).squeeze(1)
    a = F.softmax(a, dim=-1)
    a = a * m
    a = a / (a.sum(dim=-1, keepdim=True) + 1e-12)
    return torch.where(m == 0, -10000.0, a)


def gumbel_hard(logits, gumbels, tau=1.0):
    """F.gumbel_softmax(hard=True, dim=1) with the noise injected instead of
    drawn: y_soft = softmax((logits + g)/tau, 1); ret = onehot(argmax) - y + y."""
    y = F.softmax((logits + gumbels) / tau, dim=1)
    idx = y.max(1, keepdim=True)[1]
    hard = torch.zeros_like(logits).scatter_(1, idx, 1.0)
    return hard - y + y


def question_pool(sd, prefix, txt, txt_mask):
    """Grounding_Module q_linear + _calculate_self_attn (models/t2s.py:453-459,
    472-473): softmax over all tokens, then mask and renormalise (Q12)."""
    qp = _lin(txt, sd, prefix + ".q_linear")
    attn = _lin(qp, sd, prefix + ".self_attn").squeeze(-1)
    attn = F.softmax(attn, dim=-1)
    attn = attn * txt_mask
    attn = attn / (attn.sum(1, keepdim=True) + 1e-12)
    return torch.bmm(attn.unsqueeze(1), qp)


def temporal_indicator(q, frames, frame_mask, frame_id, topk, gumbels, neg_topk_override=None):
    """Temporal_Grounding_Indicator.forward (stg.py:34-68)."""
    B, Fn, _ = frames.shape
    pos = attention_score(q, frames, frame_mask)
    neg = attention_score(q, frames, frame_mask)
    score = torch.cat((pos.unsqueeze(1), neg.unsqueeze(1)), 1)
    hard = gumbel_hard(score, gumbels)
    pos_mask = hard[:, 0, :] * frame_mask
    neg_mask = hard[:, 1, :] * frame_mask
    pos = pos * pos_mask
    pos = torch.where(pos_mask == 0, -10000.0, pos)
    _, pidx = torch.topk(pos, topk, dim=1, largest=True, sorted=True)
    pos_topk = torch.zeros_like(pos).scatter_(1, pidx, 1)
    neg = neg * neg_mask
    neg = torch.where(neg_mask == 0, -10000.0, neg)
    if neg_topk_override is None:
        _, nidx = torch.topk(neg, topk, dim=1, largest=False, sorted=True)
        neg_topk = torch.zeros_like(neg).scatter_(1, nidx, 1)
    else:
        neg_topk = neg_topk_override
    pos_f = torch.nonzero(pos_topk, as_tuple=False)[:, 1].view(B, topk)
    ground_frame = torch.gather(frame_id, 1, pos_f)
    return ground_frame, pos_topk, neg_topk, dict(frame_score=pos, frame_pos_sel=pos_mask, frame_neg_score=neg,
                                                  frame_pos_topk=pos_topk, frame_neg_topk=neg_topk,
                                                  n_pos_frames=(pos_mask != 0).sum(1))


def spatial_indicator(q, ocr, boxes, attn_mask, o_topk, frame_num, o_frame_num, gumbels):
    """Spatial_Grounding_Indicator.forward (stg.py:79-142).  `attn_mask` is the
    grounded-frame slot mask (pads included, Q4); `pos_topk_mask` is NOT
    multiplied by it (Q2/Q3), `neg_topk_mask` is (stg.py:117)."""
    B, O, _ = ocr.shape
    pos = attention_score(q, ocr, attn_mask)
    neg = attention_score(q, ocr, attn_mask)
    score = torch.cat((pos.unsqueeze(1), neg.unsqueeze(1)), 1)
    hard = gumbel_hard(score, gumbels)
    pos_mask = hard[:, 0, :] * attn_mask
    neg_mask = hard[:, 1, :] * attn_mask
    pos = pos * pos_mask
    pos = torch.where(pos_mask == 0, -10000.0, pos)
    neg = neg * neg_mask
    neg = torch.where(neg_mask == 0, -10000.0, neg)
    rp = pos.view(B, frame_num, o_frame_num)
    _, sp = torch.sort(rp, descending=True, dim=-1, stable=True)
    pos_topk = torch.zeros_like(rp).scatter_(2, sp[:, :, :o_topk], 1).view(B, -1)
    rn = neg.view(B, frame_num, o_frame_num)
    _, sn = torch.sort(rn, descending=False, dim=-1, stable=True)
    neg_topk = torch.zeros_like(rn).scatter_(2, sn[:, :, :o_topk], 1).view(B, -1)
    neg_topk = neg_topk * attn_mask
    box_mask = pos_topk.unsqueeze(-1).expand(B, -1, 4)
    ground_box = torch.masked_select(boxes, box_mask.bool()).view(B, -1, 4)
    return ground_box, pos_topk, neg_topk, dict(ocr_score=pos, ocr_neg_score=neg)


def grounding_t2s(sd, d, inp, txt, txt_mask, frames, ocr, neg_frame_override=None):
    """Grounding_Module.forward (models/t2s.py:461-518); d.ablation selects the two ablation models, which differ
    from it only here: "wo_sg" (models/t2s_wo_sg.py:496-506) and "wo_tg" (models/t2s_wo_tg.py:477-537)."""
    B = ocr.size(0)
    frame_mask = inp["frame_mask"]
    ablation = getattr(d, "ablation", "")
    gq = question_pool(sd, "Grounding_Module", txt, txt_mask)
    dbg = {}
    if ablation == "wo_tg":
        ground_frame = inp["frame_id"]                                           # t2s_wo_tg.py:483
    else:
        ground_frame, gf_mask, nf_mask, dbg = temporal_indicator(
            gq, frames, frame_mask, inp["frame_id"], d.frame_topk, inp["gumbel_frame"], neg_frame_override)
        gf_mask = gf_mask * frame_mask
        nf_mask = nf_mask * frame_mask
    t1 = torch.where(ground_frame == 0, torch.tensor(1), ground_frame)           # Q9
    eq = torch.eq(inp["temporal_id"].unsqueeze(1), t1.unsqueeze(-1))
    new_idx = torch.nonzero(eq, as_tuple=True)[2].view(B, -1)
    new_ocr_mask = torch.zeros((B, ocr.size(1))).scatter_(1, new_idx, 1)
    boxes = inp["ocr_bbox_coordinates"]
    if ablation == "wo_sg":
        go_mask = new_ocr_mask                                                   # t2s_wo_sg.py:503
        no_mask = torch.ones_like(go_mask) - go_mask
        gbox = torch.masked_select(boxes, new_ocr_mask.unsqueeze(-1).expand(B, -1, 4).bool()).view(B, -1, 4)
    else:
        o_topk = d.frame_topk * d.ocr_topk if ablation == "wo_tg" else d.ocr_topk     # t2s_wo_tg.py:504
        gbox, go_mask, no_mask, dbg2 = spatial_indicator(
            gq, ocr, boxes, new_ocr_mask, o_topk, d.frames, d.ocr_per_frame, inp["gumbel_ocr"])
        dbg.update(dbg2)
    if ablation == "wo_tg":
        ocr_mask = inp["ocr_mask"]
        go_mask = go_mask * ocr_mask                                             # t2s_wo_tg.py:506-507
        no_mask = no_mask * ocr_mask

        def frames_of(mask):                                                     # t2s_wo_tg.py:511-535 (the 5 is literal)
            any_f = mask.view(B, d.frames, -1).any(dim=2)
            rows = []
            for i in range(B):
                idx = torch.where(any_f[i])[0]
                if len(idx) < 5:
                    idx = torch.cat([idx, torch.full((5 - len(idx),), -1, dtype=torch.long)])
                rows.append(idx[:5])
            idx = torch.stack(rows)
            fm = torch.zeros((B, d.frames), dtype=torch.long)
            fm[torch.arange(B).unsqueeze(1), idx] = 1
            return idx, fm
        ground_frame, gf_mask = frames_of(go_mask)
        _, nf_mask = frames_of(no_mask)
    dbg.update(global_q=gq, new_ocr_mask=new_ocr_mask)
    return dict(ground_frame=ground_frame, ground_bbox=gbox, pos_obj_mask=gf_mask, pos_ocr_mask=go_mask,
                neg_obj_mask=nf_mask, neg_ocr_mask=no_mask, debug=dbg)


def posthoc_m4c(sd, d, inp, txt, txt_mask, frames, ocr):
    """PostHoc_Attention.forward (models/m4c.py:356-422): deterministic."""
    B = frames.size(0)
    frame_mask, ocr_mask = inp["frame_mask"], inp["ocr_mask"]
    mid_id, mid_idx = inp["middel_frame_id"], inp["middel_frame_idx"]
    new_frame_mask = torch.zeros((B, d.frames)).scatter_(1, mid_idx - 1, 1)
    eq = torch.eq(inp["temporal_id"].unsqueeze(1), mid_id.unsqueeze(-1))
    new_idx = torch.nonzero(eq, as_tuple=True)[2].view(B, -1)
    new_ocr_mask = torch.zeros((B, ocr.size(1))).scatter_(1, new_idx, 1)
    middle_ocr_mask = new_ocr_mask * ocr_mask
    gq = question_pool(sd, "PostHoc", txt, txt_mask)
    score = attention_score(gq, ocr, ocr_mask)
    rs = score.view(B, d.frames, d.ocr_per_frame)
    _, si = torch.sort(rs, descending=True, dim=-1, stable=True)
    topk_mask = torch.zeros_like(rs).scatter_(2, si[:, :, :d.ocr_topk], 1).view(B, -1)
    g_mask = topk_mask * new_ocr_mask
    gbox = torch.masked_select(inp["ocr_bbox_coordinates"], g_mask.unsqueeze(-1).expand(B, -1, 4).bool()).view(B, -1, 4)
    g_ocr_mask = torch.masked_select(ocr_mask, g_mask.bool()).view(B, -1)
    gbox = gbox * g_ocr_mask.unsqueeze(-1).expand(B, -1, 4)
    return dict(ground_frame=mid_id, ground_bbox=gbox,
                obj_mask=torch.ones((B, 1), dtype=torch.float32), ocr_mask=middle_ocr_mask,
                debug=dict(global_q=gq, ocr_score=score))


# ------------------------------------------------------------------ MMT + output heads
def prev_pred_embeddings(sd, ans_emb, ocr_emb, prev_inds):
    """PrevPredEmbeddings.forward (models/t2s.py:690-723)."""
    p = "mmt.prev_pred_embeddings"
    B, T = prev_inds.shape
    ans_num = ans_emb.size(0)
    ans = _ln(ans_emb, sd, p + ".ans_layer_norm", 1e-12)
    ocr = _ln(ocr_emb, sd, p + ".ocr_layer_norm", 1e-12)
    cat = torch.cat([ans.unsqueeze(0).expand(B, -1, -1), ocr], dim=1)
    raw = torch.gather(cat, 1, prev_inds.unsqueeze(-1).expand(B, T, cat.size(-1)))   # == _batch_gather
    pos_ids = torch.arange(T, dtype=torch.long).unsqueeze(0).expand(B, T)
    emb = (F.embedding(pos_ids, sd[p + ".position_embeddings.weight"])
           + F.embedding(prev_inds.ge(ans_num).long(), sd[p + ".token_type_embeddings.weight"]))
    return _drop("prev", raw + _ln(emb, sd, p + ".emb_layer_norm", 1e-12))


def causal_mask(n):
    """models/t2s.py:735-742 `_get_causal_mask` (lower triangular ones)."""
    return torch.tril(torch.ones(n, n))


def mmt(sd, d, txt, txt_mask, obj, obj_mask, ocr, ocr_mask, prev_inds):
    """MMT.forward (models/t2s.py:556-633): prefix-LM mask; returns OCR and
    decoder rows of the last layer."""
    dec = prev_pred_embeddings(sd, sd["classifier.module.weight"], ocr, prev_inds)
    B, T = prev_inds.shape
    dec_mask = torch.zeros(B, T, dtype=torch.float32)
    x = torch.cat([txt, obj, ocr, dec], dim=1)
    m = torch.cat([txt_mask, obj_mask, ocr_mask, dec_mask], dim=1)
    L = m.size(1)
    ext = m.unsqueeze(1).unsqueeze(2).repeat(1, 1, L, 1)
    ext[:, :, -T:, -T:] = causal_mask(T)
    ext = (1.0 - ext) * NEG
    y = bert_encoder(sd, "mmt.encoder", x, ext, d.mmt_layers)
    nt, no = txt.size(1), obj.size(1)
    return y[:, nt + no:nt + no + ocr.size(1)], y[:, -T:]


def ocr_ptr_net(sd, dec_out, ocr_out, mask):
    """OcrPtrNet.forward (models/t2s.py:648-670): adds the RAW 0/1 mask (Q1)."""
    q = _lin(dec_out, sd, "ocr_ptr_net.query")
    k = _lin(ocr_out, sd, "ocr_ptr_net.key")
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(q.size(-1))
    return s + mask.unsqueeze(1)


def forward_output(sd, ocr_out, dec_out, mask):
    """T2S._forward_output (models/t2s.py:279-286)."""
    fixed = _lin(dec_out, sd, "classifier.module")
    return torch.cat([fixed, ocr_ptr_net(sd, dec_out, ocr_out, mask)], dim=-1)


# ------------------------------------------------------------------ full forward
def front_t2s(sd, d, inp, ln_eps_embed=1e-5, neg_frame_override=None):
    txt_mask = get_mask(inp["text_len"], inp["text"].size(1))
    txt = text_bert(sd, d, inp["text"], txt_mask)
    obj = encode_obj(sd, d, inp, ln_eps_embed)
    ocr = encode_ocr(sd, d, inp, ln_eps_embed)
    pre = dict(txt0=txt, obj0=obj, ocr0=ocr)
    txt, obj, ocr = qtv(sd, d, txt, txt_mask, obj, inp["frame_mask"], ocr, inp["ocr_mask"])
    g = grounding_t2s(sd, d, inp, txt, txt_mask, obj, ocr, neg_frame_override)
    return txt_mask, txt, obj, ocr, g, pre


def forward_t2s(sd, d, inp, training=False, schedule="literal", ln_eps_embed=1e-5, bos_idx=1,
                neg_frame_override=None, return_debug=False):
    """T2S.forward (models/t2s.py:153-175) incl. `_forward_mmt_and_output`
    (288-354).  schedule="literal": the reference's loop (3 passes x 12 steps
    in eval).  schedule="dedup": mathematically identical shortcut -- greedy
    decode on the `pos` variant only, then one `ref` and one `neg` pass with
    the final prev_inds (decoder rows are causal and encoder rows cannot see
    decoder rows, t2s.py:574-579,609-615); used where the literal loop is too
    slow (large-batch GPU parity)."""
    txt_mask, txt, obj, ocr, g, pre = front_t2s(sd, d, inp, ln_eps_embed, neg_frame_override)
    variants = {
        "ref": (inp["frame_mask"], inp["ocr_mask"]),
        "pos": (g["pos_obj_mask"], g["pos_ocr_mask"]),
        "neg": (g["neg_obj_mask"], g["neg_ocr_mask"]),
    }

    def one_pass(name, prev):
        global DROPOUT_CONTEXT
        om, cm = variants[name]
        DROPOUT_CONTEXT = name
        try:
            ocr_out, dec_out = mmt(sd, d, txt, txt_mask, obj, om, ocr, cm, prev)
        finally:
            DROPOUT_CONTEXT = ""
        return forward_output(sd, ocr_out, dec_out, cm)

    scores = {}
    if training:
        prev = inp["train_prev_inds"].clone()
        for name in ("ref", "pos", "neg"):
            scores[name] = one_pass(name, prev)
    else:
        T = inp["train_prev_inds"].size(1)
        prev = torch.zeros_like(inp["train_prev_inds"])
        prev[:, 0] = bos_idx
        if schedule == "literal":
            for _ in range(T):
                for name in ("ref", "pos", "neg"):
                    scores[name] = one_pass(name, prev)
                prev[:, 1:] = scores["pos"].argmax(dim=-1)[:, :-1]
        else:
            for _ in range(T):
                used = prev.clone()          # the literal loop's last iteration scores with this
                scores["pos"] = one_pass("pos", used)
                prev[:, 1:] = scores["pos"].argmax(dim=-1)[:, :-1]
            scores["ref"] = one_pass("ref", used)
            scores["neg"] = one_pass("neg", used)
    out = {
        "ref_scores": scores["ref"], "pos_scores": scores["pos"], "neg_scores": scores["neg"],
        "ground_box": g["ground_bbox"], "ground_frame": g["ground_frame"],
        "frame_topk": torch.tensor(d.frame_topk), "ocr_topk": torch.tensor(d.ocr_topk),
    }
    if return_debug:
        out["debug"] = dict(txt_mask=txt_mask, txt=txt, obj=obj, ocr=ocr, prev_inds=prev, **pre,
                            pos_obj_mask=g["pos_obj_mask"], pos_ocr_mask=g["pos_ocr_mask"],
                            neg_obj_mask=g["neg_obj_mask"], neg_ocr_mask=g["neg_ocr_mask"], **g["debug"])
    return out


def forward_m4c(sd, d, inp, training=False, ln_eps_embed=1e-5, bos_idx=1, return_debug=False):
    """M4C.forward (models/m4c.py:149-167): TextBert re-run per step is
    idempotent (m4c.py:257-261), so it is evaluated once here."""
    txt_mask = get_mask(inp["text_len"], inp["text"].size(1))
    txt = text_bert(sd, d, inp["text"], txt_mask)
    obj = encode_obj(sd, d, inp, ln_eps_embed)
    ocr = encode_ocr(sd, d, inp, ln_eps_embed)
    g = posthoc_m4c(sd, d, inp, txt, txt_mask, obj, ocr)

    def one_pass(prev):
        ocr_out, dec_out = mmt(sd, d, txt, txt_mask, obj, g["obj_mask"], ocr, g["ocr_mask"], prev)
        return forward_output(sd, ocr_out, dec_out, g["ocr_mask"])

    if training:
        scores = one_pass(inp["train_prev_inds"].clone())
    else:
        T = inp["train_prev_inds"].size(1)
        prev = torch.zeros_like(inp["train_prev_inds"])
        prev[:, 0] = bos_idx
        for _ in range(T):
            scores = one_pass(prev)
            prev[:, 1:] = scores.argmax(dim=-1)[:, :-1]
    out = {"pos_scores": scores, "ground_box": g["ground_bbox"], "ground_frame": g["ground_frame"],
           "frame_topk": torch.tensor(d.frame_topk), "ocr_topk": torch.tensor(d.ocr_topk)}
    if return_debug:
        out["debug"] = dict(txt=txt, obj=obj, ocr=ocr, ocr_mask=g["ocr_mask"], **g["debug"])
    return out


def posthoc_t5vitevqa(sd, d, inp, txt, txt_mask, frames, ocr):
    """PostHoc_Attention.forward of the T5-ViteVQA baseline (models/t5vitevqa.py:357-416): the frame_topk * ocr_topk OCR
    tokens with the largest post-hoc attention over ALL frames; the answer transformer sees the dataset masks."""
    B = frames.size(0)
    ocr_mask = inp["ocr_mask"]
    gq = question_pool(sd, "PostHoc", txt, txt_mask)
    score = attention_score(gq, ocr, ocr_mask)
    _, si = torch.sort(score, descending=True, dim=-1, stable=True)
    topk_mask = torch.zeros_like(score).scatter_(1, si[:, :d.ocr_topk * d.frame_topk], 1)
    gbox = torch.masked_select(inp["ocr_bbox_coordinates"], topk_mask.unsqueeze(-1).expand(B, -1, 4).bool()).view(B, -1, 4)
    g_ocr_mask = torch.masked_select(ocr_mask, topk_mask.bool()).view(B, -1)
    gbox = gbox * g_ocr_mask.unsqueeze(-1).expand(B, -1, 4)
    return dict(ground_frame=inp["frame_id"], ground_bbox=gbox, obj_mask=inp["frame_mask"], ocr_mask=ocr_mask,
                debug=dict(global_q=gq, ocr_score=score))


def forward_t5vitevqa(sd, d, inp, training=False, ln_eps_embed=1e-5, bos_idx=1, return_debug=False):
    """T5VITEVQA.forward (models/t5vitevqa.py:151-169): M4C's single-variant answer transformer over the question,
    all frames (ViT + frame-id embedding) and all OCR tokens (FastText + PHOC + temporal / track ids)."""
    txt_mask = get_mask(inp["text_len"], inp["text"].size(1))
    txt = text_bert(sd, d, inp["text"], txt_mask)
    obj = encode_obj(sd, d, inp, ln_eps_embed)
    ocr = encode_ocr(sd, d, inp, ln_eps_embed)
    g = posthoc_t5vitevqa(sd, d, inp, txt, txt_mask, obj, ocr)

    def one_pass(prev):
        ocr_out, dec_out = mmt(sd, d, txt, txt_mask, obj, g["obj_mask"], ocr, g["ocr_mask"], prev)
        return forward_output(sd, ocr_out, dec_out, g["ocr_mask"])

    if training:
        scores = one_pass(inp["train_prev_inds"].clone())
    else:
        T = inp["train_prev_inds"].size(1)
        prev = torch.zeros_like(inp["train_prev_inds"])
        prev[:, 0] = bos_idx
        for _ in range(T):
            scores = one_pass(prev)
            prev[:, 1:] = scores.argmax(dim=-1)[:, :-1]
    out = {"pos_scores": scores, "ground_box": g["ground_bbox"], "ground_frame": g["ground_frame"],
           "frame_topk": torch.tensor(d.frame_topk), "ocr_topk": torch.tensor(d.ocr_topk)}
    if return_debug:
        out["debug"] = dict(txt=txt, obj=obj, ocr=ocr, ocr_mask=g["ocr_mask"], **g["debug"])
    return out


def forward_gt_box(sd, d, inp, training=False, ln_eps_embed=1e-5, bos_idx=1):
    """GTBOX.forward (models/gt_box.py:158-175): the upper-bound model that is GIVEN the annotated frames / OCR boxes.
    OCR encoder over the annotated fields (gt_box.py:262-288), no QTV, no grounding computation (gt_box.py:474-488:
    outputs = the annotation, frame_topk / ocr_topk = the literals 64 / 15), one answer-transformer pass masked by
    frame_mask_embedding / ocr_mask_embedding."""
    txt_mask = get_mask(inp["text_len"], inp["text"].size(1))
    txt = text_bert(sd, d, inp["text"], txt_mask)
    obj = encode_obj(sd, d, inp, ln_eps_embed)
    gt_inp = dict(inp, temporal_id=inp["ocr_temporal_id"], track_id=inp["ocr_track_id"],
                  ocr_bbox_coordinates=inp["ocr_bbox_list"])
    ocr = encode_ocr(sd, d, gt_inp, ln_eps_embed)
    om, cm = inp["frame_mask_embedding"], inp["ocr_mask_embedding"]

    def one_pass(prev):
        ocr_out, dec_out = mmt(sd, d, txt, txt_mask, obj, om, ocr, cm, prev)
        return forward_output(sd, ocr_out, dec_out, cm)

    if training:
        scores = one_pass(inp["train_prev_inds"].clone())
    else:
        T = inp["train_prev_inds"].size(1)
        prev = torch.zeros_like(inp["train_prev_inds"])
        prev[:, 0] = bos_idx
        for _ in range(T):
            scores = one_pass(prev)
            prev[:, 1:] = scores.argmax(dim=-1)[:, :-1]
    return {"pos_scores": scores, "ground_box": inp["ocr_bbox_list"], "ground_frame": inp["frame_list"],
            "frame_topk": torch.tensor(64), "ocr_topk": torch.tensor(15)}


# ------------------------------------------------------------------ losses
def pos_bce_loss(pos_scores, targets, loss_mask):
    """POSBCEWithMaskLoss.forward (modules/losses.py:329-343)."""
    losses = F.binary_cross_entropy_with_logits(pos_scores, targets, reduction="none")
    losses = losses * loss_mask.unsqueeze(-1)
    count = torch.max(torch.sum(loss_mask), torch.tensor(1.0))
    return torch.sum(losses) / count


def info_nce(ref, pos, neg, temperature=0.1):
    """InfoNCE.forward (modules/losses.py:361-385): per-sample 2-way softmax
    between cos(ref,pos) and cos(ref,neg); temperature from the forward default
    (Q19)."""
    q, p, n = (F.normalize(x, dim=-1) for x in (ref, pos, neg))
    B = q.size(0)
    q, p, n = q.view(B, -1), p.view(B, -1), n.view(B, -1)
    pl = F.cosine_similarity(q, p, dim=1).unsqueeze(1)
    nl = F.cosine_similarity(q, n, dim=1).unsqueeze(1)
    logits = torch.cat([pl, nl], dim=1)
    labels = torch.zeros(B, dtype=torch.long)
    return F.cross_entropy(logits / temperature, labels, reduction="mean")
