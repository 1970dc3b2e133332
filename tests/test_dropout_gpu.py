"""Dropout of the training step (reference: p = 0.1 at BertEmbeddings, BertSelfOutput / BertOutput, the attention
probabilities, obj_drop / ocr_drop and emb_dropout: models/t2s.py:95,118,688,720 + BertConfig defaults).

Masks are counter-based (csrc/common.cuh) and never stored; what has to hold, and is tested here:
  * the mask a kernel applies is the documented function of (p, seed, site, element): `t2s_dropout_mask` equals a numpy
    restatement of the hash bit for bit, its keep rate and scaling are right, sites / seeds decorrelate;
  * every forward kernel applies exactly that mask (torch reference fed with the same mask);
  * every backward kernel recomputes exactly the forward's mask (autograd of the torch reference with the same mask);
  * the whole training step with the shipped probabilities == autograd through the CPU oracle with the SAME masks
    injected at every nn.Dropout site (oracle.t2s_oracle.DROPOUT_HOOK).
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from vitxt_gqa_b200 import lib as tlib  # noqa: E402

H = 768
M32 = np.uint32


@pytest.fixture(scope="module")
def L():
    return tlib.get_lib()


def stream():
    return torch.cuda.current_stream().cuda_stream


def P(t):
    return None if t is None else t.data_ptr()


def rnd(*shape, scale=1.0, dtype=torch.float32, seed=0):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(dtype)


def rel_l2(got, ref):
    got, ref = got.double().flatten(), ref.double().flatten()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def np_hash(s0, s1, x, y):
    """csrc/common.cuh drop_hash, restated."""
    with np.errstate(over="ignore"):
        h = (x ^ M32(s0)).astype(np.uint32)
        h = h * M32(0x9E3779B1)
        h ^= h >> M32(15)
        h = h + (np.asarray(y, dtype=np.uint32) * M32(0x85EBCA6B) + M32(s1))
        h = h * M32(0xC2B2AE35)
        h ^= h >> M32(13)
        h = h * M32(0x27D4EB2F)
        h ^= h >> M32(16)
        h = h * M32(0x165667B1)
        h ^= h >> M32(15)
    return h


def rows_mask(L, rows, Hd, p, seed, site):
    out = torch.empty(rows, Hd, device="cuda")
    L.dropout_mask(P(out), 0, rows, Hd, 0, 0, p, seed, site, stream())
    return out


def attn_mask(L, BH, nq, nk, p, seed, site):
    out = torch.empty(BH, nq, nk, device="cuda")
    L.dropout_mask(P(out), 1, BH, 0, nq, nk, p, seed, site, stream())
    return out


def test_mask_is_the_documented_hash_and_has_the_right_statistics(L):
    p, seed, site = 0.1, 0x1234_5678_9ABC_DEF1, 37
    rows = 4096
    m = rows_mask(L, rows, H, p, seed, site).cpu().numpy()
    thr = round(p * 65536)
    scale = np.float32(65536.0) / np.float32(65536 - thr)
    x = np.arange(rows * (H // 2), dtype=np.uint32)
    h = np_hash(seed & 0xFFFFFFFF, seed >> 32, x, M32(site << 20))
    want = np.stack([(h & M32(0xFFFF)) >= thr, (h >> M32(16)) >= thr], -1).reshape(rows, H).astype(np.float32) * scale
    assert np.array_equal(m, want), "device mask differs from the documented hash"
    keep = m != 0
    n = keep.size
    assert abs(keep.mean() - (1 - thr / 65536)) <= 4 * math.sqrt(0.09 / n)
    assert abs(m.mean() - 1.0) <= 5 * math.sqrt(0.09 / n) * float(scale)          # E[mask] = 1: unbiased scaling
    assert abs(keep.mean(1).std() - math.sqrt(0.09 / H)) <= 0.15 * math.sqrt(0.09 / H)
    k = keep.astype(np.float64)
    for a, b in ((k[:, :-1], k[:, 1:]), (k[:-1], k[1:])):                        # neighbours are independent
        assert abs(np.corrcoef(a.ravel(), b.ravel())[0, 1]) <= 5 / math.sqrt(n)
    for other in (rows_mask(L, rows, H, p, seed, site + 1), rows_mask(L, rows, H, p, seed + 1, site),
                  rows_mask(L, rows, H, p, seed + (1 << 32), site)):              # sites / seeds are independent
        assert abs(np.corrcoef(k.ravel(), (other.cpu().numpy() != 0).astype(np.float64).ravel())[0, 1]) <= 5 / math.sqrt(n)
    # attention addressing: x = (query << 16) | (key >> 1), y = (site << 20) | bh
    BH, nq, nk = 5, 77, 131
    am = attn_mask(L, BH, nq, nk, p, seed, site).cpu().numpy()
    q = np.arange(nq, dtype=np.uint32)[:, None]
    kk = np.arange(nk, dtype=np.uint32)[None, :]
    for bh in range(BH):
        h = np_hash(seed & 0xFFFFFFFF, seed >> 32, (q << M32(16)) | (kk >> M32(1)), M32((site << 20) | bh))
        u = np.where(kk & M32(1), h >> M32(16), h & M32(0xFFFF))
        assert np.array_equal(am[bh], (u >= thr).astype(np.float32) * scale)
    assert rows_mask(L, 8, H, 0.0, seed, site).eq(1).all()                        # p = 0: identity


@pytest.mark.parametrize("x_bf16,res_bf16,split", [(0, 0, 0), (0, 0, 1), (1, 1, 0), (0, 1, 0)])
def test_add_ln_dropout_and_its_backward(L, x_bf16, res_bf16, split):
    rows, p, seed, site = 333, 0.1, 99, 5
    xd, rd = (torch.bfloat16 if x_bf16 else torch.float32), (torch.bfloat16 if res_bf16 else torch.float32)
    x, res = rnd(rows, H, dtype=xd, seed=1), rnd(rows, H, dtype=rd, seed=2)
    gamma, beta = 1 + 0.1 * rnd(H, seed=3), 0.1 * rnd(H, seed=4)
    out32 = torch.empty(rows, H, device="cuda")
    out16 = torch.empty(rows, 2 * H if split else H, device="cuda", dtype=torch.bfloat16)
    h_out = torch.empty_like(x)
    L.add_ln_dropout(P(x), x_bf16, H, P(res), res_bf16, H, P(gamma), P(beta), 1e-12, rows, H, None, 0, P(out32), H,
                     P(out16), out16.shape[1], split, 0, 0, 0, P(h_out), p, seed, site, stream())
    torch.cuda.synchronize()
    m = rows_mask(L, rows, H, p, seed, site)
    hh = (x.float() * m + res.float()).double().requires_grad_(True)
    ref = F.layer_norm(hh, (H,), gamma.double(), beta.double(), 1e-12)
    assert (h_out.float() - hh.detach().float()).abs().max().item() <= (2e-2 if x_bf16 else 1e-6)
    assert (out32 - ref.detach().float()).abs().max().item() <= (3e-2 if x_bf16 else 2e-5)
    got16 = out16[:, :H].float() + (out16[:, H:].float() if split else 0)
    assert (got16 - ref.detach().float()).abs().max().item() <= (1e-4 if split else 3e-2)
    # in place (h_out aliases x), as the training engine calls it
    x2 = x.clone()
    L.add_ln_dropout(P(x2), x_bf16, H, P(res), res_bf16, H, P(gamma), P(beta), 1e-12, rows, H, None, 0, P(out32), H,
                     None, 0, 0, 0, 0, 0, P(x2), p, seed, site, stream())
    torch.cuda.synchronize()
    assert torch.equal(x2, h_out)
    # backward: dh (residual branch) and dh_drop = dh * mask (the Linear's output), dbias = column sums of dh_drop
    dy = rnd(rows, H, dtype=torch.bfloat16, seed=6)
    ref.backward(dy.double())
    dh = torch.empty(rows, H, device="cuda", dtype=torch.bfloat16)
    dhm = torch.empty(rows, H, device="cuda", dtype=torch.bfloat16)
    dg, db, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
    L.ln_bwd_dropout(P(h_out), x_bf16, H, P(dy), 1, H, 0, 0, 0, P(gamma), P(beta), 1e-12, rows, H, 0, P(dh), 1, H,
                     P(dg), P(db), P(dbias), P(dhm), p, seed, site, stream())
    torch.cuda.synchronize()
    want_dh = hh.grad.float()
    tol = 3e-2 if x_bf16 else 1e-2          # h_out rounded to bf16 moves the statistics slightly
    assert rel_l2(dh.float(), want_dh) <= tol
    assert rel_l2(dhm.float(), want_dh * m) <= tol
    assert torch.equal(dhm.float() == 0, (m == 0) | (dh.float() == 0))
    assert rel_l2(dbias, (want_dh * m).sum(0)) <= tol


def test_dropout_rows_in_place_with_row_map(L):
    B, Le, Lt, F_, p, seed, site = 3, 50, 20, 8, 0.1, 7, 9
    for dt in (torch.float32, torch.bfloat16):
        J = rnd(B * Le, H, dtype=dt, seed=1)
        before = J.clone()
        L.dropout_rows(P(J), int(dt == torch.bfloat16), H, B * F_, H, F_, Le, Lt, p, seed, site, stream())
        torch.cuda.synchronize()
        m = rows_mask(L, B * F_, H, p, seed, site).view(B, F_, H)
        want = before.view(B, Le, H).float().clone()
        want[:, Lt:Lt + F_] *= m
        assert torch.equal(J.view(B, Le, H).float(), want.to(dt).float())
        L.dropout_rows(P(J), int(dt == torch.bfloat16), H, 4, H, 0, 0, 0, 0.0, seed, site, stream())      # p = 0: no-op


def _keys(B, Le, frac, seed):
    g = torch.Generator().manual_seed(seed)
    keep = torch.rand(B, Le, generator=g) < frac
    keep[:, 0] = True
    keep[0] = True
    key_idx = torch.zeros(B, Le, dtype=torch.int32)
    n_keys = torch.zeros(B, dtype=torch.int32)
    lists = []
    for b in range(B):
        idx = keep[b].nonzero().flatten()
        key_idx[b, :idx.numel()] = idx.int()
        n_keys[b] = idx.numel()
        lists.append(idx.cuda())
    return key_idx.cuda(), n_keys.cuda(), lists


def _attn_reference_dropout(q, k, v, lists, Le, T, mask_site):
    """Masked softmax attention with the prefix-LM rule and the dropout mask of a site: mask_site [B*heads, Le+T, nkv]
    is indexed by VIRTUAL key position (list position, then nk + decoder position) -- mapped back to rows here."""
    B, Ltot, heads, dh = q.shape
    outs = []
    for b in range(B):
        nk = lists[b].numel()
        allowed = torch.zeros(Ltot, Ltot, dtype=torch.bool, device=q.device)
        allowed[:, lists[b]] = True
        if T:
            allowed[Le:, Le:] = torch.tril(torch.ones(T, T, dtype=torch.bool, device=q.device))
        s = torch.einsum("ihd,jhd->hij", q[b], k[b]) / math.sqrt(dh)
        s = s.masked_fill(~allowed, float("-inf"))
        prob = torch.softmax(s, -1)
        m = torch.ones(heads, Ltot, Ltot, dtype=torch.float64, device=q.device)
        ms = mask_site[b * heads:(b + 1) * heads].double()
        m[:, :, lists[b]] = ms[:, :, :nk]
        if T:
            m[:, :, Le:] = ms[:, :, nk:nk + T]
        outs.append(torch.einsum("hij,jhd->ihd", prob * m, v[b]))
    return torch.stack(outs)


@pytest.mark.parametrize("B,Le,T,x3", [(2, 150, 12, False), (2, 300, 0, True), (3, 20, 0, True), (1, 1044, 12, False)])
def test_attention_dropout_forward_and_backward(L, B, Le, T, x3):
    heads, p, seed, site = 12, 0.1, 4242, 11
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv_e32 = torch.randn(B * Le, 3 * H, device="cuda", generator=g) * 0.8
    qkv_d = (torch.randn(max(B * T, 1), 3 * H, device="cuda", generator=g) * 0.8).to(torch.bfloat16)
    key_idx, n_keys, lists = _keys(B, Le, 0.7, seed=Le + T)
    nkv_max = int(n_keys.max()) + T
    ms = attn_mask(L, B * heads, Le + T, nkv_max, p, seed, site)
    # {log2-sum-exp, dO . O} per (sample, head, row): the forward kernels fill slot 0 for the backward
    stats = torch.full((B * heads * (Le + T) * 2,), float("nan"), device="cuda")
    # ---- forward: encoder rows through the tcgen05 kernel (bf16 or bf16 hi|lo), decoder rows through attn_dec
    if x3:
        hi = qkv_e32.to(torch.bfloat16)
        qs = torch.cat([hi, (qkv_e32 - hi.float()).to(torch.bfloat16)], 1).contiguous()
        o_e = torch.empty(B * Le, 2 * H, device="cuda", dtype=torch.bfloat16)
        L.attn_tc_dropout(P(qs), 6 * H, 3 * H, B, Le, H, heads, P(key_idx), P(n_keys), Le, P(o_e), 2 * H, p, seed, site,
                          P(stats), Le + T, stream())
        qkv_e = qkv_e32
    else:
        qkv_e16 = qkv_e32.to(torch.bfloat16)
        o_e = torch.empty(B * Le, H, device="cuda", dtype=torch.bfloat16)
        L.attn_tc_dropout(P(qkv_e16), 3 * H, 0, B, Le, H, heads, P(key_idx), P(n_keys), Le, P(o_e), H, p, seed, site,
                          P(stats), Le + T, stream())
        qkv_e = qkv_e16.float()
    o_d = torch.zeros(max(B * T, 1), H, device="cuda", dtype=torch.bfloat16)
    if T:
        src = qs if x3 else qkv_e16
        L.attn_dec_dropout(P(src), src.shape[1], Le, P(qkv_d), 3 * H, T, B, H, heads, P(key_idx), P(n_keys), Le, 0, T,
                           P(o_d), H, p, seed, site, P(stats), stream())
    torch.cuda.synchronize()
    assert torch.isfinite(stats.view(-1, 2)[:, 0]).all(), "a forward kernel left a row's log-sum-exp unwritten"

    def joint(e, d, width):
        e = e.view(B, Le, -1)[..., :width]
        return torch.cat([e, d[:B * T].view(B, T, width)], 1) if T else e

    qkv = joint(qkv_e.double(), qkv_d.double(), 3 * H)
    q, k, v = (qkv[..., i * H:(i + 1) * H].reshape(B, Le + T, heads, 64).clone().requires_grad_(True) for i in range(3))
    o = _attn_reference_dropout(q, k, v, lists, Le, T, ms)
    ref_o = o.detach().reshape(B, Le + T, H)
    got_e = (o_e[:, :H].float() + o_e[:, H:].float()) if x3 else o_e.float()
    err_e = (got_e.view(B, Le, H).double() - ref_o[:, :Le]).abs().max().item()
    assert err_e <= (1e-4 if x3 else 4e-2), err_e
    if T:
        err_d = (o_d[:B * T].view(B, T, H).double() - ref_o[:, Le:]).abs().max().item()
        assert err_d <= 3e-2, err_d
    # ---- backward: recomputes the same mask
    do_e = torch.randn(B * Le, H, device="cuda", generator=g).to(torch.bfloat16)
    do_d = torch.randn(max(B * T, 1), H, device="cuda", generator=g).to(torch.bfloat16)
    o.backward(joint(do_e.double(), do_d.double(), H).view(B, Le + T, heads, 64))
    o16 = ref_o.to(torch.bfloat16)
    oe16 = o16[:, :Le].reshape(B * Le, H).contiguous()
    od16 = o16[:, Le:].reshape(B * T, H).contiguous() if T else None
    qe16 = qkv_e32.to(torch.bfloat16) if x3 else qkv_e16
    dqkv_e = torch.full((B * Le, 3 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
    dqkv_d = torch.full((max(B * T, 1), 3 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
    ws = torch.empty(int(L.attn_bwd_workspace_bytes(B, Le, T, heads)), device="cuda", dtype=torch.uint8)
    ref = torch.cat([t.grad.reshape(B, Le + T, H) for t in (q, k, v)], -1)
    for stats_arg in (None, stats):       # statistics recomputed by the backward / taken from the forward kernels
        dqkv_e.fill_(float("nan"))
        dqkv_d.fill_(float("nan"))
        L.attn_bwd_dropout(P(qe16), 3 * H, P(qkv_d) if T else None, 3 * H, P(oe16), H, P(od16), H, P(do_e), H,
                           P(do_d) if T else None, H, P(dqkv_e), 3 * H, P(dqkv_d) if T else None, 3 * H, B, Le, T, H, heads,
                           P(key_idx), P(n_keys), Le, Le, P(ws), p, seed, site, P(stats_arg), stream())
        torch.cuda.synchronize()
        got = joint(dqkv_e.float(), dqkv_d.float(), 3 * H)
        assert torch.isfinite(got).all()
        for i, name in enumerate(("dq", "dk", "dv")):
            e = rel_l2(got[..., i * H:(i + 1) * H], ref[..., i * H:(i + 1) * H])
            assert e <= 3e-2, (name, e, stats_arg is not None)
    # a backward with ANOTHER site's mask is measurably wrong: the test can tell the masks apart
    L.attn_bwd_dropout(P(qe16), 3 * H, P(qkv_d) if T else None, 3 * H, P(oe16), H, P(od16), H, P(do_e), H,
                       P(do_d) if T else None, H, P(dqkv_e), 3 * H, P(dqkv_d) if T else None, 3 * H, B, Le, T, H, heads,
                       P(key_idx), P(n_keys), Le, Le, P(ws), p, seed, site + 1, None, stream())
    torch.cuda.synchronize()
    assert rel_l2(joint(dqkv_e.float(), dqkv_d.float(), 3 * H)[..., 2 * H:], ref[..., 2 * H:]) > 0.1


# ------------------------------------------------------------------------------- the whole step at the shipped p = 0.1
def _site_masks(L, eng, model, seed, B, Lt, F_, O, T):
    """name -> mask tensor for every dropout site the oracle hook asks for, built with t2s_dropout_mask from the
    engine's (seed, site id) and mapped to the oracle's tensor layouts."""
    Le = Lt + F_ + O
    ids = eng._site_ids
    cache = {}

    def rows(name, n, p):
        key = ("r", name)
        if key not in cache:
            cache[key] = rows_mask(L, n, H, p, seed, ids[name]).cpu()
        return cache[key]

    mws = [w for w in model._ws.values() if "keys" in w][0]

    def attn(name, nq, p, keys, nk, has_dec):
        key = ("a", name)
        if key in cache:
            return cache[key]
        nkv = int(nk.max()) + (T if has_dec else 0)
        ms = attn_mask(L, B * 12, nq, nkv, p, seed, ids[name]).cpu().view(B, 12, nq, nkv)
        full = torch.ones(B, 12, nq, nq)
        for b in range(B):
            n = int(nk[b])
            idx = keys[b, :n].long()
            full[b][:, :, idx] = ms[b][:, :, :n]
            if has_dec:
                full[b][:, :, Le:] = ms[b][:, :, n:n + T]
        cache[key] = full
        return full

    d = eng.drop_cfg

    def hook(site, x):
        ctx, _, name = site.rpartition("|")
        if name == "text_bert.emb":
            return x * rows("text.emb", B * Lt, d["text"][0]).view(B, Lt, H)
        if name == "obj":
            return x * rows("obj", B * F_, d["obj"]).view(B, F_, H)
        if name == "ocr":
            return x * rows("ocr", B * O, d["ocr"]).view(B, O, H)
        if name == "prev":
            return x * rows("mmt.%s.prev" % ctx, B * T, d["mmt"][0]).view(B, T, H)
        mod, li, kind = name.split(".encoder.layer.")[0], int(name.split(".layer.")[1].split(".")[0]), name.rsplit(".", 1)[1]
        if mod in ("text_bert", "TransLayer"):
            tag, n, ph, pa = ("text", Lt, *d["text"]) if mod == "text_bert" else ("qtv", Le, *d["qtv"])
            keys, nk = (mws["keys_txt"], mws["nk_txt"]) if mod == "text_bert" else (mws["keys"]["ref"], mws["nk"]["ref"])
            if kind == "attn":
                return x * attn("%s.%d.attn" % (tag, li), n, pa, keys.cpu(), nk.cpu(), False)
            return x * rows("%s.%d.%s" % (tag, li, kind), B * n, ph).view(B, n, H)
        assert mod == "mmt", site
        if kind == "attn":
            return x * attn("mmt.%s.%d.attn" % (ctx, li), Le + T, d["mmt"][1], mws["keys"][ctx].cpu(), mws["nk"][ctx].cpu(), True)
        e = rows("mmt.%s.%d.enc.%s" % (ctx, li, kind), B * Le, d["mmt"][0]).view(B, Le, H)
        dd = rows("mmt.%s.%d.dec.%s" % (ctx, li, kind), B * T, d["mmt"][0]).view(B, T, H)
        return x * torch.cat([e, dd], 1)

    return hook


def test_training_step_with_dropout_matches_oracle_autograd_under_the_same_masks(L):
    from oracle import t2s_oracle as O
    from parity_utils import build_b200_model, sample_list
    from vitxt_gqa_b200 import synth
    d = synth.Dims(frames=8, ocr_per_frame=4, vocab=200, frame_topk=3, ocr_topk=2)
    sd = synth.make_state_dict(d, seed=0, variant="stress")
    inp = synth.make_inputs(d, 3, seed=11, train=True)
    B, Lt, F_, Oc, T = 3, d.txt_len, d.frames, d.ocr, d.dec_steps
    m = build_b200_model(d, sd, train=True, dropout=True)
    eng = m.train_engine()
    eng.set_dropout(True, seed=20261017)
    assert eng.dropout_p == pytest.approx(0.1)
    sl = sample_list(inp)
    # pass 1 (device): finds the grounding the dropped-out chain produces; the oracle then runs under the same masks
    out = m(sl)
    seed = eng.saved["seed"]
    torch.cuda.synchronize()
    hook = _site_masks(L, eng, m, seed, B, Lt, F_, Oc, T)
    sd_g = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    O.DROPOUT_HOOK = hook
    try:
        ref = O.forward_t2s(sd_g, d, inp, training=True, return_debug=True)
    finally:
        O.DROPOUT_HOOK = None
    dbg = ref["debug"]
    if not torch.equal(out["ground_frame"].cpu(), ref["ground_frame"]):
        # tie order among -10000 entries (SURVEY hard part 3): rerun the device step with the oracle's choice injected
        m.parity_hooks = {"pos_frame_topk": dbg["frame_pos_topk"].float(), "neg_frame_topk": dbg["frame_neg_topk"].float()}
        eng.fwd_gen -= 1                      # same step seed as pass 1
        eng.saved = None
        out = m(sl)
        assert eng.saved["seed"] == seed
    else:
        m.parity_hooks = {"neg_frame_topk": dbg["frame_neg_topk"].float()}
        eng.fwd_gen -= 1
        eng.saved = None
        out = m(sl)
    assert torch.equal(out["ground_frame"].cpu(), ref["ground_frame"])
    for k in ("ref_scores", "pos_scores", "neg_scores"):
        err = (out[k].detach().cpu() - ref[k].detach()).abs().max().item()
        assert err <= 6e-2, (k, err)
    # dropout really is on: the same step with dropout off gives different scores
    bce = O.pos_bce_loss(ref["pos_scores"], inp["targets"], inp["train_loss_mask"])
    nce = O.info_nce(ref["ref_scores"], ref["pos_scores"], ref["neg_scores"])
    (bce + 100.0 * nce).backward()
    ref_g = {k: v.grad for k, v in sd_g.items() if v.requires_grad}
    w_cfg = {name: float(w) for name, w, _ in m.losses.losses}
    losses = out["losses"]
    bce_key = [k for k in losses if "bce" in k.lower()][0]
    nce_key = [k for k in losses if "nce" in k.lower()][0]
    (losses[bce_key] / w_cfg["pos_bce_loss"] + losses[nce_key] * (100.0 / w_cfg["InfoNCE"])).backward()
    torch.cuda.synchronize()
    worst = {}
    for name, prm in m.named_parameters():
        r = ref_g.get(name)
        if name.startswith(eng.DEAD_PREFIXES) or r is None or r.norm().item() == 0 or name.endswith("attention.self.key.bias"):
            continue
        worst[name] = rel_l2(prm.grad.cpu(), r)
    bad = {k: v for k, v in worst.items() if v > 6e-2}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:12]
    a = torch.cat([m.get_parameter(k).grad.flatten().cpu().double() for k in worst])
    b = torch.cat([ref_g[k].flatten().double() for k in worst])
    assert F.cosine_similarity(a, b, dim=0).item() >= 0.998
    # and the masks matter: with dropout off the scores differ well beyond the tolerance
    eng.set_dropout(False)
    m.parity_hooks = {}
    off = m(sl)
    assert (off["pos_scores"].detach() - out["pos_scores"].detach()).abs().max().item() > 0.1


def test_dropout_is_reproducible_per_seed_and_off_in_eval(L):
    from parity_utils import build_b200_model, sample_list
    from vitxt_gqa_b200 import synth
    d = synth.Dims(frames=8, ocr_per_frame=4, vocab=200, frame_topk=3, ocr_topk=2)
    sd = synth.make_state_dict(d, seed=0, variant="stress")
    inp = synth.make_inputs(d, 2, seed=5, train=True)
    m = build_b200_model(d, sd, train=True, dropout=True)
    eng = m.train_engine()
    sl = sample_list(inp)
    eng.set_dropout(True, seed=1)
    a = m(sl)["pos_scores"].detach().clone()
    b = m(sl)["pos_scores"].detach().clone()          # next step: new masks
    eng.set_dropout(True, seed=1)
    eng.fwd_gen = 0
    c = m(sl)["pos_scores"].detach().clone()          # same seed, same step index: same masks
    assert torch.equal(a, c) and not torch.equal(a, b)
    m.eval()
    with torch.no_grad():
        e1 = m(sl)["pos_scores"].clone()
        e2 = m(sl)["pos_scores"].clone()
    assert torch.equal(e1, e2)
