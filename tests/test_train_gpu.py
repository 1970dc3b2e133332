"""GPU tests of the training step: every backward kernel through the C ABI against torch
autograd of the same op, then the whole step -- `model(sample_list)` in training mode, the registered losses,
`loss.backward()` -- against autograd through the CPU oracle (oracle/t2s_oracle.py, fp32) on the same seeded
inputs and weights.

Tolerances: activation gradients are bf16 (2^-8 relative rounding per stored value) with fp32 accumulation, so
parameter gradients are compared by relative L2 error; kernels with fp32 outputs are held to 1e-4-class.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from vitxt_gqa_b200 import lib as tlib  # noqa: E402

H = 768


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


def gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


# ------------------------------------------------------------------------------- weight-gradient GEMM (MN-major operands)
@pytest.mark.parametrize("rows,Pn,Qn,splits", [
    (64, 128, 256, 1), (128, 128, 64, 1), (1000, 768, 768, 0), (4176, 2304, 768, 0), (2088, 768, 3072, 7),
    (24, 200, 768, 0), (576, 5000, 768, 0), (3072, 768, 1088, 0), (130, 100, 70, 2),
])
def test_gemm_wgrad(L, rows, Pn, Qn, splits):
    ldg, ldx = (Pn + 7) // 8 * 8 + 8 * (rows % 3), 2 * ((Qn + 7) // 8 * 8)       # pitched operands, like the hi|lo buffers
    G = rnd(rows, ldg, dtype=torch.bfloat16, seed=1)
    X = rnd(rows, ldx, dtype=torch.bfloat16, seed=2)
    ldd = (Qn + 3) // 4 * 4
    base = rnd(Pn, ldd, seed=3)
    dW = base.clone()
    L.gemm_wgrad_bf16(P(G), ldg, P(X), ldx, P(dW), ldd, rows, Pn, Qn, splits, stream())
    torch.cuda.synchronize()
    ref = G[:, :Pn].float().t() @ X[:, :Qn].float()
    got = dW[:, :Qn] - base[:, :Qn]
    assert torch.equal(dW[:, Qn:], base[:, Qn:]), "columns beyond Q were touched"
    tol = 1e-3 * ref.abs().max().item() + 1e-4          # fp32 accumulation, order varies with the split
    assert (got - ref).abs().max().item() <= tol, ((got - ref).abs().max().item(), tol)


def test_gemm_dgelu_epilogue(L):
    M, N, K = 300, 3072, 768
    A = rnd(M, K, dtype=torch.bfloat16, seed=1)
    W = rnd(N, K, scale=0.05, dtype=torch.bfloat16, seed=2)
    for aux_f32 in (False, True):
        u = rnd(M, N, scale=1.5, seed=3)
        aux = u if aux_f32 else u.to(torch.bfloat16)
        C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        L.gemm_bf16(P(A), K, P(W), K, None, P(aux), N, P(C), N, M, N, K,
                    tlib.GEMM_DGELU | (tlib.GEMM_RES_F32 if aux_f32 else 0), 0, stream())
        torch.cuda.synchronize()
        uu = aux.float().requires_grad_(True)
        gelu(uu).sum().backward()
        ref = (A.float() @ W.float().t()) * uu.grad
        assert (C.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item() + 1e-3


def test_gemm_accumulate_in_place(L):
    M, N, K = 200, 768, 768
    A = rnd(M, K, dtype=torch.bfloat16, seed=1)
    W = rnd(N, K, scale=0.05, dtype=torch.bfloat16, seed=2)
    C = rnd(M, N, dtype=torch.bfloat16, seed=3)
    ref = A.float() @ W.float().t() + C.float()
    L.gemm_bf16(P(A), K, P(W), K, None, P(C), N, P(C), N, M, N, K, 0, 0, stream())
    torch.cuda.synchronize()
    assert (C.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item() + 1e-3


# ------------------------------------------------------------------------------- LayerNorm backward
@pytest.mark.parametrize("h_bf16,dy_bf16,tanh_out", [(1, 1, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)])
def test_ln_bwd(L, h_bf16, dy_bf16, tanh_out):
    rows = 1044 * 2 + 5
    h = rnd(rows, H, scale=2.0, seed=1)
    h = h.to(torch.bfloat16) if h_bf16 else h
    dy = rnd(rows, H, seed=2)
    dy = dy.to(torch.bfloat16) if dy_bf16 else dy
    gamma, beta = 1.0 + 0.1 * rnd(H, seed=3), 0.1 * rnd(H, seed=4)
    eps = 1e-12
    dh = torch.empty(rows, H, device="cuda", dtype=torch.bfloat16)
    dg, db, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
    L.ln_bwd(P(h), h_bf16, H, P(dy), dy_bf16, H, 0, 0, 0, P(gamma), P(beta), eps, rows, H, tanh_out, P(dh), 1, H,
             P(dg), P(db), P(dbias), stream())
    torch.cuda.synchronize()
    hh = h.double().requires_grad_(True)
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    y = F.layer_norm(hh, (H,), g64, b64, eps)
    out = torch.tanh(y) if tanh_out else y
    out.backward(dy.double())
    assert rel_l2(dh.float(), hh.grad) <= 6e-3          # bf16 store
    assert rel_l2(dg, g64.grad) <= 1e-4 and rel_l2(db, b64.grad) <= 1e-4
    assert rel_l2(dbias, hh.grad.sum(0)) <= 2e-3


def test_ln_bwd_row_gather(L):
    B, per, group, off = 3, 20, 50, 7
    rows = B * per
    h = rnd(rows, H, seed=1)
    big = rnd(B * group, H, seed=2)
    gamma, beta = 1.0 + 0.1 * rnd(H, seed=3), 0.1 * rnd(H, seed=4)
    dh = torch.empty(rows, H, device="cuda", dtype=torch.float32)
    dg, db = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    L.ln_bwd(P(h), 0, H, P(big), 0, H, per, group, off, P(gamma), P(beta), 1e-5, rows, H, 0, P(dh), 0, H, P(dg), P(db),
             None, stream())
    torch.cuda.synchronize()
    dy = big.view(B, group, H)[:, off:off + per].reshape(rows, H)
    hh = h.double().requires_grad_(True)
    F.layer_norm(hh, (H,), gamma.double(), beta.double(), 1e-5).backward(dy.double())
    assert rel_l2(dh, hh.grad) <= 1e-5


# ------------------------------------------------------------------------------- small row kernels
def test_colsum_rows_add_gelu_scatter(L):
    rows, N = 1500, 2304
    x = rnd(rows, N + 8, dtype=torch.bfloat16, seed=1)
    dst = torch.ones(N, device="cuda")
    L.colsum(P(x), 1, N + 8, rows, N, P(dst), stream())
    ref = 1.0 + x[:, :N].float().sum(0)
    assert (dst - ref).abs().max().item() <= 1e-3 * ref.abs().max().item() + 1e-3
    xf = rnd(rows, 300, seed=2)
    dst = torch.zeros(300, device="cuda")
    L.colsum(P(xf), 0, 300, rows, 300, P(dst), stream())
    assert (dst - xf.sum(0)).abs().max().item() <= 1e-3

    a, b, c = (rnd(40, H, dtype=torch.bfloat16, seed=s) for s in (3, 4, 5))
    out = rnd(4 * 30, H, seed=6)
    ref = out.clone().view(4, 30, H)
    ref[:, 5:15] += (a.float() + b.float() + c.float()).view(4, 10, H)
    L.rows_add(P(a), P(b), P(c), H, 40, H, P(out), H, 10, 30, 5, 1, stream())
    torch.cuda.synchronize()
    assert torch.allclose(out.view(4, 30, H), ref, atol=1e-6)

    u = rnd(200, 3072, scale=2.0, seed=7)
    o = torch.empty(200, 2 * 3072, device="cuda", dtype=torch.bfloat16)
    L.gelu_rows(P(u), 0, 3072, 200, 3072, P(o), 2 * 3072, 3072, stream())
    torch.cuda.synchronize()
    rec = o[:, :3072].float() + o[:, 3072:].float()
    assert (rec - gelu(u)).abs().max().item() <= 2 ** -15 * gelu(u).abs().max().item() + 1e-6     # hi + lo: 16 mantissa bits
    ub = u.to(torch.bfloat16)
    o2 = torch.empty(200, 3072, device="cuda", dtype=torch.bfloat16)
    L.gelu_rows(P(ub), 1, 3072, 200, 3072, P(o2), 3072, 0, stream())
    torch.cuda.synchronize()
    assert (o2.float() - gelu(ub.float())).abs().max().item() <= 2 ** -8 * 8 + 1e-3

    ids = torch.randint(0, 37, (500,), device="cuda")
    src = rnd(500, 100, seed=8)
    table = torch.zeros(37, 50, device="cuda")
    L.embed_scatter_add(P(src), 0, 100, 50, 50, P(ids), 500, 3, P(table), 50, stream())
    torch.cuda.synchronize()
    ref = torch.zeros(37, 50, device="cuda").index_add_(0, ids[ids != 3], src[ids != 3][:, 50:])
    assert torch.allclose(table, ref, atol=1e-4)


# ------------------------------------------------------------------------------- attention backward
def _attn_reference(q, k, v, key_lists, Le, T):
    """q, k, v: [B, Le+T, heads, 64] fp64 with requires_grad; masked softmax attention with the prefix-LM rule."""
    B, Ltot, heads, dh = q.shape
    outs = []
    for b in range(B):
        allowed = torch.zeros(Ltot, Ltot, dtype=torch.bool, device=q.device)
        allowed[:, key_lists[b]] = True
        if T:
            allowed[Le:, Le:] = torch.tril(torch.ones(T, T, dtype=torch.bool, device=q.device))
        s = torch.einsum("ihd,jhd->hij", q[b], k[b]) / math.sqrt(dh)
        s = s.masked_fill(~allowed, float("-inf"))
        outs.append(torch.einsum("hij,jhd->ihd", torch.softmax(s, -1), v[b]))
    return torch.stack(outs)


@pytest.mark.parametrize("B,Le,T,heads_used", [(2, 150, 12, 12), (3, 100, 0, 12), (2, 20, 0, 12), (1, 1044, 12, 12)])
def test_attn_bwd(L, B, Le, T, heads_used):
    heads = 12
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv_e = (torch.randn(B * Le, 3 * H, device="cuda", generator=g) * 0.8).to(torch.bfloat16)
    qkv_d = (torch.randn(max(B * T, 1), 3 * H, device="cuda", generator=g) * 0.8).to(torch.bfloat16)
    do_e = torch.randn(B * Le, H, device="cuda", generator=g).to(torch.bfloat16)
    do_d = torch.randn(max(B * T, 1), H, device="cuda", generator=g).to(torch.bfloat16)
    key_idx = torch.zeros(B, Le, dtype=torch.int32, device="cuda")
    n_keys = torch.zeros(B, dtype=torch.int32, device="cuda")
    lists = []
    for b in range(B):
        keep = torch.rand(Le, device="cuda", generator=g) < (0.7 if b else 1.0)
        keep[0] = True
        idx = torch.nonzero(keep).flatten()
        key_idx[b, :idx.numel()] = idx.int()
        n_keys[b] = idx.numel()
        lists.append(idx)

    def joint(e, d, width):
        e = e.view(B, Le, width)
        return torch.cat([e, d[:B * T].view(B, T, width)], 1) if T else e

    qkv = joint(qkv_e, qkv_d, 3 * H).double()
    q, k, v = (qkv[..., i * H:(i + 1) * H].reshape(B, Le + T, heads, 64).clone().requires_grad_(True) for i in range(3))
    o = _attn_reference(q, k, v, lists, Le, T)
    do = joint(do_e, do_d, H).double().view(B, Le + T, heads, 64)
    o.backward(do)
    o16 = o.detach().reshape(B, Le + T, H).to(torch.bfloat16)
    o_e = o16[:, :Le].reshape(B * Le, H).contiguous()
    o_d = o16[:, Le:].reshape(B * T, H).contiguous() if T else None
    dqkv_e = torch.full((B * Le, 3 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
    dqkv_d = torch.full((max(B * T, 1), 3 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
    ws = torch.empty(int(L.attn_bwd_workspace_bytes(B, Le, T, heads)), device="cuda", dtype=torch.uint8)
    L.attn_bwd(P(qkv_e), 3 * H, P(qkv_d) if T else None, 3 * H, P(o_e), H, P(o_d), H, P(do_e), H, P(do_d) if T else None, H,
               P(dqkv_e), 3 * H, P(dqkv_d) if T else None, 3 * H, B, Le, T, H, heads, P(key_idx), P(n_keys), Le, Le, P(ws),
               stream())
    torch.cuda.synchronize()
    ref = torch.cat([t.grad.reshape(B, Le + T, H) for t in (q, k, v)], -1)
    got = joint(dqkv_e, dqkv_d, 3 * H).float()
    assert torch.isfinite(got).all()
    for i, name in enumerate(("dq", "dk", "dv")):
        e = rel_l2(got[..., i * H:(i + 1) * H], ref[..., i * H:(i + 1) * H])
        assert e <= 2e-2, (name, e)          # bf16 P / dS operands and bf16 stores
    # rows that are not in a key list get exactly zero dK / dV
    for b in range(B):
        off = torch.ones(Le, dtype=torch.bool, device="cuda")
        off[lists[b]] = False
        assert (got[b, :Le][off][:, H:] == 0).all()


# ------------------------------------------------------------------------------- heads / embeddings
def test_ptr_score_bwd(L):
    B, T, V, O = 3, 12, 200, 960
    N = V + O
    dS = rnd(B, T, N, seed=1)
    q = rnd(B * T, H, dtype=torch.bfloat16, seed=2)
    Le = 20 + O
    keyp = rnd(B * Le, H, dtype=torch.bfloat16, seed=3)
    dq = torch.empty(B * T, H, device="cuda", dtype=torch.bfloat16)
    dk = torch.zeros(B * Le, H, device="cuda", dtype=torch.bfloat16)
    L.ptr_score_bwd(P(dS), N, B, T, V, P(q), H, keyp.data_ptr() + 20 * H * 2, Le * H, H, O, H, P(dq), H,
                    dk.data_ptr() + 20 * H * 2, Le * H, H, stream())
    torch.cuda.synchronize()
    qq = q.double().view(B, T, H).requires_grad_(True)
    kk = keyp.double().view(B, Le, H)[:, 20:].clone().requires_grad_(True)
    (torch.matmul(qq, kk.transpose(1, 2)) / math.sqrt(H)).backward(dS[:, :, V:].double())
    assert rel_l2(dq.float().view(B, T, H), qq.grad) <= 6e-3
    assert rel_l2(dk.float().view(B, Le, H)[:, 20:], kk.grad) <= 6e-3
    assert (dk.view(B, Le, H)[:, :20] == 0).all()


def test_prev_embed_bwd(L):
    B, T, V, O = 4, 12, 50, 30
    Le = 20 + 8 + O
    ans_w, ocr = rnd(V, H, seed=1), rnd(B, Le, H, seed=2)
    pos_emb, type_emb = rnd(100, H, seed=3), rnd(5, H, seed=4)
    gs = [1.0 + 0.1 * rnd(H, seed=s) for s in (5, 6, 7)]
    prev = torch.randint(0, V + O, (B, T), device="cuda")
    prev[0, :3] = prev[1, 0]          # repeated indices exercise the atomics
    dx = rnd(B * T, H, dtype=torch.bfloat16, seed=8)
    d = {k: torch.zeros_like(t) for k, t in dict(ans=ans_w, ocr=ocr, pos=pos_emb, type=type_emb).items()}
    dg = [torch.zeros(H, device="cuda") for _ in range(6)]
    row0 = 28
    L.prev_embed_bwd(P(dx), H, P(prev), T, B, T, V, H, P(ans_w), ocr.data_ptr() + row0 * H * 4, Le * H, H, P(pos_emb),
                     P(type_emb), P(gs[0]), P(gs[1]), P(gs[2]), 1e-12, P(d["ans"]), d["ocr"].data_ptr() + row0 * H * 4,
                     P(d["pos"]), P(d["type"]), P(dg[0]), P(dg[1]), P(dg[2]), P(dg[3]), P(dg[4]), P(dg[5]), O, stream())
    torch.cuda.synchronize()
    t64 = [t.double().requires_grad_(True) for t in (ans_w, ocr, pos_emb, type_emb)]
    g64 = [t.double().requires_grad_(True) for t in gs]
    b64 = [torch.zeros(H, device="cuda", dtype=torch.float64, requires_grad=True) for _ in range(3)]
    ans = F.layer_norm(t64[0], (H,), g64[0], b64[0], 1e-12)
    oc = F.layer_norm(t64[1][:, row0:row0 + O], (H,), g64[1], b64[1], 1e-12)
    cat = torch.cat([ans.unsqueeze(0).expand(B, -1, -1), oc], 1)
    raw = torch.gather(cat, 1, prev.unsqueeze(-1).expand(B, T, H))
    pos_ids = torch.arange(T, device="cuda").unsqueeze(0).expand(B, T)
    emb = F.embedding(pos_ids, t64[2]) + F.embedding((prev >= V).long(), t64[3])
    out = raw + F.layer_norm(emb, (H,), g64[2], b64[2], 1e-12)
    out.backward(dx.double().view(B, T, H))
    for name, got, ref in (("ans", d["ans"], t64[0].grad), ("ocr", d["ocr"], t64[1].grad), ("pos", d["pos"], t64[2].grad),
                           ("type", d["type"], t64[3].grad), ("g_ans", dg[0], g64[0].grad), ("b_ans", dg[1], b64[0].grad),
                           ("g_ocr", dg[2], g64[1].grad), ("b_ocr", dg[3], b64[1].grad), ("g_emb", dg[4], g64[2].grad),
                           ("b_emb", dg[5], b64[2].grad)):
        assert rel_l2(got, ref) <= 1e-4, name


def test_ocr_finish_bwd_and_bert_embed_bwd(L):
    B, O, Le = 2, 60, 100
    rows = B * O
    h, bbox = rnd(rows, H, seed=1), torch.rand(rows, 4, device="cuda")
    w2, b2 = rnd(H, 4, seed=2), rnd(H, seed=3)
    g1, g2 = 1.0 + 0.1 * rnd(H, seed=4), 1.0 + 0.1 * rnd(H, seed=5)
    dout = rnd(B * Le, H, seed=6)
    dh = torch.empty(rows, H, device="cuda", dtype=torch.bfloat16)
    dc = torch.empty(rows, H, device="cuda")
    outs = [torch.zeros(H, device="cuda") for _ in range(5)] + [torch.zeros(H, 4, device="cuda"), torch.zeros(H, device="cuda")]
    L.ocr_finish_bwd(P(h), H, P(bbox), P(w2), P(b2), P(g1), P(g2), 1e-5, rows, H, P(dout), H, O, Le, 30, P(dh), H, P(dc), H,
                     *[P(t) for t in outs], stream())
    torch.cuda.synchronize()
    t = [x.double().requires_grad_(True) for x in (h, w2, b2, g1, g2)]
    z = [torch.zeros(H, device="cuda", dtype=torch.float64, requires_grad=True) for _ in range(2)]
    y = F.layer_norm(t[0], (H,), t[3], z[0], 1e-5) + F.layer_norm(F.linear(bbox.double(), t[1], t[2]), (H,), t[4], z[1], 1e-5)
    y.backward(dout.double().view(B, Le, H)[:, 30:30 + O].reshape(rows, H))
    assert rel_l2(dh.float(), t[0].grad) <= 6e-3
    for name, got, ref in (("dg1", outs[0], t[3].grad), ("db1", outs[1], z[0].grad), ("dg2", outs[2], t[4].grad),
                           ("db2ln", outs[3], z[1].grad), ("dbias1", outs[4], t[0].grad.sum(0)), ("dw2", outs[5], t[1].grad),
                           ("db2", outs[6], t[2].grad)):
        assert rel_l2(got, ref) <= 2e-4, name

    Lt, rows = 20, 3 * 20
    ids = torch.randint(0, 500, (rows,), device="cuda")
    ids[5:9] = 0
    word, pos, typ = rnd(500, H, seed=7), rnd(512, H, seed=8), rnd(2, H, seed=9)
    gamma = 1.0 + 0.1 * rnd(H, seed=10)
    dy = rnd(rows, H, dtype=torch.bfloat16, seed=11)
    dword, dpos, dtyp = torch.zeros_like(word), torch.zeros_like(pos), torch.zeros_like(typ)
    dg, db = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    L.bert_embed_bwd(P(dy), H, P(ids), rows, Lt, H, P(word), P(pos), P(typ), P(gamma), 1e-12, P(dword), P(dpos), P(dtyp),
                     P(dg), P(db), stream())
    torch.cuda.synchronize()
    t = [x.double().requires_grad_(True) for x in (word, pos, typ, gamma)]
    zb = torch.zeros(H, device="cuda", dtype=torch.float64, requires_grad=True)
    pid = torch.arange(rows, device="cuda") % Lt
    e = F.embedding(ids, t[0], padding_idx=0) + F.embedding(pid, t[1]) + F.embedding(torch.zeros_like(ids), t[2])
    F.layer_norm(e, (H,), t[3], zb, 1e-12).backward(dy.double())
    for name, got, ref in (("word", dword, t[0].grad), ("pos", dpos, t[1].grad), ("type", dtyp, t[2].grad),
                           ("gamma", dg, t[3].grad), ("beta", db, zb.grad)):
        assert rel_l2(got, ref) <= 1e-4, name


# ------------------------------------------------------------------------------- losses, optimizer
def test_loss_backward_kernels(L):
    from oracle import t2s_oracle as O
    B, T, N = 5, 12, 1160
    ref, pos, neg = (rnd(B, T, N, seed=s) for s in (1, 2, 3))
    targets = (torch.rand(B, T, N, device="cuda") < 0.01).float()
    mask = (torch.rand(B, T, device="cuda") < 0.6).float()
    go = torch.tensor([100.0], device="cuda")
    d = torch.empty_like(pos)
    L.pos_bce_loss_bwd(P(pos), P(targets), P(mask), B, T, N, P(go), P(d), 0, stream())
    x = pos.double().cpu().requires_grad_(True)
    (100.0 * O.pos_bce_loss(x, targets.double().cpu(), mask.double().cpu())).backward()
    assert rel_l2(d.cpu(), x.grad) <= 1e-5
    ws = torch.empty(int(L.loss_bwd_workspace_bytes(B, T)), device="cuda", dtype=torch.uint8)
    dr, dp, dn = (torch.empty_like(ref) for _ in range(3))
    L.info_nce_loss_bwd(P(ref), P(pos), P(neg), B, T, N, 0.1, P(ws), P(go), P(dr), P(dp), P(dn), 0, stream())
    torch.cuda.synchronize()
    t = [a.double().cpu().requires_grad_(True) for a in (ref, pos, neg)]
    (100.0 * O.info_nce(*t)).backward()
    for got, r in zip((dr, dp, dn), t):
        assert rel_l2(got.cpu(), r.grad) <= 1e-4


def test_sumsq_and_adam(L):
    n = 1_000_003
    p, g = rnd(n, seed=1), rnd(n, scale=0.01, seed=2)
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    ws = torch.empty(1024, device="cuda", dtype=torch.float64)
    ss = torch.zeros(1, device="cuda")
    for step in (1, 2, 3):
        L.sumsq(P(g), n, P(ws), P(ss), stream())
        L.adam_step(P(p), P(g), P(m), P(v), n, 1e-3, 0.9, 0.999, 1e-8, step, P(ss), 0.25, 1.0, stream())
        ref_p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([ref_p], 0.25)
        opt.step()
        torch.cuda.synchronize()
        assert abs(ss.item() - (g.double() ** 2).sum().item()) <= 1e-5 * ss.item()
        assert (p - ref_p.data).abs().max().item() <= 2e-6


# ------------------------------------------------------------------------------- the whole training step vs the oracle
def _oracle_grads(sd, d, inp, weights):
    from oracle import t2s_oracle as O
    # fp32, as the reference trains (its reduction-order noise, ~1e-6 relative, is far below the bf16 tolerance)
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    inp64 = inp
    out = O.forward_t2s(sd, d, inp64, training=True, return_debug=True)
    bce = O.pos_bce_loss(out["pos_scores"], inp64["targets"], inp64["train_loss_mask"])
    nce = O.info_nce(out["ref_scores"], out["pos_scores"], out["neg_scores"])
    (weights[0] * bce + weights[1] * nce).backward()
    return out, bce.item(), nce.item(), {k: v.grad for k, v in sd.items() if v.requires_grad}


@pytest.mark.parametrize("nce_weight", [100.0, 0.0])
def test_training_step_gradients_match_oracle_autograd(nce_weight):
    from parity_utils import build_b200_model, sample_list
    from vitxt_gqa_b200 import synth
    d = synth.Dims(frames=8, ocr_per_frame=4, vocab=200, frame_topk=3, ocr_topk=2)
    sd = synth.make_state_dict(d, seed=0, variant="stress")
    inp = synth.make_inputs(d, 3, seed=11, train=True)
    ref_out, bce, nce, ref_g = _oracle_grads(sd, d, inp, (1.0, nce_weight))
    m = build_b200_model(d, sd, train=True)
    dbg = ref_out["debug"]
    m.parity_hooks = {"pos_frame_topk": dbg["frame_pos_topk"].float(), "neg_frame_topk": dbg["frame_neg_topk"].float()}
    sl = sample_list(inp)
    out = m(sl)
    assert torch.equal(out["ground_frame"].cpu(), ref_out["ground_frame"])
    for k in ("ref_scores", "pos_scores", "neg_scores"):
        assert (out[k].detach().cpu() - ref_out[k].detach()).abs().max().item() <= 5e-2, k
    losses = out["losses"]
    names = sorted(losses)
    bce_key = [k for k in names if "bce" in k.lower()][0]
    nce_key = [k for k in names if "nce" in k.lower()][0]
    # the registered losses carry the config weights; rescale to the weights under test
    w_cfg = {name: float(w) for name, w, _ in m.losses.losses}
    assert abs(losses[bce_key].item() / w_cfg["pos_bce_loss"] - bce) <= 1e-2 * abs(bce) + 1e-4
    assert abs(losses[nce_key].item() / w_cfg["InfoNCE"] - nce) <= 1e-2
    total = losses[bce_key] / w_cfg["pos_bce_loss"] + losses[nce_key] * (nce_weight / w_cfg["InfoNCE"])
    total.backward()
    torch.cuda.synchronize()
    eng = m.train_engine()
    worst = {}
    for name, p in m.named_parameters():
        ref = ref_g.get(name)
        if name.startswith(eng.DEAD_PREFIXES):
            assert p.grad is None or p.grad.abs().max().item() == 0, name
            assert ref is None or ref.abs().max().item() == 0, name
            continue
        assert p.grad is not None, name
        if ref is None or ref.norm().item() == 0:
            assert p.grad.abs().max().item() <= 1e-6, name
            continue
        if name.endswith("attention.self.key.bias"):
            # softmax is invariant to a per-query constant, so the key bias has a mathematically zero gradient:
            # both sides hold rounding noise only; require it to be small against the query-bias gradient
            qref = ref_g[name.replace(".key.", ".query.")].abs().max().item()
            assert p.grad.abs().max().item() <= 2e-2 * qref + 1e-6, (name, p.grad.abs().max().item(), qref)
            continue
        worst[name] = rel_l2(p.grad.cpu(), ref)
    bad = {k: v for k, v in worst.items() if v > 5e-2}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:12]
    # global direction: cosine between the flat gradients
    a = torch.cat([m.get_parameter(k).grad.flatten().cpu().double() for k in worst])
    b = torch.cat([ref_g[k].flatten().double() for k in worst])
    assert F.cosine_similarity(a, b, dim=0).item() >= 0.999


def test_training_step_fused_adam_matches_torch_adam():
    from parity_utils import build_b200_model, sample_list
    from vitxt_gqa_b200 import synth
    d = synth.Dims(frames=8, ocr_per_frame=4, vocab=200, frame_topk=3, ocr_topk=2)
    sd = synth.make_state_dict(d, seed=0, variant="stress")
    inp = synth.make_inputs(d, 2, seed=12, train=True)
    m = build_b200_model(d, sd, train=True)
    sl = sample_list(inp)
    out = m(sl)
    sum(out["losses"].values()).backward()
    eng = m.train_engine()
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    # torch reference: clip_grad_norm_ over every gradient, Adam with the three groups of t2s.py:356-376
    ps = {n: torch.nn.Parameter(before[n].clone()) for n in grads}
    for n in ps:
        ps[n].grad = grads[n].clone()
    torch.nn.utils.clip_grad_norm_(list(ps.values()), 0.25)
    groups = [{"params": [ps[n] for n in ps if not n.startswith(("text_bert.", "mmt."))], "lr": 1e-4},
              {"params": [ps[n] for n in ps if n.startswith("text_bert.")], "lr": 1e-5},
              {"params": [ps[n] for n in ps if n.startswith("mmt.")], "lr": 1e-4}]
    finetune_text = any(f["module"] is m.text_bert for f in m.finetune_modules)
    if not finetune_text:
        groups[1]["lr"] = 1e-4
    torch.optim.Adam(groups, lr=1e-4, eps=1e-8).step()
    eng.step(lr=1e-4, lr_scale_text_bert=0.1, lr_scale_mmt=1.0, max_grad_l2_norm=0.25)
    torch.cuda.synchronize()
    for n, p in m.named_parameters():
        if n in ps:
            assert (p.detach() - ps[n].detach()).abs().max().item() <= 1e-6, n
        else:
            assert torch.equal(p.detach(), before[n]), n
    # and the model still runs (weights repacked)
    m.eval()
    with torch.no_grad():
        m(sl)


def test_lr_schedule_is_the_references():
    """general.py:20-30 with the training_parameters of configs/t2s_clipocr.yml."""
    from vitxt_gqa_b200.train import lr_lambda_update
    cfg = {"training_parameters": {"use_warmup": True, "warmup_iterations": 1000, "warmup_factor": 0.2,
                                   "lr_steps": [10000, 20000], "lr_ratio": 0.1}}
    assert lr_lambda_update(0, cfg) == pytest.approx(0.2) and lr_lambda_update(500, cfg) == pytest.approx(0.6)
    assert lr_lambda_update(1000, cfg) == pytest.approx(1.0) and lr_lambda_update(1001, cfg) == 1.0
    assert lr_lambda_update(10000, cfg) == pytest.approx(0.1) and lr_lambda_update(25000, cfg) == pytest.approx(0.01)


def test_optimizer_state_and_checkpoint_interchange_with_torch_adam(tmp_path):
    """The fused Adam's state exported in torch.optim.Adam's format (what the reference's checkpoint stores,
    checkpoint.py:226-232): a torch Adam that loads it takes the SAME next step as the engine, and a checkpoint
    round trip restores parameters, moments and step count."""
    from parity_utils import build_b200_model, sample_list
    from vitxt_gqa_b200 import synth
    from vitxt_gqa_b200.pythia_api import ConfigNode
    d = synth.Dims(frames=8, ocr_per_frame=4, vocab=200, frame_topk=3, ocr_topk=2)
    m = build_b200_model(d, synth.make_state_dict(d, seed=0, variant="stress"), train=True)
    cfg = ConfigNode({"optimizer_attributes": {"params": {"lr": 1e-4}}})
    eng = m.train_engine()

    def fwd_bwd(seed):
        for p in m.parameters():
            p.grad = None
        out = m(sample_list(synth.make_inputs(d, 2, seed=seed, train=True)))
        sum(out["losses"].values()).backward()

    for it in range(2):                       # two engine steps build non-trivial moments
        fwd_bwd(30 + it)
        eng.step(lr=1e-4, max_grad_l2_norm=0.25)
    sd = eng.optimizer_state_dict(cfg)
    assert len(sd["state"]) == len(eng.live_names) and all(int(s["step"]) == 2 for s in sd["state"].values())
    path = str(tmp_path / "ckpt.pth")
    eng.save_checkpoint(path, cfg, best_iteration=2, best_metric_value=0.5)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"model", "optimizer", "best_iteration", "best_metric_value", "config"}
    assert set(ck["model"]) == set(m.state_dict())

    # third step, two ways from the same parameters / gradients / moments
    fwd_bwd(40)
    p0 = {n: p.detach().clone() for n, p in m.named_parameters()}
    g0 = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    opt = torch.optim.Adam(m.get_optimizer_parameters(cfg), lr=1e-4, eps=1e-8)
    opt.load_state_dict(sd)
    torch.nn.utils.clip_grad_norm_([p for p in m.parameters() if p.grad is not None], 0.25)
    opt.step()
    torch.cuda.synchronize()
    p_torch = {n: p.detach().clone() for n, p in m.named_parameters()}
    with torch.no_grad():                     # back to the state before the step
        for n, p in m.named_parameters():
            p.copy_(p0[n])
            if n in g0:
                p.grad.copy_(g0[n])
    eng.step(lr=1e-4, max_grad_l2_norm=0.25)
    torch.cuda.synchronize()
    for n, p in m.named_parameters():
        assert (p.detach() - p_torch[n]).abs().max().item() <= 1e-6, n

    # checkpoint round trip into a fresh model / engine
    m2 = build_b200_model(d, synth.make_state_dict(d, seed=3, variant="default"), train=True)
    eng2 = m2.train_engine()
    eng2.load_checkpoint(path, cfg)
    assert eng2.step_count == 2
    for (n, a), (_, b) in zip(m2.state_dict().items(), ck["model"].items()):
        assert torch.equal(a.cpu(), b), n
    sd2 = eng2.optimizer_state_dict(cfg)          # (torch's load_state_dict aliased `sd`'s tensors: compare with the file)
    assert set(sd2["state"]) == set(ck["optimizer"]["state"])
    for i, s in ck["optimizer"]["state"].items():
        assert torch.equal(sd2["state"][i]["exp_avg"].cpu(), s["exp_avg"].cpu())
        assert torch.equal(sd2["state"][i]["exp_avg_sq"].cpu(), s["exp_avg_sq"].cpu())


@pytest.mark.parametrize("kind", ["m4c", "t5vitevqa", "gt_box"])
def test_single_variant_training_gradients_match_oracle_autograd(kind):
    """The training step of the single-variant models (M4C, the T5-ViteVQA baseline, the GT-box upper bound): one
    answer-transformer variant straight on the encoders' output, masked BCE only; gradients against autograd through the
    oracle (whose forward is pinned to the real classes' goldens)."""
    from oracle import t2s_oracle as O
    from parity_utils import build_b200_model, sample_list
    from vitxt_gqa_b200 import synth
    d = synth.Dims(frames=8, ocr_per_frame=4, vocab=200, frame_topk=1 if kind == "m4c" else 2, ocr_topk=1 if kind == "m4c" else 3,
                   model=kind)
    sd = synth.make_state_dict(d, seed=0, variant="stress")
    inp = synth.make_inputs(d, 3, seed=13, train=True)
    sdg = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    fwd = {"m4c": O.forward_m4c, "t5vitevqa": O.forward_t5vitevqa, "gt_box": O.forward_gt_box}[kind]
    ref_out = fwd(sdg, d, inp, training=True)
    bce = O.pos_bce_loss(ref_out["pos_scores"], inp["targets"], inp["train_loss_mask"])
    bce.backward()
    ref_g = {k: v.grad for k, v in sdg.items() if v.requires_grad}
    m = build_b200_model(d, sd, train=True)
    out = m(sample_list(inp))
    assert torch.equal(out["ground_frame"].cpu(), ref_out["ground_frame"])
    assert torch.equal(out["ground_box"].cpu(), ref_out["ground_box"].detach())
    assert (out["pos_scores"].detach().cpu() - ref_out["pos_scores"].detach()).abs().max().item() <= 5e-2
    (loss,) = out["losses"].values()
    assert abs(loss.item() - bce.item()) <= 1e-2 * abs(bce.item()) + 1e-4
    loss.sum().backward()
    torch.cuda.synchronize()
    eng = m.train_engine()
    worst = {}
    for name, p in m.named_parameters():
        ref = ref_g.get(name)
        if name.startswith(eng.DEAD_PREFIXES):
            assert p.grad is None or p.grad.abs().max().item() == 0, name
            assert ref is None or ref.abs().max().item() == 0, name
            continue
        assert p.grad is not None, name
        if ref is None or ref.norm().item() == 0:
            assert p.grad.abs().max().item() <= 1e-6, name
            continue
        if name.endswith("attention.self.key.bias"):
            qref = ref_g[name.replace(".key.", ".query.")].abs().max().item()
            assert p.grad.abs().max().item() <= 2e-2 * qref + 1e-6, (name, p.grad.abs().max().item(), qref)
            continue
        worst[name] = rel_l2(p.grad.cpu(), ref)
    bad = {k: v for k, v in worst.items() if v > 5e-2}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:12]
    a = torch.cat([m.get_parameter(k).grad.flatten().cpu().double() for k in worst])
    b = torch.cat([ref_g[k].flatten().double() for k in worst])
    assert F.cosine_similarity(a, b, dim=0).item() >= 0.999
    # and the fused optimizer steps it
    eng.step(lr=1e-4, max_grad_l2_norm=0.25)
    m.eval()
    with torch.no_grad():
        m(sample_list(inp))


def test_operand_copies_are_rewritten_in_place_after_the_step():
    """After `eng.step()` the bf16 / hi|lo / transposed weight copies are refreshed by ONE kernel from a job table
    (t2s_repack_weights) instead of ~300 tensor ops: every copy must equal, bit for bit, what packing from scratch
    through torch produces from the updated parameters -- and the tensors must be the same objects (pointers stable)."""
    from parity_utils import build_b200_model, sample_list
    from vitxt_gqa_b200 import synth
    d = synth.Dims(frames=8, ocr_per_frame=4, vocab=200, frame_topk=3, ocr_topk=2)
    sd = synth.make_state_dict(d, seed=0, variant="stress")
    m = build_b200_model(d, sd, train=True)
    sl = sample_list(synth.make_inputs(d, 2, seed=12, train=True))
    eng = m.train_engine()
    for _ in range(2):
        sum(m(sl)["losses"].values()).backward()
        eng.step(lr=1e-2, max_grad_l2_norm=0.25)          # a large step: every weight moves
    torch.cuda.synchronize()
    tab = eng._repack_tab
    assert tab is not None and tab["jobs"] is not None and tab["n_jobs"] > 40, "the in-place path was not taken"
    P, W = m._packed, eng._wt
    assert P is tab["P"] and W is tab["W"]

    def flat(x, pre=""):
        out = {}
        if torch.is_tensor(x):
            out[pre] = x
        elif isinstance(x, dict):
            for k, v in x.items():
                out.update(flat(v, pre + "/" + str(k)))
        elif isinstance(x, (list, tuple)):
            for i, v in enumerate(x):
                out.update(flat(v, pre + "/%d" % i))
        return out

    got = {k: v.clone() for k, v in {**flat(P, "P"), **flat(W, "W")}.items()}
    m._packed, eng._wt = None, None
    want = {**flat(m._pack(eng.dev), "P"), **flat(eng._wt_pack(), "W")}
    assert set(want) <= set(got) and set(got) - set(want) <= {"P/q_linear"}    # (the lazily split dead weight of the grounding module)
    for k in want:
        assert got[k].shape == want[k].shape and torch.equal(got[k], want[k]), "stale or wrong operand copy: " + k
    # and a forward on the refreshed copies equals a forward on freshly packed ones
    m.eval()
    with torch.no_grad():
        a = m(sl)["pos_scores"].clone()
        m._packed = None
        b = m(sl)["pos_scores"].clone()
    assert torch.equal(a, b)
