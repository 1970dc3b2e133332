"""Per-kernel GPU tests: every entry point of libt2s_sm100 is called through the C ABI
and compared with a closed-form fp32/fp64 torch computation of the same op.

Tolerances are written beside each check: bit-level for index/mask work, 1e-5-class
for fp32 kernels (reduction order only), bf16-class (2^-8 relative on the output
rounding) for kernels that store bf16.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from vitxt_gqa_b200 import lib as tlib  # noqa: E402


@pytest.fixture(scope="module")
def L():
    return tlib.get_lib()


def stream():
    return torch.cuda.current_stream().cuda_stream


def P(t):
    return None if t is None else t.data_ptr()


def gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def rnd(*shape, scale=1.0, dtype=torch.float32, seed=None):
    g = torch.Generator(device="cuda")
    g.manual_seed(1234 if seed is None else seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(dtype)


# ------------------------------------------------------------------------------- K1 tcgen05 GEMM
@pytest.mark.parametrize("M,N,K,bn", [
    (128, 256, 64, 256), (128, 128, 128, 128), (128, 64, 64, 64),      # single tile, single/two k-blocks
    (256, 768, 768, 0), (300, 768, 768, 0), (1044, 2304, 768, 0),       # ragged M
    (777, 3072, 768, 256), (513, 768, 3072, 128), (64, 2304, 768, 64),   # FFN shapes, small M
    (768, 5000, 768, 0), (2, 768, 768, 0), (4100, 768, 1000, 256),       # ragged N (classifier), tiny M, ragged K
])
def test_gemm_bf16_plain(L, M, N, K, bn):
    A = rnd(M, K, dtype=torch.bfloat16, seed=1)
    W = rnd(N, K, scale=0.05, dtype=torch.bfloat16, seed=2)
    C = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.gemm_bf16(P(A), K, P(W), K, None, None, 0, P(C), N, M, N, K, 0, bn, stream())
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    err = (C.float() - ref).abs().max().item()
    tol = 2 ** -8 * ref.abs().max().item() + 1e-3        # one bf16 rounding of the output
    assert torch.isfinite(C.float()).all(), "unwritten or non-finite outputs"
    assert err <= tol, (err, tol)


@pytest.mark.parametrize("flags,res_kind", [(0, None), (tlib.GEMM_GELU, None), (0, "bf16"),
                                             (tlib.GEMM_OUT_F32, None), (tlib.GEMM_OUT_F32 | tlib.GEMM_RES_F32, "f32"),
                                             (tlib.GEMM_GELU | tlib.GEMM_OUT_F32, None)])
def test_gemm_bf16_epilogues(L, flags, res_kind):
    M, N, K = 1000, 1536, 768
    A = rnd(M, K, dtype=torch.bfloat16, seed=3)
    W = rnd(N, K, scale=0.05, dtype=torch.bfloat16, seed=4)
    bias = rnd(N, seed=5)
    res = None
    if res_kind == "bf16":
        res = rnd(M, N, dtype=torch.bfloat16, seed=6)
    elif res_kind == "f32":
        res = rnd(M, N, seed=6)
    out_f32 = bool(flags & tlib.GEMM_OUT_F32)
    C = torch.zeros(M, N, device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    L.gemm_bf16(P(A), K, P(W), K, P(bias), P(res), N, P(C), N, M, N, K, flags, 0, stream())
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias
    if flags & tlib.GEMM_GELU:
        ref = gelu(ref)
    if res is not None:
        ref = ref + res.float()
    err = (C.float() - ref).abs().max().item()
    tol = (1e-3 if out_f32 else 2 ** -8 * ref.abs().max().item() + 1e-3)   # fp32 out: accumulation order only
    assert err <= tol, (err, tol)


@pytest.mark.parametrize("M,cap", [(5000, 0), (5000, 37), (3 * 128, 2), (66816, 124)])
def test_gemm_bf16_cta_pairs(L, M, cap):
    """The 128 x 256 throughput tile runs as 2-CTA clusters (W tile halves multicast into both CTAs): odd numbers of
    row blocks (the second CTA of the last pair gets an out-of-range tile), odd / tiny SM caps, many tiles per pair."""
    N, K = 2304, 768
    A = rnd(M, K, dtype=torch.bfloat16, seed=31)
    W = rnd(N, K, scale=0.05, dtype=torch.bfloat16, seed=32)
    bias = rnd(N, seed=33)
    R = rnd(M, N, dtype=torch.bfloat16, seed=34)
    C = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.gemm_bf16(P(A), K, P(W), K, P(bias), P(R), N, P(C), N, M, N, K, cap << tlib.GEMM_SM_CAP_SHIFT, 256, stream())
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias + R.float()
    assert torch.isfinite(C.float()).all(), "unwritten or non-finite outputs"
    err = (C.float() - ref).abs().max().item()
    assert err <= 2 ** -8 * ref.abs().max().item() + 1e-3, err


def test_gemm_bf16_strided_rows(L):
    """Decoder-step addressing: one row per sample, T rows apart, fp32 output with a wide pitch."""
    B, T, H, V, N = 64, 12, 768, 5000, 5960
    X = rnd(B * T, H, dtype=torch.bfloat16, seed=7)
    W = rnd(V, H, scale=0.05, dtype=torch.bfloat16, seed=8)
    bias = rnd(V, seed=9)
    S = torch.zeros(B, T, N, device="cuda")
    t0 = 5
    L.gemm_bf16(X.data_ptr() + t0 * H * 2, T * H, P(W), H, P(bias), None, 0, S.data_ptr() + t0 * N * 4, T * N,
                B, V, H, tlib.GEMM_OUT_F32, 0, stream())
    torch.cuda.synchronize()
    ref = X.view(B, T, H)[:, t0].float() @ W.float().t() + bias
    assert (S[:, t0, :V] - ref).abs().max().item() <= 1e-3
    assert S[:, t0, V:].abs().max().item() == 0 and S[:, :t0].abs().max().item() == 0


# ------------------------------------------------------------------------------- K1x bf16x3 GEMM
def _split(x, kp=None):
    k = x.shape[1]
    kp = k if kp is None else kp
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    out = torch.zeros(x.shape[0], 2 * kp, device=x.device, dtype=torch.bfloat16)
    out[:, :k], out[:, kp:kp + k] = hi, lo
    return out


@pytest.mark.parametrize("M,N,K,flags", [
    (1280, 2304, 768, tlib.GEMM_OUT_F32), (1044, 768, 768, tlib.GEMM_OUT_F32 | tlib.GEMM_RES_F32),
    (777, 3072, 768, tlib.GEMM_GELU | tlib.GEMM_OUT_SPLIT), (300, 768, 3072, tlib.GEMM_OUT_F32 | tlib.GEMM_RES_F32),
    (64, 768, 1088, tlib.GEMM_OUT_F32), (4100, 768, 1024, tlib.GEMM_OUT_F32)])
def test_gemm_bf16x3(L, M, N, K, flags):
    """fp32-class accuracy from three bf16 tensor-core products (hi.hi + hi.lo + lo.hi)."""
    A, W, bias = rnd(M, K, seed=21), rnd(N, K, scale=0.05, seed=22), rnd(N, seed=23)
    res = rnd(M, N, seed=24) if flags & tlib.GEMM_RES_F32 else None
    As, Ws = _split(A), _split(W)
    split_out = bool(flags & tlib.GEMM_OUT_SPLIT)
    C = torch.zeros(M, 2 * N if split_out else N, device="cuda", dtype=torch.bfloat16 if split_out else torch.float32)
    L.gemm_bf16x3(P(As), 2 * K, P(Ws), 2 * K, P(bias), P(res), N, P(C), C.shape[1], M, N, K, flags, 0, stream())
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t() + bias.double()
    if flags & tlib.GEMM_GELU:
        ref = gelu(ref)
    if res is not None:
        ref = ref + res.double()
    got = (C[:, :N].float() + C[:, N:].float()) if split_out else C
    err = (got.double() - ref).abs().max().item()
    scale = (A.abs().double() @ W.abs().double().t()).max().item()      # sum |a||w|: what the 2^-17 applies to
    assert err <= 2 ** -15 * scale + 1e-6, (err, scale)
    # and it really is fp32-class: a plain bf16 contraction of the same operands is >30x worse
    bf = (A.to(torch.bfloat16).double() @ W.to(torch.bfloat16).double().t() + bias.double())
    if not (flags & tlib.GEMM_GELU) and res is None:
        assert err * 30 < (bf - ref).abs().max().item()


def test_split_bf16(L):
    rows, K, ldx, lo_off = 1000, 1074, 1076, 1088       # ragged K, 16-byte aligned row pitch
    x = rnd(rows, ldx, seed=31)[:, :K]
    out = torch.full((rows, 2 * lo_off), 7.0, device="cuda", dtype=torch.bfloat16)
    L.split_bf16(P(x), ldx, rows, K, lo_off, P(out), 2 * lo_off, 0, 0, 0, stream())
    torch.cuda.synchronize()
    assert torch.equal(out[:, :K], x.to(torch.bfloat16))
    assert torch.equal(out[:, lo_off:lo_off + K], (x - x.to(torch.bfloat16).float()).to(torch.bfloat16))
    assert out[:, K:lo_off].abs().max().item() == 0 and out[:, lo_off + K:].abs().max().item() == 0


def test_add_ln_split(L):
    rows, H = 999, 768
    x, r = rnd(rows, H, seed=32), rnd(rows, H, seed=33)
    g, b = 1 + 0.1 * rnd(H, seed=34), 0.1 * rnd(H, seed=35)
    o32 = torch.zeros(rows, H, device="cuda")
    o16 = torch.zeros(rows, 2 * H, device="cuda", dtype=torch.bfloat16)
    L.add_ln_split(P(x), 0, H, P(r), 0, H, P(g), P(b), 1e-12, rows, H, None, 0, P(o32), H, P(o16), 2 * H, 0, 0, 0, stream())
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm((x + r).double(), (H,), g.double(), b.double(), 1e-12).float()
    assert (o32 - ref).abs().max().item() <= 1e-5
    assert torch.equal(o16[:, :H], o32.to(torch.bfloat16))
    assert torch.equal(o16[:, H:], (o32 - o32.to(torch.bfloat16).float()).to(torch.bfloat16))


# ------------------------------------------------------------------------------- K1f fp32 GEMM
@pytest.mark.parametrize("M,N,K", [(1280, 2304, 768), (100, 768, 3072), (4097, 768, 1088), (1, 768, 768), (960, 768, 16)])
def test_gemm_f32(L, M, N, K):
    A, W, bias, res = rnd(M, K, seed=1), rnd(N, K, scale=0.05, seed=2), rnd(N, seed=3), rnd(M, N, seed=4)
    C = torch.zeros(M, N, device="cuda")
    L.gemm_f32(P(A), K, P(W), K, P(bias), P(res), N, P(C), N, M, N, K, tlib.GEMM_GELU, 0, 0, 0, stream())
    torch.cuda.synchronize()
    ref = (gelu(A.double() @ W.double().t() + bias.double()) + res.double()).float()
    assert (C - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())   # fp32 accumulation over K


def test_gemm_f32_row_gather(L):
    B, Le, Lt, H = 5, 52, 20, 768
    J = rnd(B * Le, H, seed=5)
    W, bias = rnd(H, H, scale=0.05, seed=6), rnd(H, seed=7)
    C = torch.zeros(B * Lt, H, device="cuda")
    L.gemm_f32(P(J), H, P(W), H, P(bias), None, 0, P(C), H, B * Lt, H, H, 0, Lt, Le, 0, stream())
    torch.cuda.synchronize()
    ref = (J.view(B, Le, H)[:, :Lt].reshape(-1, H).double() @ W.double().t() + bias.double()).float()
    assert (C - ref).abs().max().item() <= 2e-4


# ------------------------------------------------------------------------------- K2 attention
def _attn_ref(qkv, B, L, H, keys, nk):
    q, k, v = qkv.view(B, L, 3, 12, 64).double().unbind(2)
    out = torch.zeros(B, L, 12, 64, dtype=torch.float64, device=qkv.device)
    for b in range(B):
        idx = keys[b, :nk[b]].long()
        s = torch.einsum("qhd,khd->hqk", q[b], k[b, idx]) / 8.0
        out[b] = torch.einsum("hqk,khd->qhd", s.softmax(-1), v[b, idx])
    return out.reshape(B * L, H)


def _keys(B, L, seed=0, frac=0.6):
    g = torch.Generator().manual_seed(seed)
    mask = (torch.rand(B, L, generator=g) < frac)
    mask[:, 0] = True
    keys = torch.zeros(B, L, dtype=torch.int32)
    nk = torch.zeros(B, dtype=torch.int32)
    for b in range(B):
        idx = mask[b].nonzero()[:, 0]
        keys[b, :len(idx)] = idx.int()
        nk[b] = len(idx)
    return mask.float().cuda(), keys.cuda(), nk.cuda()


@pytest.mark.parametrize("B,L", [(3, 20), (2, 52), (2, 200), (1, 1044)])
def test_attn_f32(B, L):
    lib = tlib.get_lib()
    H = 768
    qkv = rnd(B * L, 3 * H, seed=11)
    mask, keys, nk = _keys(B, L, seed=B + L)
    out = torch.zeros(B * L, H, device="cuda")
    osp = torch.zeros(B * L, 2 * H, device="cuda", dtype=torch.bfloat16)
    lib.attn_f32(P(qkv), 3 * H, B, L, H, 12, P(keys), P(nk), L, P(out), H, P(osp), 2 * H, stream())
    torch.cuda.synchronize()
    ref = _attn_ref(qkv, B, L, H, keys, nk).float()
    assert (out - ref).abs().max().item() <= 2e-5        # fp32, softmax-normalised outputs of O(1)
    # bf16 hi|lo copy: hi is the bf16 rounding, hi + lo reproduces the fp32 value to 2^-17 relative
    assert torch.equal(osp[:, :H], out.to(torch.bfloat16))
    rec = osp[:, :H].float() + osp[:, H:].float()
    assert (rec - out).abs().max().item() <= 2 ** -16 * out.abs().max().item()


@pytest.mark.parametrize("B,L", [(3, 20), (2, 52), (2, 200), (1, 1044)])
def test_attn_x3(B, L):
    """fp32-class attention from bf16 hi|lo operands: three tensor-core products per contraction."""
    lib = tlib.get_lib()
    H = 768
    qkv = rnd(B * L, 3 * H, seed=15)
    qkv[:, :2 * H] *= 1.7                      # |q.k|/8 up to ~15: exercises the exponent range of real runs
    mask, keys, nk = _keys(B, L, seed=3 * B + L)
    qs = _split(qkv)                           # [rows, 2 * 3H], lo at column 3H
    out = torch.zeros(B * L, 2 * H, device="cuda", dtype=torch.bfloat16)
    lib.attn_x3(P(qs), 6 * H, 3 * H, B, L, H, 12, P(keys), P(nk), L, P(out), 2 * H, stream())
    torch.cuda.synchronize()
    ref = _attn_ref(qkv, B, L, H, keys, nk)
    got = out[:, :H].double() + out[:, H:].double()
    err = (got - ref).abs().max().item()
    # scores here reach |q.k|/8 ~ 15 with |q_i k_i| ~ 3: a 2^-17-relative product error moves the exponent by
    # ~5e-5, hence the probabilities by the same relative amount (real runs have |q_i k_i| ~ 0.2: ~3e-6).
    # bf16 attention on the same data is ~1e-2.
    assert err <= 1.5e-4, err


@pytest.mark.parametrize("B,L", [(2, 64), (2, 52), (3, 200), (1, 1044)])
def test_attn_bf16(B, L):
    lib = tlib.get_lib()
    H = 768
    qkv = rnd(B * L, 3 * H, dtype=torch.bfloat16, seed=12)
    mask, keys, nk = _keys(B, L, seed=B * 7 + L)
    out = torch.zeros(B * L, H, device="cuda", dtype=torch.bfloat16)
    lib.attn_bf16(P(qkv), 3 * H, B, L, H, 12, P(keys), P(nk), L, P(out), H, stream())
    torch.cuda.synchronize()
    ref = _attn_ref(qkv.float(), B, L, H, keys, nk).float()
    # P is rounded to bf16 before P.V and the output is stored in bf16: ~2^-8 relative each
    assert (out.float() - ref).abs().max().item() <= 3e-2, (out.float() - ref).abs().max().item()
    assert (out.float() - ref).abs().mean().item() <= 3e-3


@pytest.mark.parametrize("B,L,frac", [(2, 64, 0.6), (2, 52, 0.6), (3, 200, 0.6), (1, 1044, 0.6), (2, 300, 1.0),
                                      (2, 1044, 0.05)])
def test_attn_tc_bf16(B, L, frac):
    """tcgen05 attention, bf16 operands: S and O accumulate in TMEM, P goes through shared memory as bf16."""
    lib = tlib.get_lib()
    H = 768
    qkv = rnd(B * L, 3 * H, dtype=torch.bfloat16, seed=16)
    mask, keys, nk = _keys(B, L, seed=B * 5 + L, frac=frac)
    out = torch.full((B * L, H), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib.attn_tc(P(qkv), 3 * H, 0, B, L, H, 12, P(keys), P(nk), L, P(out), H, stream())
    torch.cuda.synchronize()
    ref = _attn_ref(qkv.float(), B, L, H, keys, nk).float()
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs()
    assert err.max().item() <= 3e-2, err.max().item()
    assert err.mean().item() <= 3e-3


def test_attn_tc_lazy_rescale():
    """Keys ordered so that the row maximum keeps rising by more than 2^8 from tile to tile: exercises the
    in-TMEM rescale of O."""
    lib = tlib.get_lib()
    B, L, H = 1, 512, 768
    qkv = rnd(B * L, 3 * H, dtype=torch.bfloat16, seed=17)
    ramp = torch.linspace(0.2, 3.0, L, device="cuda")[:, None]
    qkv[:, H:2 * H] = (qkv[:, H:2 * H].float() * ramp).to(torch.bfloat16)        # later keys score much higher
    qkv[:, :H] = (qkv[:, :H].float().abs() * 1.5).to(torch.bfloat16)
    qkv[:, H:2 * H] = qkv[:, H:2 * H].float().abs().to(torch.bfloat16) * torch.sign(ramp - 0.1).to(torch.bfloat16)
    keys = torch.arange(L, dtype=torch.int32, device="cuda")[None].contiguous()
    nk = torch.tensor([L], dtype=torch.int32, device="cuda")
    out = torch.zeros(B * L, H, device="cuda", dtype=torch.bfloat16)
    lib.attn_tc(P(qkv), 3 * H, 0, B, L, H, 12, P(keys), P(nk), L, P(out), H, stream())
    torch.cuda.synchronize()
    ref = _attn_ref(qkv.float(), B, L, H, keys, nk).float()
    assert (out.float() - ref).abs().max().item() <= 4e-2


@pytest.mark.parametrize("B,L", [(3, 20), (2, 52), (2, 200), (1, 1044)])
def test_attn_tc_x3(B, L):
    """tcgen05 attention with bf16 hi|lo operands: fp32-class."""
    lib = tlib.get_lib()
    H = 768
    qkv = rnd(B * L, 3 * H, seed=18)
    mask, keys, nk = _keys(B, L, seed=3 * B + L)
    qs = _split(qkv)
    out = torch.zeros(B * L, 2 * H, device="cuda", dtype=torch.bfloat16)
    lib.attn_tc(P(qs), 6 * H, 3 * H, B, L, H, 12, P(keys), P(nk), L, P(out), 2 * H, stream())
    torch.cuda.synchronize()
    ref = _attn_ref(qkv, B, L, H, keys, nk)
    got = out[:, :H].double() + out[:, H:].double()
    err = (got - ref).abs().max().item()
    assert err <= 5e-5, err


@pytest.mark.parametrize("t0,nq", [(0, 1), (5, 1), (11, 1), (0, 12), (2, 4), (3, 7), (0, 3)])
def test_attn_dec(t0, nq):
    lib = tlib.get_lib()
    B, Le, T, H = 3, 116, 12, 768
    enc = rnd(B * Le, 3 * H, dtype=torch.bfloat16, seed=13)
    dec = rnd(B * T, 3 * H, dtype=torch.bfloat16, seed=14)
    mask, keys, nk = _keys(B, Le, seed=99)
    out = torch.zeros(B * T, H, device="cuda", dtype=torch.bfloat16)
    lib.attn_dec(P(enc), 3 * H, Le, P(dec), 3 * H, T, B, H, 12, P(keys), P(nk), Le, t0, nq, P(out), H, stream())
    torch.cuda.synchronize()
    qe, ke, ve = enc.view(B, Le, 3, 12, 64).double().unbind(2)
    qd, kd, vd = dec.view(B, T, 3, 12, 64).double().unbind(2)
    for b in range(B):
        idx = keys[b, :nk[b]].long()
        for i in range(nq):
            t = t0 + i
            k = torch.cat([ke[b, idx], kd[b, :t + 1]], 0)
            v = torch.cat([ve[b, idx], vd[b, :t + 1]], 0)
            s = torch.einsum("hd,khd->hk", qd[b, t], k) / 8.0
            ref = torch.einsum("hk,khd->hd", s.softmax(-1), v).reshape(H).float()
            got = out.view(B, T, H)[b, t].float()
            assert (got - ref).abs().max().item() <= 2e-2, (b, t, (got - ref).abs().max().item())
    if nq == 1:     # other rows untouched
        other = [t for t in range(T) if t != t0]
        assert out.view(B, T, H)[:, other].abs().max().item() == 0


# ------------------------------------------------------------------------------- K3 / K4
def test_bert_embed_ln(L):
    B, Lt, H = 7, 20, 768
    ids = torch.randint(0, 30522, (B, Lt), device="cuda")
    word, pos, typ = rnd(30522, H, seed=1), rnd(512, H, seed=2), rnd(2, H, seed=3)
    g, b = 1 + 0.1 * rnd(H, seed=4), 0.1 * rnd(H, seed=5)
    out = torch.zeros(B * Lt, H, device="cuda")
    L.bert_embed_ln(P(ids), B * Lt, Lt, H, P(word), P(pos), P(typ), P(g), P(b), 1e-12, P(out), H, stream())
    torch.cuda.synchronize()
    e = word[ids] + pos[torch.arange(Lt, device="cuda")][None] + typ[0]
    ref = torch.nn.functional.layer_norm(e, (H,), g, b, 1e-12).view(-1, H)
    assert (out - ref).abs().max().item() <= 1e-5


def test_feat_concat(L):
    rows, kp = 333, 1008
    f0, f1 = rnd(rows, 300, seed=1), (rnd(rows, 604, seed=2) > 1.5).float()
    f1[5] = 0          # zero row stays zero (Q20)
    id0 = torch.randint(0, 4000, (rows,), device="cuda")
    id1 = torch.randint(0, 4000, (rows,), device="cuda")
    t0, t1 = rnd(4000, 50, seed=3), rnd(4000, 50, seed=4)
    out = torch.full((rows, kp), float("nan"), device="cuda")
    osp = torch.full((rows, 2 * kp), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.feat_concat(P(f0), 300, P(f1), 604, P(id0), P(t0), P(id1), P(t1), 50, rows, P(out), kp, kp, P(osp), 2 * kp, stream())
    torch.cuda.synchronize()
    F = torch.nn.functional
    ref = torch.cat([F.normalize(f0, dim=-1), F.normalize(f1, dim=-1), t0[id0], t1[id1],
                     torch.zeros(rows, kp - 1004, device="cuda")], -1)
    assert (out - ref).abs().max().item() <= 1e-6
    assert out[5, 300:904].abs().max().item() == 0
    assert torch.equal(osp, _split(out))          # fused bf16 hi|lo copy == split of the fp32 rows


def test_split_bf16_row_gather(L):
    B, Le, Lt, H = 5, 52, 20, 768
    J = rnd(B * Le, H, seed=41)
    out = torch.zeros(B * Lt, 2 * H, device="cuda", dtype=torch.bfloat16)
    L.split_bf16(P(J), H, B * Lt, H, H, P(out), 2 * H, Lt, Le, 0, stream())
    torch.cuda.synchronize()
    assert torch.equal(out, _split(J.view(B, Le, H)[:, :Lt].reshape(-1, H)))


@pytest.mark.parametrize("x_bf16,res_kind,tanh,remap", [(0, None, False, False), (0, "f32", True, False),
                                                         (1, None, False, False), (1, "bf16", False, True),
                                                         (0, "f32", False, True)])
def test_add_ln(L, x_bf16, res_kind, tanh, remap):
    rows, H, per, group, off = 120, 768, 20, 52, 7
    x = rnd(rows, H, seed=1, dtype=torch.bfloat16 if x_bf16 else torch.float32)
    res = None if res_kind is None else rnd(rows, H, seed=2, dtype=torch.bfloat16 if res_kind == "bf16" else torch.float32)
    g, b = 1 + 0.1 * rnd(H, seed=3), 0.1 * rnd(H, seed=4)
    out_rows = (rows // per) * group if remap else rows
    base = rnd(out_rows, H, seed=5) if tanh else None
    o32 = torch.zeros(out_rows, H, device="cuda")
    o16 = torch.zeros(out_rows, H, device="cuda", dtype=torch.bfloat16)
    L.add_ln(P(x), x_bf16, H, P(res), int(res_kind == "bf16"), H, P(g), P(b), 1e-12, rows, H, P(base), H,
             P(o32), H, P(o16), H, per if remap else 0, group, off, stream())
    torch.cuda.synchronize()
    v = x.float() + (res.float() if res is not None else 0)
    ref = torch.nn.functional.layer_norm(v, (H,), g, b, 1e-12)
    r = torch.arange(rows, device="cuda")
    orow = (r // per) * group + off + r % per if remap else r
    if tanh:
        ref = base[orow] + torch.tanh(ref)
    assert (o32[orow] - ref).abs().max().item() <= 2e-5
    assert (o16[orow].float() - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item() + 1e-3


def test_ocr_finish(L):
    rows, H = 96, 768
    h, bbox = rnd(rows, H, seed=1), torch.rand(rows, 4, device="cuda")
    w2, b2 = rnd(H, 4, seed=2), rnd(H, seed=3)
    g1, be1, g2, be2 = (1 + 0.1 * rnd(H, seed=4), 0.1 * rnd(H, seed=5), 1 + 0.1 * rnd(H, seed=6), 0.1 * rnd(H, seed=7))
    out = torch.zeros(2 * 60, H, device="cuda")
    L.ocr_finish(P(h), H, P(bbox), P(w2), P(b2), P(g1), P(be1), P(g2), P(be2), 1e-5, rows, H, P(out), H, 48, 60, 12, stream())
    torch.cuda.synchronize()
    F = torch.nn.functional
    ref = F.layer_norm(h, (H,), g1, be1, 1e-5) + F.layer_norm(bbox @ w2.t() + b2, (H,), g2, be2, 1e-5)
    assert (out.view(2, 60, H)[:, 12:].reshape(-1, H) - ref).abs().max().item() <= 2e-5


def test_prev_embed(L):
    B, T, V, O, H, Le = 4, 12, 200, 32, 768, 60
    prev = torch.randint(0, V + O, (B, T), device="cuda")
    prev[0, 0], prev[1, 1] = 1, V          # first fixed-vocab and first OCR index
    ans = rnd(V, H, seed=1)
    J = rnd(B * Le, H, seed=2)
    ocr_row0 = Le - O
    pos, typ = rnd(100, H, seed=3), rnd(5, H, seed=4)
    ln = [1 + 0.1 * rnd(H, seed=10 + i) if i % 2 == 0 else 0.1 * rnd(H, seed=10 + i) for i in range(6)]
    out = torch.zeros(B * T, H, device="cuda", dtype=torch.bfloat16)
    o32 = torch.zeros(B * T, H, device="cuda")
    L.prev_embed(P(prev), T, B, 0, T, T, V, H, P(ans), J.data_ptr() + ocr_row0 * H * 4, Le * H, H, P(pos), P(typ),
                 P(ln[0]), P(ln[1]), P(ln[2]), P(ln[3]), P(ln[4]), P(ln[5]), 1e-12, P(out), P(o32), H, O, stream())
    torch.cuda.synchronize()
    F = torch.nn.functional
    ocr = J.view(B, Le, H)[:, ocr_row0:]
    cat = torch.cat([F.layer_norm(ans, (H,), ln[0], ln[1], 1e-12)[None].expand(B, -1, -1),
                     F.layer_norm(ocr, (H,), ln[2], ln[3], 1e-12)], 1)
    raw = torch.gather(cat, 1, prev[:, :, None].expand(B, T, H))
    emb = F.layer_norm(pos[:T][None] + typ[(prev >= V).long()], (H,), ln[4], ln[5], 1e-12)
    ref = (raw + emb).view(-1, H)
    assert (o32 - ref).abs().max().item() <= 2e-5
    assert (out.float() - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item() + 1e-3


def test_mask_prep_build_keys(L):
    B, Lt, F, O = 5, 20, 8, 32
    tl = torch.randint(1, Lt + 1, (B,), device="cuda")
    fm = (torch.rand(B, F, device="cuda") < 0.7).long()
    om = (torch.rand(B, O, device="cuda") < 0.4).long()
    Le = Lt + F + O
    jm = torch.zeros(B, Le, device="cuda")
    keys = torch.full((B, Le), -1, device="cuda", dtype=torch.int32)
    nk = torch.zeros(B, device="cuda", dtype=torch.int32)
    L.mask_prep(P(tl), P(fm), P(om), B, Lt, F, O, P(jm), stream())
    L.build_keys(P(jm), B, Le, P(keys), P(nk), Le, stream())
    torch.cuda.synchronize()
    ref = torch.cat([(torch.arange(Lt, device="cuda")[None] < tl[:, None]).float(), fm.float(), om.float()], 1)
    assert torch.equal(jm, ref)
    for b in range(B):
        idx = ref[b].nonzero()[:, 0].int()
        assert nk[b].item() == len(idx) and torch.equal(keys[b, :len(idx)], idx)


# ------------------------------------------------------------------------------- K6 / K7
def test_ptr_score_and_argmax(L):
    B, T, V, O, H, Le = 3, 12, 200, 64, 768, 100
    N = V + O
    q = rnd(B * T, H, dtype=torch.bfloat16, seed=1)
    keyp = rnd(B * Le, H, dtype=torch.bfloat16, seed=2)
    jm = (torch.rand(B, Le, device="cuda") < 0.5).float()
    off = Le - O
    S = rnd(B, T, N, seed=3).contiguous()
    S0 = S.clone()
    L.ptr_score(P(q), H, B, T, 0, T, keyp.data_ptr() + off * H * 2, Le * H, H, O, H, jm.data_ptr() + off * 4, Le,
                P(S), N, V, stream())
    torch.cuda.synchronize()
    ref = torch.einsum("btd,bod->bto", q.view(B, T, H).float(), keyp.view(B, Le, H)[:, off:].float()) / math.sqrt(H) \
        + jm[:, None, off:]
    assert (S[:, :, V:] - ref).abs().max().item() <= 1e-4
    assert torch.equal(S[:, :, :V], S0[:, :, :V])
    prev = torch.zeros(B, T, device="cuda", dtype=torch.int64)
    am = torch.zeros(B, T, device="cuda", dtype=torch.int64)
    S[0, 3, 17] = S[0, 3, 150] = 1e6          # tie: first index wins
    L.argmax_feedback(P(S), N, B, T, 0, T, N, P(prev), T, P(am), stream())
    torch.cuda.synchronize()
    assert torch.equal(am, S.argmax(-1)) and am[0, 3].item() == 17
    assert torch.equal(prev[:, 1:], am[:, :-1]) and prev[:, 0].abs().max().item() == 0


@pytest.mark.parametrize("O,t0,nq", [(960, 0, 12), (50, 2, 4), (77, 0, 16), (960, 5, 2)])
def test_ptr_score_several_rows_tensor_core_kernel(L, O, t0, nq):
    """nq > 1 decoder rows go through the mma.sync kernel (keys = M, queries = N), one row through the scalar kernel:
    both against torch, ragged key counts, row offsets, and untouched neighbours."""
    B, T, V, H = 3, 16, 40, 768
    Le, N = 20 + O, V + O
    q = rnd(B * T, H, dtype=torch.bfloat16, seed=1)
    keyp = rnd(B * Le, H, dtype=torch.bfloat16, seed=2)
    jm = (torch.rand(B, Le, device="cuda") < 0.5).float()
    off = Le - O
    S = torch.full((B, T, N), float("nan"), device="cuda")
    L.ptr_score(P(q), H, B, T, t0, nq, keyp.data_ptr() + off * H * 2, Le * H, H, O, H, jm.data_ptr() + off * 4, Le,
                P(S), N, V, stream())
    torch.cuda.synchronize()
    ref = torch.einsum("btd,bod->bto", q.view(B, T, H).float(), keyp.view(B, Le, H)[:, off:].float()) / math.sqrt(H) \
        + jm[:, None, off:]
    assert (S[:, t0:t0 + nq, V:] - ref[:, t0:t0 + nq]).abs().max().item() <= 1e-4
    assert torch.isnan(S[:, :, :V]).all() and torch.isnan(S[:, :t0]).all() and torch.isnan(S[:, t0 + nq:]).all()
    one = torch.full((B, T, N), float("nan"), device="cuda")
    for t in range(t0, t0 + nq):      # the one-row (greedy) kernel gives the same numbers to fp32 summation order
        L.ptr_score(P(q), H, B, T, t, 1, keyp.data_ptr() + off * H * 2, Le * H, H, O, H, jm.data_ptr() + off * 4, Le,
                    P(one), N, V, stream())
    torch.cuda.synchronize()
    assert (one[:, t0:t0 + nq, V:] - S[:, t0:t0 + nq, V:]).abs().max().item() <= 2e-5


def test_losses(L):
    B, T, N = 5, 12, 5960
    ref_s, pos_s, neg_s = rnd(B, T, N, seed=1), rnd(B, T, N, seed=2), rnd(B, T, N, seed=3)
    tg = (torch.rand(B, T, N, device="cuda") < 0.001).float()
    lm = (torch.rand(B, T, device="cuda") < 0.5).float()
    ws = torch.empty(int(L.loss_workspace_bytes(B, T)), device="cuda", dtype=torch.uint8)
    o1, o2 = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    L.pos_bce_loss(P(pos_s), P(tg), P(lm), B, T, N, P(ws), P(o1), stream())
    L.info_nce_loss(P(ref_s), P(pos_s), P(neg_s), B, T, N, 0.1, P(ws), P(o2), stream())
    torch.cuda.synchronize()
    F = torch.nn.functional
    bce = (F.binary_cross_entropy_with_logits(pos_s.double(), tg.double(), reduction="none") * lm[:, :, None]).sum() \
        / max(lm.sum().item(), 1.0)
    q, p, n = (F.normalize(x.double(), dim=-1).view(B, -1) for x in (ref_s, pos_s, neg_s))
    logits = torch.stack([F.cosine_similarity(q, p, dim=1), F.cosine_similarity(q, n, dim=1)], 1) / 0.1
    nce = F.cross_entropy(logits, torch.zeros(B, dtype=torch.long, device="cuda"))
    assert abs(o1.item() - bce.item()) <= 1e-5 * abs(bce.item())      # rtol: fp32 elementwise, fp64 reduction
    assert abs(o2.item() - nce.item()) <= 1e-5
