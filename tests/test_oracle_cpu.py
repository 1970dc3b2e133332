"""Pins the CPU oracle (oracle/t2s_oracle.py) against golden outputs of the REAL reference model
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference in the dev
container).  Runs without a GPU.

fp32 bar: the oracle follows the reference's op order, so logits agree to reduction-order noise
(1.9e-6 floor); index outputs are exact.  torch.topk's order among tied -10000 entries is
implementation defined (SURVEY hard part 3), so the reference's own neg-frame choice (captured in
the fixture) is injected where it matters.
"""
import numpy as np
import pytest
import torch

from parity_utils import load_golden

from oracle import t2s_oracle as O

FP32_ATOL = 2e-4     # logits have std ~0.6; observed agreement is ~1e-5


def _neg_override(z):
    return torch.from_numpy(z["neg_frame_topk_mask"]) if "neg_frame_topk_mask" in z.files else None


@pytest.mark.parametrize("fixture,schedule", [("t2s_small_eval", "literal"), ("t2s_small_eval", "dedup"),
                                              ("t2s_small_default", "literal"), ("t2s_small_train", "literal"),
                                              # ablation models (reference models/t2s_wo_sg.py, t2s_wo_tg.py)
                                              ("t2s_wo_sg_small_eval", "dedup"), ("t2s_wo_sg_small_train", "literal"),
                                              ("t2s_wo_tg_small_eval", "dedup"), ("t2s_wo_tg_all_eval", "dedup"),
                                              # BASELINE configs[4]: 128 frames x 15 OCR slots, batch 1
                                              ("t2s_stress_f128_eval", "dedup"),
                                              ("t2s_stress_f256_eval", "dedup")])
def test_oracle_t2s_matches_reference_golden(fixture, schedule):
    z, meta, d, sd, inp = load_golden(fixture)
    train = meta["mode"] == "train"
    with torch.no_grad():
        out = O.forward_t2s(sd, d, inp, training=train, schedule=schedule, neg_frame_override=_neg_override(z),
                            return_debug=True)
    assert np.array_equal(out["ground_frame"].numpy(), z["ground_frame"])
    assert np.array_equal(out["ground_box"].numpy(), z["ground_box"])
    if "pos_frame_topk_mask" in z.files:          # the w/o TG ablation has no temporal indicator
        assert np.array_equal(out["debug"]["frame_pos_topk"].numpy(), z["pos_frame_topk_mask"])
    for k in ("ref_scores", "pos_scores", "neg_scores"):
        err = np.abs(out[k].numpy() - z[k]).max()
        assert err <= FP32_ATOL, (fixture, k, err)
        if not train:
            assert np.array_equal(out[k].numpy().argmax(-1), z[k].argmax(-1)), (fixture, k)
    tg, lm = inp["targets"], inp["train_loss_mask"]
    bce = O.pos_bce_loss(out["pos_scores"], tg, lm)
    nce = O.info_nce(out["ref_scores"], out["pos_scores"], out["neg_scores"])
    assert abs(float(bce) - float(z["loss_pos_bce"][0])) <= 1e-5 * max(1.0, abs(float(z["loss_pos_bce"][0])))
    assert abs(float(nce) - float(z["loss_info_nce"])) <= 1e-4


def test_oracle_m4c_matches_reference_golden():
    z, meta, d, sd, inp = load_golden("m4c_small_eval")
    with torch.no_grad():
        out = O.forward_m4c(sd, d, inp)
    assert np.array_equal(out["ground_frame"].numpy(), z["ground_frame"])
    assert np.array_equal(out["ground_box"].numpy(), z["ground_box"])
    err = np.abs(out["pos_scores"].numpy() - z["pos_scores"]).max()
    assert err <= FP32_ATOL, err
    assert np.array_equal(out["pos_scores"].numpy().argmax(-1), z["pos_scores"].argmax(-1))
    bce = O.pos_bce_loss(out["pos_scores"], inp["targets"], inp["train_loss_mask"])
    assert abs(float(bce) - float(z["loss_pos_bce"][0])) <= 1e-5 * max(1.0, abs(float(z["loss_pos_bce"][0])))


@pytest.mark.parametrize("fixture", ["t5vitevqa_small_eval", "t5vitevqa_small_train"])
def test_oracle_t5vitevqa_matches_reference_golden(fixture):
    """The T5-ViteVQA baseline (reference models/t5vitevqa.py), golden from the real class."""
    z, meta, d, sd, inp = load_golden(fixture)
    with torch.no_grad():
        out = O.forward_t5vitevqa(sd, d, inp, training=meta["mode"] == "train")
    assert np.array_equal(out["ground_frame"].numpy(), z["ground_frame"])
    assert np.array_equal(out["ground_box"].numpy(), z["ground_box"])
    err = np.abs(out["pos_scores"].numpy() - z["pos_scores"]).max()
    assert err <= FP32_ATOL, err
    if meta["mode"] != "train":
        assert np.array_equal(out["pos_scores"].numpy().argmax(-1), z["pos_scores"].argmax(-1))
    bce = O.pos_bce_loss(out["pos_scores"], inp["targets"], inp["train_loss_mask"])
    assert abs(float(bce) - float(z["loss_pos_bce"][0])) <= 1e-5 * max(1.0, abs(float(z["loss_pos_bce"][0])))


@pytest.mark.parametrize("fixture", ["gt_box_small_eval", "gt_box_small_train"])
def test_oracle_gt_box_matches_reference_golden(fixture):
    """The GT-box upper bound (reference models/gt_box.py), golden from the real class."""
    z, meta, d, sd, inp = load_golden(fixture)
    with torch.no_grad():
        out = O.forward_gt_box(sd, d, inp, training=meta["mode"] == "train")
    assert np.array_equal(out["ground_frame"].numpy(), z["ground_frame"])
    assert np.array_equal(out["ground_box"].numpy(), z["ground_box"])
    assert int(out["frame_topk"]) == int(z["frame_topk"]) == 64 and int(out["ocr_topk"]) == int(z["ocr_topk"]) == 15
    assert np.abs(out["pos_scores"].numpy() - z["pos_scores"]).max() <= FP32_ATOL
    if meta["mode"] != "train":
        assert np.array_equal(out["pos_scores"].numpy().argmax(-1), z["pos_scores"].argmax(-1))


def test_oracle_front_matches_reference_golden_at_baseline_shape():
    """t2s_abinet shapes (F=64, 15 OCR/frame, V=5000): the grounding front end is cheap enough on CPU;
    the full 36-pass decode at this shape is covered on the GPU box and by the dedup schedule below."""
    z, meta, d, sd, inp = load_golden("t2s_abinet_eval")
    with torch.no_grad():
        _, _, _, _, g, _ = O.front_t2s(sd, d, inp, neg_frame_override=_neg_override(z))
    assert np.array_equal(g["ground_frame"].numpy(), z["ground_frame"])
    assert np.array_equal(g["ground_bbox"].numpy(), z["ground_box"])
    assert np.array_equal(g["debug"]["frame_pos_topk"].numpy(), z["pos_frame_topk_mask"])


def test_oracle_dedup_schedule_matches_reference_golden_at_baseline_shape():
    z, meta, d, sd, inp = load_golden("t2s_abinet_eval")
    one = {k: (v[:1] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == meta["batch"] else v)
           for k, v in inp.items()}
    ovr = _neg_override(z)[:1]
    with torch.no_grad():
        out = O.forward_t2s(sd, d, one, schedule="dedup", neg_frame_override=ovr)
    for k in ("ref_scores", "pos_scores", "neg_scores"):
        err = np.abs(out[k].numpy() - z[k][:1]).max()
        assert err <= FP32_ATOL, (k, err)
        assert np.array_equal(out[k].numpy().argmax(-1), z[k][:1].argmax(-1))


def test_oracle_is_deterministic_and_gumbel_sensitive():
    z, meta, d, sd, inp = load_golden("t2s_small_eval")
    with torch.no_grad():
        a = O.forward_t2s(sd, d, inp, schedule="dedup")
        b = O.forward_t2s(sd, d, inp, schedule="dedup")
    for k in ("pos_scores", "ground_frame", "ground_box"):
        assert torch.equal(a[k], b[k])
    inp2 = dict(inp)
    g = torch.Generator().manual_seed(99)
    inp2["gumbel_frame"] = -torch.empty_like(inp["gumbel_frame"]).exponential_(generator=g).log()
    with torch.no_grad():
        c = O.forward_t2s(sd, d, inp2, schedule="dedup")
    # the ref variant sees the dataset masks, not the grounded ones: its first decoder row (prev = BOS)
    # cannot depend on the noise; later rows may, through the greedy `pos` decode that feeds prev_inds
    assert torch.equal(a["ref_scores"][:, 0], c["ref_scores"][:, 0])


def test_oracle_autograd_matches_reference_backward_golden():
    """loss.backward() of the REAL reference (training mode, dropout 0, its own loss classes; fixture
    t2s_small_train_grads) against autograd through the oracle: pins the gradient oracle that the B200 backward is
    held to in tests/test_train_gpu.py."""
    z, meta, d, sd, inp = load_golden("t2s_small_train_grads")
    w_bce, w_nce = meta["loss_weights"]
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    out = O.forward_t2s(sd, d, inp, training=True, neg_frame_override=_neg_override(z))
    assert np.array_equal(out["ground_frame"].numpy(), z["ground_frame"])
    bce = O.pos_bce_loss(out["pos_scores"], inp["targets"], inp["train_loss_mask"])
    nce = O.info_nce(out["ref_scores"], out["pos_scores"], out["neg_scores"])
    ref_bce, ref_nce = float(z["loss_pos_bce"].reshape(-1)[0]), float(z["loss_info_nce"].reshape(-1)[0])
    assert abs(float(bce) - ref_bce) <= 1e-5 * max(1.0, abs(ref_bce))
    assert abs(float(nce) - ref_nce) <= 1e-4
    (w_bce * bce + w_nce * nce).backward()
    names = [str(n) for n in z["grad_names"]]
    norms = dict(zip(names, z["grad_norms"]))
    scale = max(norms.values())
    checked = 0
    for n, p in sd.items():
        if not p.requires_grad:
            continue
        g = p.grad
        if n not in norms:
            assert g is None or float(g.abs().max()) == 0.0, n        # no gradient in the reference either
            continue
        gref = z["g:" + n]
        if g is None:
            assert norms[n] == 0.0, n
            continue
        stride = max(1, g.numel() // meta["grad_sample"])
        got = g.flatten()[::stride][:meta["grad_sample"]].numpy()
        # fp32 reduction-order noise only; tiny gradients (e.g. key biases, mathematically zero) are compared
        # against the scale of the whole gradient
        assert abs(float(g.double().norm()) - norms[n]) <= 2e-3 * norms[n] + 1e-6 * scale, (n, float(g.norm()), norms[n])
        assert np.abs(got - gref).max() <= 2e-3 * np.abs(gref).max() + 1e-6 * scale, n
        checked += 1
    assert checked >= 150
