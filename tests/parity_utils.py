"""Helpers shared by the CPU and GPU parity tests (test infrastructure only)."""
import ast
import os

import numpy as np
import torch

from vitxt_gqa_b200 import synth
from vitxt_gqa_b200.pythia_api import ConfigNode, SampleList, register_defaults

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Stated tolerances (SURVEY 8d "Tolerance anchors"): the answer transformer runs in bf16
# (fp32 accumulate), the grounding chain in fp32.
SCORES_MAX_ABS = 5e-2      # max |logit - reference logit|
SCORES_MEAN_ABS = 5e-3
FP32_CHAIN_ATOL = 2e-4     # joint [txt; frames; ocr] features after TextBert / encoders / QTV
ARGMAX_MARGIN = 1e-1       # answer indices must agree wherever the reference's top1-top2 margin exceeds this
LOSS_RTOL = 1e-2
INFO_NCE_ATOL = 1e-2       # on the unweighted InfoNCE (temperature 0.1 amplifies cosine errors 10x)


# Flip / low-margin counts of every margin-aware comparison of a test session; tests/conftest.py prints them in the
# terminal summary so that tolerated flips are REPORTED, not swallowed (SURVEY hard part 8).
PARITY_REPORT = []


def report_parity(name, **fields):
    PARITY_REPORT.append((name, fields))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = ast.literal_eval(str(z["meta"]))
    d = synth.Dims(**meta["dims"])
    sd = synth.make_state_dict(d, seed=meta["w_seed"], variant=meta["variant"])
    inp = synth.make_inputs(d, meta["batch"], seed=meta["in_seed"], train=(meta["mode"] == "train"))
    return z, meta, d, sd, inp


def build_b200_model(d, sd, train=False, dropout=False):
    """`dropout=False` (default): the training step runs at p = 0, where parity with autograd is defined; True keeps the
    configured probabilities (obj / ocr dropout_prob 0.1, BERT defaults 0.1) like the reference's training."""
    from vitxt_gqa_b200 import model as tmodel
    register_defaults(vocab_size=d.vocab, ocr_max_num=d.ocr)
    cfg = ConfigNode(synth.model_config_for_dims(d))
    cls = {"": tmodel.T2S, "wo_sg": tmodel.T2SWithoutSG, "wo_tg": tmodel.T2SWithoutTG}[d.ablation]
    m = {"t2s": cls, "m4c": tmodel.M4C, "t5vitevqa": tmodel.T5ViteVQA, "gt_box": tmodel.GTBox}[d.model](cfg)
    m.build()
    m.init_losses_and_metrics()
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    m.train(train)
    m.train_dropout = bool(dropout)
    return m


def sample_list(inp, device="cuda", with_noise=True):
    return synth.to_sample_list(inp, SampleList, with_noise=with_noise).to(device)


def margin_aware_argmax_check(ref_scores, got_scores, margin=ARGMAX_MARGIN):
    """Greedy decode is autoregressive: compare answer indices row by row and stop at the first row
    whose reference margin is inside the tolerance band (later rows may legitimately diverge).
    Returns (n_checked, n_mismatch, n_low_margin_rows)."""
    ref = torch.as_tensor(ref_scores).float().cpu()
    got = torch.as_tensor(got_scores).float().cpu()
    top2 = ref.topk(2, dim=-1).values
    marg = top2[..., 0] - top2[..., 1]
    ra, ga = ref.argmax(-1), got.argmax(-1)
    checked = mism = low = 0
    B, T = ra.shape
    for b in range(B):
        for t in range(T):
            if marg[b, t] <= margin:
                low += 1
                if ra[b, t] != ga[b, t]:
                    break          # a tolerated flip: stop checking this sample's later rows
                continue
            checked += 1
            if ra[b, t] != ga[b, t]:
                mism += 1
                break
    return checked, mism, low


def score_errors(ref, got):
    ref = torch.as_tensor(ref).float().cpu()
    got = torch.as_tensor(got).float().cpu()
    diff = (ref - got).abs()
    return diff.max().item(), diff.mean().item()


def reference_prev_inds(ref_pos_scores, bos_idx=1):
    """prev_inds of the reference's LAST decode iteration (t2s.py:321,353): [BOS, argmax(pos_scores)[:-1]]."""
    ref = torch.as_tensor(ref_pos_scores)
    prev = torch.zeros(ref.shape[:2], dtype=torch.int64)
    prev[:, 0] = bos_idx
    prev[:, 1:] = ref.argmax(-1)[:, :-1]
    return prev


def agreeing_prefix_mask(ref_pos_scores, got_pos_scores):
    """[B,T] bool: rows whose decoder INPUTS are identical in both runs -- row t of a sample is comparable
    iff the answer indices agreed on every earlier row (greedy decode feeds them back)."""
    ra = torch.as_tensor(ref_pos_scores).float().cpu().argmax(-1)
    ga = torch.as_tensor(got_pos_scores).float().cpu().argmax(-1)
    agree = (ra == ga).long()
    before = torch.cumprod(torch.cat([torch.ones_like(agree[:, :1]), agree[:, :-1]], 1), 1)
    return before.bool()
