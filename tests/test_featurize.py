"""PHOC featuriser: oracle vs the reference binary's golden vectors on CPU; the CUDA kernel
(through the C ABI) vs golden + oracle on the GPU, bit for bit."""
import os

import numpy as np
import pytest
import torch

from oracle import build_ref, phoc_oracle
from vitxt_gqa_b200 import featurize, lib as tlib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "phoc_golden.npz"))
    tokens = bytes(z["tokens_utf8"]).decode("utf-8").split("\x00")
    rows = np.unpackbits(z["bits"], axis=1)[:, :604].astype(np.float32)
    assert len(tokens) == rows.shape[0] == 1500
    return tokens, rows


# ------------------------------------------------------------------------------------------ CPU
def test_golden_tokens_are_the_seeded_synthetic_ones():
    tokens, _ = _golden()
    assert tokens == synth.make_ocr_tokens(1500, seed=2024)


def test_phoc_oracle_matches_reference_golden_bit_exact():
    tokens, rows = _golden()
    got = np.stack([phoc_oracle.build_phoc(t) for t in tokens])
    assert np.array_equal(got, rows)
    # known structure: "<pad>" is the word "pad"; the empty word is all zero; level-2 halves of "ab"
    assert np.array_equal(phoc_oracle.build_phoc("<pad>"), phoc_oracle.build_phoc("PAD"))
    assert phoc_oracle.build_phoc("?!").sum() == 0
    ab = phoc_oracle.build_phoc("ab")
    assert ab[0] == 1 and ab[36 + 1] == 1 and ab[36] == 0


def test_phoc_oracle_matches_compiled_reference_when_present():
    """oracle/_ref/cphoc.so = the reference's own cphoc.c compiled in place (oracle/build_ref.py)."""
    ref = build_ref.load()
    if ref is None:
        pytest.skip("oracle/_ref/cphoc.so not built (no /root/reference on this box)")
    for t in synth.make_ocr_tokens(400, seed=7):
        clean = phoc_oracle.clean_token(t)
        assert np.array_equal(np.asarray(ref.build_phoc(clean), np.float32), phoc_oracle.phoc_of_clean(clean)), t


def test_phoc_processor_oracle_pads_and_truncates():
    out = phoc_oracle.phoc_processor(["stop", "<pad>", "x"], 5)
    assert out.shape == (5, 604) and out[3:].sum() == 0 and out[0].sum() > 0
    assert np.array_equal(phoc_oracle.phoc_processor(["a", "b", "c"], 2), phoc_oracle.phoc_processor(["a", "b"], 2))


def test_pack_tokens_and_no_cpu_fallback():
    data, off = featurize.pack_tokens(["Hello", "<pad>", "", "café", "K"])
    assert bytes(data.numpy()) == b"Hello<pad>caf\xc3\xa9k" and off.tolist() == [0, 5, 10, 10, 15, 16]
    data, off = featurize.pack_tokens([])
    assert data.numel() == 0 and off.tolist() == [0]
    from vitxt_gqa_b200.pythia_api import registry
    assert issubclass(registry.mapping["processor_name_mapping"]["phoc"], featurize.PhocProcessor)
    if not torch.cuda.is_available():
        with pytest.raises(tlib.T2SLibraryError):
            featurize.phoc_rows(["abc"])
        with pytest.raises(tlib.T2SLibraryError):       # argument errors are caught before any launch
            tlib.get_lib().phoc_build(None, None, 3, 2, None, 604, None)


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_phoc_kernel_matches_reference_golden_bit_exact():
    tokens, rows = _golden()
    got = featurize.phoc_rows(tokens, rows=len(tokens) + 3)
    assert torch.equal(got[: len(tokens)].cpu(), torch.from_numpy(rows))
    assert got[len(tokens):].abs().sum().item() == 0


@pytest.mark.gpu
def test_phoc_kernel_long_and_ragged_tokens_vs_oracle():
    import random
    rng = random.Random(3)
    al = "abcdefghijklmnopqrstuvwxyz0123456789 -.'AZé"
    tokens = ["", "a" * 31, "b" * 32, "c" * 33, "th" * 100, "z9" * 1000]
    tokens += ["".join(rng.choice(al) for _ in range(rng.randint(0, 130))) for _ in range(300)]
    want = np.stack([phoc_oracle.build_phoc(t) for t in tokens])
    # strided output view (ldo > 604, unaligned for float4) and the contiguous one
    buf = torch.full((len(tokens), 607), 7.0, device="cuda")
    got = featurize.phoc_rows(tokens, out=buf[:, 1:605])
    assert torch.equal(got.cpu(), torch.from_numpy(want))
    assert torch.all(buf[:, 0] == 7.0) and torch.all(buf[:, 605:] == 7.0)
    assert torch.equal(featurize.phoc_rows(tokens).cpu(), torch.from_numpy(want))
    assert featurize.phoc_rows([], rows=0).shape == (0, 604)
    assert featurize.phoc_rows([], rows=4).abs().sum().item() == 0


@pytest.mark.gpu
def test_phoc_processor_and_batch_match_the_oracle_at_dataset_shape():
    toks = [synth.make_ocr_tokens(960, seed=s) for s in (1, 2)]
    toks[1] = toks[1][:700]                                   # short sample: zero rows behind it
    feat = featurize.phoc_batch(toks, 960)
    assert feat.shape == (2, 960, 604) and feat.dtype == torch.float32
    for b in range(2):
        assert torch.equal(feat[b].cpu(), torch.from_numpy(phoc_oracle.phoc_processor(toks[b], 960)))
    proc = featurize.PhocProcessor({"max_length": 960})
    r = proc({"tokens": toks[1]})
    assert torch.equal(r["text"], feat[1]) and int(r["length"]) == 700 and r["tokens"][700] == "<pad>"
    r2 = featurize.PhocProcessor({"max_length": 10})({"tokens": toks[0]})
    assert torch.equal(r2["text"], feat[0, :10]) and int(r2["length"]) == 10


@pytest.mark.gpu
def test_forward_takes_ocr_token_text_instead_of_phoc_rows():
    """SURVEY 8f rank 2, wired: a SampleList that carries `ocr_token_bytes` (the OCR token text, 64 B per slot) instead
    of `context_feature_1` (604 fp32 per slot) gives bit-identical outputs to one that carries the PHOC rows the
    reference's CPU processor would have produced for the same tokens (checked against the compiled-reference golden
    rows elsewhere in this file), in eval and in the training step."""
    import torch
    from parity_utils import build_b200_model, sample_list
    from vitxt_gqa_b200 import featurize, synth
    from oracle import phoc_oracle
    d = synth.Dims(frames=8, ocr_per_frame=4, vocab=200, frame_topk=3, ocr_topk=2)
    sd = synth.make_state_dict(d, seed=0, variant="stress")
    for train in (False, True):
        a = synth.make_inputs(d, 3, seed=41, train=train)
        b = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in a.items()}
        tokens = synth.attach_ocr_tokens(b, seed=7)
        assert "context_feature_1" not in b and b["ocr_token_bytes"].shape == (3, d.ocr, 64)
        # the float path gets the rows of the CPU restatement of cphoc.c for the same tokens
        a["context_feature_1"] = torch.from_numpy(np.stack(
            [np.stack([phoc_oracle.build_phoc(t) for t in toks]) for toks in tokens])).float()
        m = build_b200_model(d, sd, train=train)
        with torch.no_grad():
            oa = m(sample_list(a))
            ob = m(sample_list(b))
        torch.cuda.synchronize()
        for k in ("pos_scores", "ref_scores", "neg_scores", "ground_frame", "ground_box"):
            assert torch.equal(oa[k], ob[k]), (train, k)
    rec = featurize.pack_tokens_fixed(["Hello", "<pad>", "naïve-42"]).cuda()
    rows = featurize.phoc_from_records(rec)
    assert torch.equal(rows.cpu(), featurize.phoc_rows(["Hello", "<pad>", "naïve-42"]).cpu())
    with pytest.raises(ValueError):
        featurize.pack_tokens_fixed(["x" * 65])


# ------------------------------------------------------------------------------------------------------------------
# frame sampling + per-frame OCR truncate / pad / pack (vtextgqa/dataset.py:103-253): goldens produced by the
# reference's own source text (tests/golden/make_pack_golden.py)
# ------------------------------------------------------------------------------------------------------------------
PACK_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ocr_pack_golden.npz")
PACK_FIELDS = ("ocr_bbox_coordinates", "track_id", "temporal_id", "ocr_mask", "frame_id", "frame_mask", "frame_num",
               "middel_frame_id", "middel_frame_idx")


def _pack_cases():
    z = np.load(PACK_GOLDEN)
    for ci in range(int(z["n_cases"])):
        p = "c%d_" % ci
        n_frames, n_info, F, Of = [int(v) for v in z[p + "geom"]]
        yield ci, {k[len(p):]: z[k] for k in z.files if k.startswith(p)}, n_frames, n_info, F, Of


def test_pack_oracle_matches_the_reference_method_bit_exact():
    from oracle import pack_oracle
    for ci, g, n_frames, n_info, F, Of in _pack_cases():
        r = pack_oracle.pack_ocr_frames(g["det_points"], g["det_track"], g["det_tokens"], g["frame_ptr"], n_info, n_frames,
                                        float(g["size"][0]), float(g["size"][1]), F, Of)
        for k in PACK_FIELDS:
            assert np.array_equal(np.asarray(r[k]), g[k]), (ci, k)
        assert r["ocr_bbox_coordinates"].dtype == g["ocr_bbox_coordinates"].dtype == np.float32
        nt = g["ocr_tokens"].shape[0]                      # len(sampled frames) * Of entries, then the processors' padding
        assert np.array_equal(r["ocr_token_bytes"][:nt], g["ocr_tokens"]) and not r["ocr_token_bytes"][nt:].any(), ci
        assert featurize.sample_frames(n_frames, F) == [int(v) for v in g["frame_id"] if v > 0] == \
            pack_oracle.sample_frames(n_frames, F)


def test_ocr_info_to_csr_round_trip():
    info = {"1": [{"points": [1, 2, 30, 2, 30, 20, 1, 20], "ocr": "Exit", "ID": 7}], "2": [],
            "3": [{"points": [5.5, 6, 9, 6, 9, 8, 5, 8.25], "ocr": "a", "ID": 1},
                  {"points": [0, 0, 1, 0, 1, 1, 0, 1], "ocr": "b", "ID": 2}]}
    c = featurize.ocr_info_to_csr(info, token_processor=str.lower, width=16)
    assert c["frame_ptr"].tolist() == [0, 1, 1, 3] and c["det_track"].tolist() == [7, 1, 2]
    assert c["det_points"].shape == (3, 8) and c["det_points"].dtype == np.float32
    assert bytes(c["det_tokens"][0]).rstrip(b"\0") == b"exit" and c["det_tokens"].shape == (3, 16)
    if not torch.cuda.is_available():          # no CPU fallback: the call refuses without a device
        with pytest.raises(tlib.T2SLibraryError):
            featurize.pack_ocr_frames([dict(c, n_frames=3, width=10, height=10)], 4, 2)


@pytest.mark.gpu
def test_pack_kernel_matches_the_reference_method_bit_exact():
    """Videos of one geometry go through ONE launch as a batch; every field equals what the reference's
    add_sample_details produced for that video, and the token records feed the PHOC kernel."""
    groups = {}
    for ci, g, n_frames, n_info, F, Of in _pack_cases():
        groups.setdefault((F, Of), []).append((ci, g, n_frames))
    assert len(groups) == 2
    for (F, Of), cases in groups.items():
        videos = [{"det_points": g["det_points"], "det_track": g["det_track"], "det_tokens": g["det_tokens"],
                   "frame_ptr": g["frame_ptr"], "n_frames": n_frames, "width": float(g["size"][0]),
                   "height": float(g["size"][1])} for _, g, n_frames in cases]
        out = featurize.pack_ocr_frames(videos, F, Of)
        torch.cuda.synchronize()
        for b, (ci, g, _) in enumerate(cases):
            for k in PACK_FIELDS:
                got = out[k][b].cpu().numpy()
                assert np.array_equal(got.reshape(g[k].shape), g[k]), (ci, k)
            tok = out["ocr_token_bytes"][b].cpu().numpy()
            nt = g["ocr_tokens"].shape[0]
            assert np.array_equal(tok[:nt], g["ocr_tokens"]) and not tok[nt:].any(), ci
        # the records are what t2s_phoc_build_fixed takes: "<pad>" slots get the descriptor of the word "pad", missing frames zero rows
        rows = featurize.phoc_from_records(out["ocr_token_bytes"]).cpu().numpy().reshape(len(cases), F * Of, -1)
        for b, (ci, g, _) in enumerate(cases):
            words = [bytes(r).rstrip(b"\0").decode("utf-8") for r in out["ocr_token_bytes"][b].cpu().numpy()]
            want = np.stack([phoc_oracle.build_phoc(w) for w in words])
            assert np.array_equal(rows[b], want), ci
            assert rows[b][g["ocr_tokens"].shape[0]:].sum() == 0
