"""Evaluation step.  CPU: the oracle restatement and the host string side against the golden
numbers produced by the REAL reference metric classes (tests/golden/make_metrics_golden.py).  GPU: the kernels
through the C ABI and the six registered metrics end to end against the same golden numbers, bit for bit."""
import json
import os

import pytest
import torch

from oracle import metrics_oracle as MO
from vitxt_gqa_b200 import lib as tlib, metrics as M, synth
from vitxt_gqa_b200.pythia_api import SampleList, registry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "metrics_golden.json")))
CASES = sorted(GOLDEN["cases"])


def _case(name):
    g = GOLDEN["cases"][name]
    return g, synth.make_metrics_case(**g["kwargs"])


def _entries(case):
    by_id = {r["question_id"]: r for r in case["records"]}
    out = []
    for b, q in enumerate(case["question_id"]):
        r = by_id[q]
        out.append({"pred_frame": case["ground_frame"][b].tolist(), "pred_box": case["ground_box"][b].tolist(),
                    "ocr_topk": case["ocr_topk"], "st_gt": r["spatial_temporal_gt"], "video_fps": r["fps"],
                    "width": r["width"], "height": r["height"]})
    return out


# ------------------------------------------------------------------------------------------------ CPU
def test_contraction_table_is_the_reference_table():
    assert M.EvalAIAnswerProcessor.CONTRACTIONS == GOLDEN["contractions"]


def test_normaliser_matches_the_reference_on_the_corpus():
    proc = M.EvalAIAnswerProcessor()
    assert len(GOLDEN["normalise"]) > 150
    for raw, want in GOLDEN["normalise"]:
        assert proc(raw) == want, raw


def test_soft_accuracy_and_anls_match_the_reference():
    tv = M.TextVQAAccuracyEvaluator()
    for gts, want in GOLDEN["soft_accuracy"]:
        got = tv.eval_pred_list([], [{"pred_answer": "Stop", "gt_answers": gts}])[0][0]
        assert got == want                      # same float, not close: the summation order is the reference's
    an = M.STVQAANLSEvaluator()
    for pred, gts, want in GOLDEN["anls"]:
        assert an.eval_pred_list([], [{"pred_answer": pred, "gt_answers": gts}])[0][0] == want
    assert M.edit_distance("kitten", "sitting") == 3 and M.edit_distance("", "abc") == 3
    assert M.edit_distance("flaw", "lawn") == 2 and M.edit_distance("same", "same") == 0


def test_edit_distance_bit_parallel_equals_the_textbook_recurrence():
    import random

    def dp(a, b):
        prev = list(range(len(b) + 1))
        for i, ca in enumerate(a, 1):
            cur = [i]
            for j, cb in enumerate(b, 1):
                cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
            prev = cur
        return prev[-1]
    rng = random.Random(5)
    for _ in range(1500):
        a = "".join(rng.choice("abcd '") for _ in range(rng.randint(0, 70)))
        b = "".join(rng.choice("abcd '") for _ in range(rng.randint(0, 70)))
        assert M.edit_distance(a, b) == dp(a, b), (a, b)


@pytest.mark.parametrize("name", CASES)
def test_oracle_box_and_temporal_scores_match_the_reference(name):
    g, case = _case(name)
    entries = _entries(case)
    for thr in (0.3, 0.5):
        scores, acc = MO.box_accuracy(entries, thr)
        assert scores == g["box_scores@%s" % thr]
        assert torch.tensor(acc).float().item() == g["metrics"]["val"]["val/vtextgqa/IOU@%s" % thr]
    assert MO.temporal_accuracy(entries)[1] == g["temporal_accuracy"]
    assert len(g["box_scores@0.3"]) != len(entries) or name == "m4c"     # E1: the list is not one entry per sample


@pytest.mark.parametrize("name", CASES)
def test_oracle_answer_cut_spells_the_reference_answers(name):
    g, case = _case(name)
    ids = case["pos_scores"].argmax(-1)
    assert torch.equal(ids, case["planted_ids"])          # ties resolve to the lowest index
    V = case["V"]
    for b in range(ids.shape[0]):
        kept = MO.answer_cut(ids[b].tolist(), V, 2)
        words = [M.ocr_word(case["ocr_tokens"][b][a - V]) if a >= V else case["vocab"][a] for a in kept]
        assert " ".join(words).replace(" 's", "'s") == g["pred_answers"][b]


def test_annotation_packing():
    recs = [{"header": 1}, {"question_id": 7, "fps": 10, "width": 100, "height": 50, "spatial_temporal_gt": [
        {"temporal_gt": [0.25, 1.0], "bbox_gt": {"3": [1, 2, 3, 4], "03": [9, 9, 9, 9], "x": [0, 0, 0, 0]}},
        {"temporal_gt": [2.0, 2.5], "bbox_gt": {}}]},
        {"question_id": 7, "fps": 1, "width": 1, "height": 1, "spatial_temporal_gt": []}]
    ann = M.GroundAnnotations(recs)
    h = ann._host
    assert ann.n_records == 2 and ann.index == {7: 0}                       # the first record of an id wins
    assert h["span_ptr"].tolist() == [0, 2, 2] and h["span_st"].tolist() == [3, 21] and h["span_ed"].tolist() == [11, 26]
    assert h["box_ptr"].tolist() == [0, 1, 1] and h["box_frame"].tolist() == [3]    # only canonical integer keys
    assert ann.record_index(torch.tensor([7, 8])).tolist() == [0, -1]


def test_metric_registry_keys_and_container_errors():
    for key in ("textvqa_accuracy", "stvqa_anls", "IOU@0.3", "IOU@0.5", "GQA@0.3", "GQA@0.5"):
        assert registry.mapping["metric_name_mapping"][key].NAME == key
    with pytest.raises(ValueError):
        M.Metrics(["no_such_metric"])
    with pytest.raises(TypeError):
        M.Metrics([3])
    m = M.Metrics([{"type": "IOU@0.3"}, "textvqa_accuracy"])
    assert list(m.metrics) == ["IOU@0.3", "textvqa_accuracy"]
    sl = SampleList()
    sl.add_field("x", torch.zeros(2, 1))
    assert m(sl, {}) == {}                                  # no `targets`: nothing is evaluated (metrics.py:103-104)


def test_grounding_metrics_have_no_cpu_fallback():
    g, case = _case("t2s")
    registry.register("vtextgqa_answer_processor", synth.SynthAnswerProcessor(case["vocab"]))
    registry.register("ground_annotations", {"val": M.GroundAnnotations(case["records"])})
    sl, out = synth.metrics_sample_list(case, SampleList)
    with pytest.raises(tlib.T2SLibraryError):
        M.Metrics(["IOU@0.3"])(sl, out)
    with pytest.raises(tlib.T2SLibraryError):
        M.Metrics(["textvqa_accuracy"])(sl, out)


# ------------------------------------------------------------------------------------------------ GPU
def _on_gpu(case, dataset_type="val"):
    registry.register("vtextgqa_answer_processor", synth.SynthAnswerProcessor(case["vocab"]))
    ann = M.GroundAnnotations(case["records"])
    registry.register("ground_annotations", {"val": ann, "test": ann})
    sl, out = synth.metrics_sample_list(case, SampleList, dataset_type=dataset_type)
    sl = sl.to("cuda")
    out = {k: v.cuda() for k, v in out.items()}
    return sl, out


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_six_metrics_match_the_reference_classes(name):
    g, case = _case(name)
    names = ["textvqa_accuracy", "stvqa_anls", "IOU@0.3", "IOU@0.5", "GQA@0.3", "GQA@0.5"]
    before = tlib.get_lib().launches
    sl, out = _on_gpu(case)
    m = M.Metrics(names)
    vals = m(sl, out)
    assert tlib.get_lib().launches - before == 2             # one answer_decode + one ground_metrics for all six
    assert set(vals) == set(g["metrics"]["val"])
    for k, want in g["metrics"]["val"].items():
        assert vals[k].shape == (1,) and vals[k].dtype == torch.float32
        assert vals[k].item() == want, k
    assert registry.get("metrics.vtextgqa.val") is vals
    # Q23: a training batch evaluates the two answer metrics only, and drops the others from the object for good
    sl, out = _on_gpu(case, dataset_type="train")
    vals = m(sl, out)
    assert {k: v.item() for k, v in vals.items()} == g["metrics"]["train"]
    assert list(m.metrics) == ["textvqa_accuracy", "stvqa_anls"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernels_reproduce_the_evaluator_lists(name):
    g, case = _case(name)
    sl, out = _on_gpu(case)
    ev = M.BatchEval(sl, out)
    ids, lens, V, _ = ev.answer_ids()
    assert torch.equal(ids.long(), case["planted_ids"])
    for b in range(ids.shape[0]):
        assert ids[b, :int(lens[b])].tolist() == MO.answer_cut(case["planted_ids"][b].tolist(), V, 2)
    assert [p["pred_answer"] for p in ev.qa_predictions()] == g["pred_answers"]
    gr = ev.grounding()
    for t, thr in enumerate((0.3, 0.5)):
        flat = []
        for b in range(ids.shape[0]):
            flat += [1] * int(gr["ones"][t, b]) + [0] * int(gr["tail_zero"][t, b])
        assert flat == g["box_scores@%s" % thr]
        assert gr["head"][t].tolist() == flat[:ids.shape[0]]
    assert gr["acc"][2].item() == torch.tensor(g["temporal_accuracy"]).float().item()


@pytest.mark.gpu
def test_large_random_batch_against_the_oracle():
    case = synth.make_metrics_case(B=512, T=12, V=5000, O=960, frame_topk=5, ocr_topk=5, n_boxes=320, seed=99)
    sl, out = _on_gpu(case)
    ev = M.BatchEval(sl, out)
    ids, lens, V, _ = ev.answer_ids()
    assert torch.equal(ids.long(), case["pos_scores"].argmax(-1))
    gr = ev.grounding()
    entries = _entries(case)
    for t, thr in enumerate((0.3, 0.5)):
        scores, acc = MO.box_accuracy(entries, thr)
        per = [MO.box_scores(e, thr) for e in entries]
        assert gr["ones"][t].tolist() == [sum(p) for p in per]
        assert gr["tail_zero"][t].tolist() == [len(p) - sum(p) for p in per]
        assert gr["acc"][t].item() == torch.tensor(acc).float().item()
        assert gr["head"][t].tolist() == scores[:512]
    assert gr["t_hit"].tolist() == MO.temporal_accuracy(entries)[0]


@pytest.mark.gpu
def test_evaluator_errors_surface_like_the_reference():
    g, case = _case("t2s")
    sl, out = _on_gpu(case)
    bad = dict(out)
    bad["ground_box"] = out["ground_box"].flip(-1).contiguous()      # x1 > x2: `assert pred_bbox[0]<=pred_bbox[2]`
    flipped = any(sum(MO.box_scores(e, 0.3)) for e in _entries(case))
    assert flipped
    with pytest.raises(AssertionError):
        M.BatchEval(sl, bad).grounding()
    sl["question_id"] = sl["question_id"] + 100000                  # no annotation: `None['spatial_temporal_gt']`
    with pytest.raises(TypeError):
        M.BatchEval(sl, out).grounding()


@pytest.mark.gpu
def test_prediction_dump_matches_the_reference_loop():
    """`format_for_evalai` (vtextgqa/dataset.py:315-362) restated as the plain python loop over the same report."""
    g, case = _case("t2s")
    sl, out = _on_gpu(case)
    proc = synth.SynthAnswerProcessor(case["vocab"])
    B, T, N = out["pos_scores"].shape
    report = {"question_id": sl["question_id"], "image_id": ["vid%d" % i for i in range(B)],
              "context_tokens": case["ocr_tokens"], "scores": out["pos_scores"].view(-1, N),
              "ground_frame": out["ground_frame"], "ground_box": out["ground_box"]}
    got = M.format_for_evalai(report, proc)
    V = case["V"]
    pred = out["pos_scores"].argmax(dim=-1).view(B, -1).cpu()
    for b in range(B):
        words, src = [], []
        for a in pred[b].tolist():
            if a >= V:
                words.append(M.ocr_word(case["ocr_tokens"][b][a - V])); src.append("OCR")
            else:
                if a == proc.EOS_IDX:
                    break
                words.append(proc.idx2word(a)); src.append("VOCAB")
        assert got[b] == {"question_id": int(sl["question_id"][b]), "video_id": "vid%d" % b,
                          "answer": " ".join(words).replace(" 's", "'s"),
                          "grounded frame": out["ground_frame"][b].tolist(), "grounded box": out["ground_box"][b].tolist(),
                          "pred_source": src}
        assert got[b]["answer"] == g["pred_answers"][b]
