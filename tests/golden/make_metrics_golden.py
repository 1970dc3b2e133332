"""Generates tests/golden/metrics_golden.json by running the REAL reference evaluation code
(/root/reference/pythia/modules/metrics.py and pythia/utils/m4c_evaluators.py, imported in place -- nothing is
copied) on seeded synthetic batches (vitxt_gqa_b200/synth.py make_metrics_case, regenerated identically by the tests).

    python tests/golden/make_metrics_golden.py

What is patched, and why: `np.load` inside pythia.modules.metrics returns the synthetic annotation records (the
reference hard-codes /data/zsheng/... paths, metrics.py:250-253); `editdistance` (absent third-party dependency) is
the Levenshtein restatement of vitxt_gqa_b200/metrics.py, so the ANLS numbers pin the wrapper, not the distance.
The evaluators' `eval_pred_list` are wrapped to record the lists they build.
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("T2S_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)

from vitxt_gqa_b200 import synth  # noqa: E402

CASES = {
    "t2s": dict(B=8, T=12, V=120, O=24, frame_topk=3, ocr_topk=2, n_boxes=16, seed=21),
    "t2s_k5": dict(B=16, T=12, V=200, O=30, frame_topk=5, ocr_topk=5, n_boxes=40, seed=22),
    "m4c": dict(B=8, T=12, V=120, O=24, frame_topk=1, ocr_topk=1, n_boxes=1, seed=23),
    "short_boxes": dict(B=6, T=6, V=60, O=10, frame_topk=4, ocr_topk=3, n_boxes=7, seed=24),   # slices run off the list
}


def main():
    # import the product module BEFORE the reference is importable: with `pythia` on sys.path it would bind to the
    # real registry and take over the metric keys, and this script must run the reference's own classes
    from vitxt_gqa_b200.metrics import edit_distance
    sys.path.insert(0, REF)
    ed = types.ModuleType("editdistance")
    ed.eval = edit_distance
    sys.modules["editdistance"] = ed
    import warnings
    warnings.simplefilter("ignore")
    from pythia.common.registry import registry
    from pythia.common.sample import SampleList
    import pythia.modules.metrics as ref_metrics
    import pythia.utils.m4c_evaluators as ref_eval

    class Writer:
        def write(self, *a, **k):
            pass
    registry.register("writer", Writer())
    golden = {"cases": {}}

    # ---- the text normaliser and the two answer evaluators on their own
    proc = ref_eval.EvalAIAnswerProcessor()
    corpus = sorted(set(list(proc.CONTRACTIONS) + list(synth._ANSWER_WORDS) + [
        "The Coca, Cola?", "a  an the", "it's 5 o'clock", "1,000 dollars", "no. 5", "3.14", "u.s.a", "x ; y", "x;y",
        "tab\tsep\nline", "  padded  ", "Don't Stop", "mr. smith's car", "none of the above", "ten, nine", "a.b.c.d",
        "." * 40, "www.site.com", "50%", "#1", "rock & roll", "é accent", "UPPER lower", ""]))
    golden["normalise"] = [[s, proc(s)] for s in corpus]
    golden["contractions"] = dict(proc.CONTRACTIONS)
    tv = ref_eval.TextVQAAccuracyEvaluator()
    soft = []
    for agree in range(11):
        gts = ["stop"] * agree + ["go %d" % i for i in range(10 - agree)]
        soft.append([gts, tv.eval_pred_list([], [{"pred_answer": "Stop", "gt_answers": gts}])[0][0]])
    golden["soft_accuracy"] = soft
    an = ref_eval.STVQAANLSEvaluator()
    pairs = [["coca cola", ["coca-cola", "pepsi"]], ["stop", ["STOP "]], ["abc", ["xyz", "abd"]], ["a", ["bb"]],
             ["main street", ["main st", "main st."]]]
    golden["anls"] = [[p, g, an.eval_pred_list([], [{"pred_answer": p, "gt_answers": g}])[0][0]] for p, g in pairs]

    # ---- the six registered metrics end to end, plus the lists the evaluators build on the way
    for name, kw in CASES.items():
        case = synth.make_metrics_case(**kw)
        registry.register("vtextgqa_answer_processor", synth.SynthAnswerProcessor(case["vocab"]))
        records = np.array([{"header": 1}] + case["records"], dtype=object)        # the file's first row is skipped
        real_load = ref_metrics.np.load
        captured = {}

        def wrap(cls, key):
            orig = cls.eval_pred_list

            def rec(self, pred_scores, pred_list, *a, **k):
                res = orig(self, pred_scores, pred_list, *a, **k)
                tag = key + ("@%s" % k["threshold"] if "threshold" in k else "")
                captured[tag] = {"pred_list": pred_list, "scores": list(res[0]), "accuracy": res[1]}
                return res
            cls.eval_pred_list = rec
            return orig
        saved = [(c, wrap(c, k)) for c, k in ((ref_eval.TextVQAAccuracyEvaluator, "vqa"),
                                              (ref_eval.STVQAANLSEvaluator, "anls"),
                                              (ref_eval.BoxGroundAccuracyEvaluator, "box"))]
        ref_metrics.np.load = lambda *a, **k: records
        try:
            out = {}
            for dataset_type in ("val", "train"):
                sl, mo = synth.metrics_sample_list(case, SampleList, dataset_type=dataset_type)
                m = ref_metrics.Metrics(["textvqa_accuracy", "stvqa_anls", "IOU@0.3", "IOU@0.5", "GQA@0.3", "GQA@0.5"])
                assert type(m.metrics["IOU@0.3"]).__module__ == "pythia.modules.metrics"
                vals = m(sl, mo)
                out[dataset_type] = {k: float(v) for k, v in vals.items()}
        finally:
            ref_metrics.np.load = real_load
            for c, o in saved:
                c.eval_pred_list = o
        temporal = ref_eval.TempGroundAccuracyEvaluator().eval_pred_list(captured["box@0.5"]["pred_list"])
        golden["cases"][name] = {
            "kwargs": kw, "metrics": out, "temporal_accuracy": temporal,
            "pred_answers": [e["pred_answer"] for e in captured["vqa"]["pred_list"]],
            "vqa_scores": captured["vqa"]["scores"], "anls_scores": captured["anls"]["scores"],
            "box_scores@0.3": captured["box@0.3"]["scores"], "box_scores@0.5": captured["box@0.5"]["scores"],
        }
        print(name, out["val"])
    path = os.path.join(ROOT, "tests", "golden", "metrics_golden.json")
    with open(path, "w") as f:
        json.dump(golden, f, indent=0, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
