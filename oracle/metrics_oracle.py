"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's evaluation step.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(vitxt_gqa_b200/metrics.py + csrc/metrics.cu) never does.

Plain python loops over lists, like the reference.  Pinned: tests/test_metrics.py holds every function here to
tests/golden/metrics_golden.json, which tests/golden/make_metrics_golden.py produced by running the REAL reference
classes (pythia/modules/metrics.py, pythia/utils/m4c_evaluators.py, imported in place) on the same seeded batch.
One dependency of the reference is absent everywhere (`editdistance`, unpinned): the ANLS numbers are pinned with
the Levenshtein restatement injected into the real evaluator -- "parity unpinned" for that one function.
"""


def answer_cut(ids, V, eos_idx):
    """ids of one sample (T ints) -> the ids the answer keeps (reference modules/metrics.py:194-206): an id >= V is an
    OCR copy and is always kept; a vocabulary id equal to EOS ends the answer."""
    kept = []
    for a in ids:
        if a >= V:
            kept.append(a)
        else:
            if a == eos_idx:
                break
            kept.append(a)
    return kept


def iou(box1, box2):
    """reference utils/m4c_evaluators.py:335-356 (python floats, +1 pixel convention)"""
    xa, ya = max(box1[0], box2[0]), max(box1[1], box2[1])
    xb, yb = min(box1[2], box2[2]), min(box1[3], box2[3])
    inter = max(0, xb - xa + 1) * max(0, yb - ya + 1)
    a1 = (box1[2] - box1[0] + 1) * (box1[3] - box1[1] + 1)
    a2 = (box2[2] - box2[0] + 1) * (box2[3] - box2[1] + 1)
    return inter / (a1 + a2 - inter)


def box_scores(entry, threshold):
    """The entries ONE prediction appends to the evaluator's score list
    (reference utils/m4c_evaluators.py:372-401 with check_iou, 358-370).
    entry: pred_frame (list of ints), pred_box (list of [x1, y1, x2, y2] normalised), ocr_topk, st_gt, video_fps,
    width, height."""
    out = []
    w, h = entry["width"], entry["height"]
    boxes = [[b[0] * w, b[1] * h, b[2] * w, b[3] * h] for b in entry["pred_box"]]
    k = entry["ocr_topk"]
    flag = False
    for span in entry["st_gt"]:
        st = int(span["temporal_gt"][0] * entry["video_fps"]) + 1
        ed = int(span["temporal_gt"][1] * entry["video_fps"]) + 1
        for i, fr in enumerate(entry["pred_frame"]):
            if st <= int(fr) <= ed and str(int(fr - 1)) in span["bbox_gt"]:
                gt = span["bbox_gt"][str(int(fr - 1))]
                assert gt[0] <= gt[2] and gt[1] <= gt[3]
                best = 0
                for pb in boxes[i * k:(i + 1) * k]:
                    assert pb[0] <= pb[2] and pb[1] <= pb[3]
                    v = iou(gt, pb)
                    if v > best:
                        best = v
                flag = best > threshold
                if flag:
                    out.append(1)
    if not flag:
        out.append(0)
    return out


def box_accuracy(entries, threshold):
    """-> (concatenated score list, accuracy): reference utils/m4c_evaluators.py:372-405"""
    scores = []
    for e in entries:
        scores.extend(box_scores(e, threshold))
    return scores, sum(scores) / len(scores)


def temporal_accuracy(entries):
    """reference utils/m4c_evaluators.py:305-328"""
    scores = []
    for e in entries:
        hit = 0
        for span in e["st_gt"]:
            st = int(span["temporal_gt"][0] * e["video_fps"]) + 1
            ed = int(span["temporal_gt"][1] * e["video_fps"]) + 1
            if any(st <= f <= ed for f in e["pred_frame"]):
                hit = 1
                break
        scores.append(hit)
    return scores, sum(scores) / len(scores)
