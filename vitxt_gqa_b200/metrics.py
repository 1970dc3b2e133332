"""Evaluation step on the device -- the consumer of the forward's outputs.

The reference evaluates every batch on the host: six metric objects each copy `ground_frame` / `ground_box` /
`pos_scores.argmax` to the CPU, `.tolist()` them, re-`np.load` the grounding annotation file and run python loops
(reference pythia/modules/metrics.py:175-546, pythia/utils/m4c_evaluators.py:225-405).  Here one `BatchEval` per
batch launches `t2s_answer_decode` and `t2s_ground_metrics` (csrc/metrics.cu) on the forward's stream and every
metric reads from it: the device does the argmax + EOS cut and the span / IoU evaluation at both thresholds, the
host keeps only the string side (vocabulary lookup, EvalAI normalisation, soft accuracy, ANLS) fed by B x T int32.

Same plugin surface as the reference: the metric classes are registered under the same registry keys
(`textvqa_accuracy`, `stvqa_anls`, `IOU@0.3`, `IOU@0.5`, `GQA@0.3`, `GQA@0.5`), `Metrics(metric_list)` is the
container `BaseModel.init_losses_and_metrics` builds from `config.metrics`, results are keyed
`<dataset_type>/<dataset_name>/<metric>` and registered under `metrics.<dataset_name>.<dataset_type>`.  The
evaluator's quirks (E1-E4 in csrc/metrics.cu, Q23 here) are reproduced; tests pin every number against the real
reference evaluators (tests/golden/make_metrics_golden.py).

There is no CPU fallback for the device part: without the CUDA library / a GPU the grounding metrics raise.
"""
import collections.abc
import pickle
import re

import numpy as np
import torch

from . import lib as _lib
from .pythia_api import registry

# ------------------------------------------------------------------------------------------------ host: strings
# EvalAI answer normalisation (reference m4c_evaluators.py:5-215).  The contraction table is stored by its structure:
# _SINGLE holds contracted forms whose key is the form without apostrophes; _DOUBLE holds forms with several
# apostrophes whose keys are the form with exactly ONE apostrophe dropped; _ODD are the table's three irregular rows.
_SINGLE = (
    "ain't aren't can't could've couldn't didn't doesn't don't hadn't hasn't haven't he'd he's how'd how'll how's "
    "I'm I've isn't it'd it'll ma'am mightn't might've mustn't must've needn't not've o'clock oughtn't shan't "
    "should've shouldn't somebody'll somebody's someone'd someone'll someone's something'd something'll that's "
    "there'd there're there's they'd they'll they're they've 'twas wasn't we've weren't what'll what're what's "
    "what've when's where'd where's where've who'd who'll who's who've why'll why're why's won't would've wouldn't "
    "y'all you'd you'll you're you've").split()
_DOUBLE = (
    "couldn't've hadn't've he'd've I'd've it'd've mightn't've 'ow's'at she'd've shouldn't've somebody'd've "
    "someone'd've something'd've there'd've they'd've we'd've who'd've wouldn't've y'all'll y'all'd've "
    "you'd've").split()
_ODD = {"let's": "let's", "she's": "she's", "somebody'd": "somebodyd"}


def _contraction_table():
    table = {form.replace("'", ""): form for form in _SINGLE}
    for form in _DOUBLE:
        cuts = [i for i, ch in enumerate(form) if ch == "'"]
        for i in cuts:
            table[form[:i] + form[i + 1:]] = form
    table.update(_ODD)
    return table


class EvalAIAnswerProcessor:
    """`processor(answer) -> normalised answer`, the EvalAI convention (m4c_evaluators.py:5-215)."""
    CONTRACTIONS = _contraction_table()
    NUMBERS = dict(zip("none zero one two three four five six seven eight nine ten".split(),
                       "0 0 1 2 3 4 5 6 7 8 9 10".split()))
    ARTICLES = ("a", "an", "the")
    # both patterns exactly as the reference writes them ("(?!<=\\d)" is a look-AHEAD for the text "<=digit")
    PERIOD = re.compile(r"(?!<=\d)(\.)(?!\d)")
    COMMA_IN_NUMBER = re.compile(r"(?<=\d)(\,)+(?=\d)")
    PUNCT = list(";/[]\"{}()=+\\_-><@`,?!")

    def __call__(self, item):
        text = item.lower().replace(",", "").replace("?", "").replace("'s", " 's").strip()
        text = text.replace("\n", " ").replace("\t", " ").strip()
        # punctuation (m4c_evaluators.py:182-193): the test looks at the ORIGINAL text, the edit at the running one
        out = text
        number_comma = self.COMMA_IN_NUMBER.search(text) is not None
        for p in self.PUNCT:
            if (p + " " in text or " " + p in text) or number_comma:
                out = out.replace(p, "")
            else:
                out = out.replace(p, " ")
        # the reference passes re.UNICODE (= 32) as the *count* argument of Pattern.sub: at most 32 periods go
        out = self.PERIOD.sub("", out, 32)
        # digits / articles / contractions (m4c_evaluators.py:195-208)
        words = []
        for w in out.lower().split():
            w = self.NUMBERS.get(w, w)
            if w not in self.ARTICLES:
                words.append(w)
        return " ".join(self.CONTRACTIONS.get(w, w) for w in words)


def edit_distance(a, b):
    """Levenshtein distance of two sequences: the published definition of the absent third-party dependency
    `editdistance` (unpinned in the reference; imported at m4c_evaluators.py:268), computed with the bit-parallel
    recurrence of Myers / Hyyro on python integers (one step per element of the longer sequence)."""
    if len(a) > len(b):
        a, b = b, a
    m = len(a)
    if m == 0:
        return len(b)
    full = (1 << m) - 1
    top = 1 << (m - 1)
    peq = {}
    for i, ch in enumerate(a):
        peq[ch] = peq.get(ch, 0) | (1 << i)
    pv, mv, score = full, 0, m
    for ch in b:
        eq = peq.get(ch, 0)
        xv = eq | mv
        xh = (((eq & pv) + pv) ^ pv) | eq
        ph = mv | (~(xh | pv) & full)
        mh = pv & xh
        if ph & top:
            score += 1
        elif mh & top:
            score -= 1
        ph = ((ph << 1) | 1) & full
        mh = (mh << 1) & full
        pv = mh | (~(xv | ph) & full)
        mv = ph & xv
    return score


class TextVQAAccuracyEvaluator:
    """Soft VQA accuracy against 10 human answers (m4c_evaluators.py:218-257)."""

    def __init__(self):
        self.answer_processor = EvalAIAnswerProcessor()

    def _compute_answer_scores(self, raw_answers):
        answers = [self.answer_processor(a) for a in raw_answers]
        assert len(answers) == 10
        scores = {}
        for cand in set(answers):
            accs = []
            for leave_out in range(len(answers)):     # leave one annotator out, count the others that agree
                agree = sum(1 for i, a in enumerate(answers) if i != leave_out and a == cand)
                accs.append(min(1, float(agree) / 3))
            scores[cand] = sum(accs) / len(accs)
        return scores

    def eval_pred_list(self, pred_scores, pred_list):
        for entry in pred_list:
            pred = self.answer_processor(entry["pred_answer"])
            pred_scores.append(self._compute_answer_scores(entry["gt_answers"]).get(pred, 0.))
        return pred_scores, sum(pred_scores) / len(pred_scores)


class STVQAANLSEvaluator:
    """Average normalised Levenshtein similarity (m4c_evaluators.py:266-288)."""

    def get_anls(self, s1, s2):
        s1, s2 = s1.lower().strip(), s2.lower().strip()
        iou = 1 - edit_distance(s1, s2) / max(len(s1), len(s2))
        return iou if iou >= .5 else 0.

    def eval_pred_list(self, pred_scores, pred_list):
        for entry in pred_list:
            pred_scores.append(max(self.get_anls(entry["pred_answer"], gt) for gt in entry["gt_answers"]))
        return pred_scores, sum(pred_scores) / len(pred_scores)


def decode_object(byte_row):
    """A python object from the byte-tensor encoding the dataset uses for token / answer lists
    (reference pythia/utils/objects_to_byte_tensor.py:34-44: two size bytes, then the pickle)."""
    row = np.asarray(byte_row, dtype=np.uint8)
    size = int(row[0]) * 256 + int(row[1])
    return pickle.loads(row[2:2 + size].tobytes())


def ocr_word(word):
    """reference pythia/utils/text_utils.py:71-78 with its default `remove`"""
    return word.lower().replace(",", "").replace("?", "").replace("'s", " 's").strip()


# ------------------------------------------------------------------------------------------------ annotations
class GroundAnnotations:
    """The grounding annotation file, packed once for the device.

    `records` is what the reference loads every batch with `np.load(path, allow_pickle=True)[1:]`
    (metrics.py:250-254): dicts with `question_id`, `fps`, `width`, `height` and `spatial_temporal_gt` = a list of
    spans {`temporal_gt`: [start_s, end_s], `bbox_gt`: {str(frame): [x1, y1, x2, y2]}}.  Span frame bounds are
    computed here with the reference's own expression `int(t * fps) + 1` (m4c_evaluators.py:386-387)."""

    def __init__(self, records):
        self.index = {}
        span_ptr, span_st, span_ed, box_ptr, box_frame, box_xyxy, wh = [0], [], [], [0], [], [], []
        n = 0
        for rec in records:
            if not (isinstance(rec, collections.abc.Mapping) and "question_id" in rec):
                continue                        # find_dict_by_id skips such rows (metrics.py:243-247)
            self.index.setdefault(rec["question_id"], n)        # the first match wins
            n += 1
            fps = rec["fps"]
            wh.append((float(rec["width"]), float(rec["height"])))
            for span in rec["spatial_temporal_gt"]:
                t0, t1 = span["temporal_gt"][0], span["temporal_gt"][1]
                span_st.append(int(t0 * fps) + 1)
                span_ed.append(int(t1 * fps) + 1)
                for key, box in span["bbox_gt"].items():
                    # the lookup is `str(int(frame - 1)) in bboxs_gt`: only canonical integer strings can match
                    try:
                        fr = int(key)
                    except (TypeError, ValueError):
                        continue
                    if str(fr) != key:
                        continue
                    box_frame.append(fr)
                    box_xyxy.append([float(box[0]), float(box[1]), float(box[2]), float(box[3])])
                box_ptr.append(len(box_frame))
            span_ptr.append(len(span_st))
        self.n_records = n
        self._host = dict(
            span_ptr=torch.tensor(span_ptr, dtype=torch.int32),
            span_st=torch.tensor(span_st, dtype=torch.int64), span_ed=torch.tensor(span_ed, dtype=torch.int64),
            box_ptr=torch.tensor(box_ptr, dtype=torch.int32), box_frame=torch.tensor(box_frame, dtype=torch.int64),
            box_xyxy=torch.tensor(box_xyxy, dtype=torch.float64).reshape(-1, 4),
            rec_wh=torch.tensor(wh, dtype=torch.float64).reshape(-1, 2))
        self._dev = {}

    @classmethod
    def from_npy(cls, path):
        return cls(np.load(path, allow_pickle=True)[1:])

    def record_index(self, question_ids):
        ids = question_ids.tolist() if hasattr(question_ids, "tolist") else list(question_ids)
        return torch.tensor([self.index.get(q, -1) for q in ids], dtype=torch.int32)

    def on(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = {k: v.to(device) for k, v in self._host.items()}
        return self._dev[key]


# the reference hard-codes these two files (metrics.py:250-253); registry key "ground_annotations" overrides them
_REFERENCE_ANNOTATION_FILES = {
    "val": "/data/zsheng/Data_T5_ViteVQA/data/m4vitevqa/ground_annotation/grouding_anno_t1s2val.npy",
    "test": "/data/zsheng/Data_T5_ViteVQA/data/m4vitevqa/ground_annotation/grouding_anno_t1s2test.npy",
}
_ANNOTATION_CACHE = {}


def ground_annotations_for(dataset_type):
    """`registry.register("ground_annotations", {"val": GroundAnnotations | path, "test": ...})`; any type but "val"
    reads the "test" entry like the reference.  Files are loaded and packed once, not once per batch."""
    which = "val" if dataset_type == "val" else "test"
    table = registry.get("ground_annotations", None, no_warning=True) or {}
    src = table.get(which, _REFERENCE_ANNOTATION_FILES[which])
    if isinstance(src, GroundAnnotations):
        return src
    if src not in _ANNOTATION_CACHE:
        _ANNOTATION_CACHE[src] = GroundAnnotations.from_npy(src)
    return _ANNOTATION_CACHE[src]


# ------------------------------------------------------------------------------------------------ device part
class BatchEval:
    """Everything the six metrics need from one (sample_list, model_output) pair, computed once."""
    THRESHOLDS = (0.3, 0.5)

    def __init__(self, sample_list, model_output):
        self.sample_list, self.out = sample_list, model_output
        self._answers = self._qa = self._ground = self._qa_scores = None

    # ---- answers
    def answer_ids(self):
        """(ids [B, T], length [B]) on the host: argmax of every decoding row and the EOS cut, made on the device."""
        if self._answers is None:
            scores = self.out["pos_scores"]
            if not scores.is_cuda:
                raise _lib.T2SLibraryError("metrics run on the CUDA device that holds the forward's outputs "
                                           "(no CPU fallback)")
            proc = registry.get(self.sample_list.dataset_name + "_answer_processor")
            V = int(proc.get_true_vocab_size())
            B, T, N = scores.shape
            assert scores.dtype == torch.float32 and scores.stride(2) == 1 and scores.stride(0) == T * scores.stride(1)
            buf = torch.empty(B * T + B, dtype=torch.int32, device=scores.device)
            with torch.cuda.device(scores.device):
                _lib.get_lib().answer_decode(scores.data_ptr(), scores.stride(1), B, T, N, V, int(proc.EOS_IDX),
                                             buf.data_ptr(), buf[B * T:].data_ptr(),
                                             torch.cuda.current_stream().cuda_stream)
            host = buf.cpu()
            self._answers = (host[:B * T].view(B, T), host[B * T:], V, proc)
        return self._answers

    def qa_predictions(self):
        """[{pred_answer, gt_answers}] exactly as metrics.py:190-214 builds them."""
        if self._qa is None:
            ids, lens, V, proc = self.answer_ids()
            ctx = self.sample_list.context_tokens_enc.cpu().numpy()
            gts = self.sample_list.gt_answers_enc.cpu().numpy()
            preds = []
            for b in range(ids.shape[0]):
                tokens = decode_object(ctx[b])
                words = []
                for a in ids[b, :int(lens[b])].tolist():
                    words.append(ocr_word(tokens[a - V]) if a >= V else proc.answer_vocab.idx2word(a))
                preds.append({"pred_answer": " ".join(words).replace(" 's", "'s"),
                              "gt_answers": decode_object(gts[b])})
            self._qa = preds
        return self._qa

    def qa_soft_scores(self, evaluator):
        """per-sample soft VQA accuracy of the predictions (shared by textvqa_accuracy and the two GQA metrics)"""
        if self._qa_scores is None:
            self._qa_scores = evaluator.eval_pred_list([], self.qa_predictions())[0]
        return self._qa_scores

    # ---- grounding
    def grounding(self):
        """dict(acc = device fp32 [3]: IOU@0.3, IOU@0.5, temporal accuracy; head = host int32 [2, B])."""
        if self._ground is None:
            gf, gb = self.out["ground_frame"], self.out["ground_box"]
            if not gb.is_cuda:
                raise _lib.T2SLibraryError("metrics run on the CUDA device that holds the forward's outputs "
                                           "(no CPU fallback)")
            dev = gb.device
            ann = ground_annotations_for(self.sample_list["dataset_type"])
            d = ann.on(dev)
            B = int(self.sample_list.frame_num.size(0))
            gf = gf.detach().to(torch.int64).contiguous().view(B, -1)
            gb = gb.detach().to(torch.float32).contiguous().view(B, -1, 4)
            rec = ann.record_index(self.sample_list["question_id"]).to(dev, non_blocking=True)
            ints = torch.empty(8 * B, dtype=torch.int32, device=dev)     # ones[2B] tail[2B] t_hit[B] status[B] head[2B]
            acc = torch.empty(3, dtype=torch.float32, device=dev)
            p = ints.data_ptr()
            with torch.cuda.device(dev):
                _lib.get_lib().ground_metrics(
                    gf.data_ptr(), gf.shape[1], gb.data_ptr(), gb.shape[1], int(self.out["ocr_topk"]),
                    rec.data_ptr(), d["span_ptr"].data_ptr(), d["span_st"].data_ptr(), d["span_ed"].data_ptr(),
                    d["box_ptr"].data_ptr(), d["box_frame"].data_ptr(), d["box_xyxy"].data_ptr(),
                    d["rec_wh"].data_ptr(), B, self.THRESHOLDS[0], self.THRESHOLDS[1],
                    p, p + 8 * B, p + 16 * B, p + 20 * B, acc.data_ptr(), p + 24 * B,
                    torch.cuda.current_stream().cuda_stream)
            host = ints.cpu()
            status = host[5 * B:6 * B]
            if bool((status & 4).any()):
                raise TypeError("'NoneType' object is not subscriptable (no grounding annotation for question %r)"
                                % (self.sample_list["question_id"][int((status & 4).nonzero()[0])],))
            if bool((status & 3).any()):
                raise AssertionError("a labelled or predicted box has x1 > x2 or y1 > y2")
            self._ground = dict(acc=acc, head=host[6 * B:].view(2, B), ones=host[:2 * B].view(2, B),
                                tail_zero=host[2 * B:4 * B].view(2, B), t_hit=host[4 * B:5 * B])
        return self._ground


_LAST_EVAL = [None]


def _batch_eval(sample_list, model_output, kwargs):
    """The BatchEval of this (sample_list, model_output): handed down by our `Metrics` container, or -- under the
    reference's own container, which calls the six metrics one by one -- the one the previous metric of the same
    batch made (it keeps that batch's outputs alive until the next batch is evaluated)."""
    ctx = kwargs.get("batch_eval") or _LAST_EVAL[0]
    if ctx is None or ctx.out is not model_output or ctx.sample_list is not sample_list:
        ctx = BatchEval(sample_list, model_output)
    _LAST_EVAL[0] = ctx
    return ctx


# ------------------------------------------------------------------------------------------------ metric plugins
class BaseMetric:
    """reference metrics.py:134-171"""

    def __init__(self, name, *args, **kwargs):
        self.name = name

    def calculate(self, sample_list, model_output, *args, **kwargs):
        raise NotImplementedError("'calculate' must be implemented in the child class")

    def __call__(self, *args, **kwargs):
        return self.calculate(*args, **kwargs)

    def _calculate_with_checks(self, *args, **kwargs):
        return self.calculate(*args, **kwargs)


class TextVQAAccuracy(BaseMetric):
    """reference metrics.py:175-221"""
    NAME = "textvqa_accuracy"

    def __init__(self):
        super().__init__(self.NAME)
        self.evaluator = TextVQAAccuracyEvaluator()

    def calculate(self, sample_list, model_output, *args, **kwargs):
        ctx = _batch_eval(sample_list, model_output, kwargs)
        if isinstance(self.evaluator, TextVQAAccuracyEvaluator):
            scores = ctx.qa_soft_scores(self.evaluator)
            accuracy = sum(scores) / len(scores)
        else:
            _, accuracy = self.evaluator.eval_pred_list([], ctx.qa_predictions())
        return torch.tensor(accuracy).to(sample_list.context_tokens_enc.device)


class STVQAANLS(TextVQAAccuracy):
    """reference metrics.py:224-229"""
    NAME = "stvqa_anls"

    def __init__(self):
        BaseMetric.__init__(self, self.NAME)
        self.evaluator = STVQAANLSEvaluator()


class _BoxGroundAccuracy(BaseMetric):
    """reference metrics.py:233-339 (`IOU@0.3`, `IOU@0.5`): the value stays on the device."""
    NAME, SLOT = None, 0

    def __init__(self):
        super().__init__(self.NAME)

    def calculate(self, sample_list, model_output, *args, **kwargs):
        return _batch_eval(sample_list, model_output, kwargs).grounding()["acc"][self.SLOT].clone()


class _GroundedQAAccuracy(BaseMetric):
    """reference metrics.py:341-546 (`GQA@0.5`, `GQA@0.3`): entry i of the box score list AND soft accuracy == 1."""
    NAME, SLOT = None, 0

    def __init__(self):
        super().__init__(self.NAME)
        self.qa_evaluator = TextVQAAccuracyEvaluator()

    def calculate(self, sample_list, model_output, *args, **kwargs):
        ctx = _batch_eval(sample_list, model_output, kwargs)
        head = ctx.grounding()["head"][self.SLOT].tolist()
        qa = ctx.qa_soft_scores(self.qa_evaluator)
        both = [1 if head[i] == 1 and qa[i] == 1 else 0 for i in range(len(qa))]
        return torch.tensor(sum(both) / len(both)).to(sample_list.frame_num.device)


class IoU03(_BoxGroundAccuracy):
    NAME, SLOT = "IOU@0.3", 0


class IoU05(_BoxGroundAccuracy):
    NAME, SLOT = "IOU@0.5", 1


class GQA05(_GroundedQAAccuracy):
    NAME, SLOT = "GQA@0.5", 1


class GQA03(_GroundedQAAccuracy):
    NAME, SLOT = "GQA@0.3", 0


METRIC_CLASSES = (TextVQAAccuracy, STVQAANLS, IoU03, IoU05, GQA05, GQA03)


def _register():
    """Take over the six metric keys.  The real registry asserts a subclass of ITS BaseMetric (registry.py:119-126),
    so when pythia's metrics module is importable the classes are mixed with it."""
    table = registry.mapping.setdefault("metric_name_mapping", {})
    real_base = None
    try:
        from pythia.modules.metrics import BaseMetric as real_base  # type: ignore
    except Exception:
        pass
    for cls in METRIC_CLASSES:
        table[cls.NAME] = type(cls.__name__, (cls, real_base), {}) if real_base is not None else cls
    return table


_register()


class Metrics:
    """The container `BaseModel.init_losses_and_metrics` builds from `config.metrics` (reference metrics.py:53-131)."""

    def __init__(self, metric_list):
        if not isinstance(metric_list, list):
            metric_list = [metric_list]
        self.writer = registry.get("writer")
        self.metrics = self._init_metrics(metric_list)

    def _init_metrics(self, metric_list):
        metrics = {}
        table = registry.mapping.get("metric_name_mapping", {})
        for metric in metric_list:
            params = {}
            if isinstance(metric, collections.abc.Mapping):
                if "type" not in metric:
                    raise ValueError("Metric {} needs to have 'type' attribute".format(metric))
                params = metric.get("params", {}) or {}
                metric = metric["type"]
            elif not isinstance(metric, str):
                raise TypeError("Metric {} has inappropriate type 'dict' or 'str' allowed".format(metric))
            cls = table.get(metric)
            if cls is None:
                raise ValueError("No metric named {} registered to registry".format(metric))
            metrics[metric] = cls(**params)
        return metrics

    def __call__(self, sample_list, model_output, *args, **kwargs):
        values = {}
        if "targets" not in sample_list:
            return values
        dataset_type, dataset_name = sample_list["dataset_type"], sample_list["dataset_name"]
        kwargs.setdefault("batch_eval", BatchEval(sample_list, model_output))
        with torch.no_grad():
            if dataset_type == "train":
                # Q23: the first training batch drops the grounding metrics from this object for good
                self.metrics = {k: v for k, v in self.metrics.items() if k in {"textvqa_accuracy", "stvqa_anls"}}
            for name, metric in self.metrics.items():
                key = "{}/{}/{}".format(dataset_type, dataset_name, name)
                value = metric._calculate_with_checks(sample_list, model_output, *args, **kwargs)
                value = value.float() if isinstance(value, torch.Tensor) else torch.tensor(value, dtype=torch.float)
                values[key] = value.view(1) if value.dim() == 0 else value
        registry.register("{}.{}.{}".format("metrics", dataset_name, dataset_type), values)
        return values


# ------------------------------------------------------------------------------------------------ prediction dump
def format_for_evalai(report, answer_processor):
    """The prediction entries the reference's test reporter writes (`format_for_evalai`, reference
    pythia/datasets/videoqa/vtextgqa/dataset.py:315-362, fed by pythia/common/test_reporter.py:134-150 with
    `report.scores` = `pos_scores` flattened to [B * T, V + O]): answer string, grounded frames / boxes and the source of
    every answer token.  The argmax and the EOS cut run on the device (`t2s_answer_decode`); `report` needs
    `question_id`, `image_id`, `context_tokens`, `scores`, `ground_frame`, `ground_box`."""
    scores = report["scores"] if isinstance(report, collections.abc.Mapping) else report.scores
    get = (lambda k: report[k]) if isinstance(report, collections.abc.Mapping) else (lambda k: getattr(report, k))
    qids = get("question_id")
    B = len(qids)
    V = int(answer_processor.get_true_vocab_size())
    if not scores.is_cuda:
        raise _lib.T2SLibraryError("format_for_evalai decodes on the CUDA device that holds the scores (no CPU fallback)")
    scores = scores.reshape(B, -1, scores.shape[-1])
    T, N = scores.shape[1], scores.shape[2]
    scores = scores.contiguous()
    buf = torch.empty(B * T + B, dtype=torch.int32, device=scores.device)
    with torch.cuda.device(scores.device):
        _lib.get_lib().answer_decode(scores.data_ptr(), N, B, T, N, V, int(answer_processor.EOS_IDX), buf.data_ptr(),
                                     buf[B * T:].data_ptr(), torch.cuda.current_stream().cuda_stream)
    host = buf.cpu()
    ids, lens = host[:B * T].view(B, T), host[B * T:]
    frames, boxes = get("ground_frame").tolist(), get("ground_box").tolist()
    out = []
    for b in range(B):
        tokens = get("context_tokens")[b]
        words, source = [], []
        for a in ids[b, :int(lens[b])].tolist():
            if a >= V:
                words.append(ocr_word(tokens[a - V]))
                source.append("OCR")
            else:
                words.append(answer_processor.answer_vocab.idx2word(a))
                source.append("VOCAB")
        q = qids[b]
        out.append({"question_id": q.item() if hasattr(q, "item") else q, "video_id": get("image_id")[b],
                    "answer": " ".join(words).replace(" 's", "'s"), "grounded frame": frames[b],
                    "grounded box": boxes[b], "pred_source": source})
    return out
