#!/usr/bin/env python
"""Measures the two components built beside the forward path the way bench.py measures the path:
device time with CUDA events after warm-up, achieved HBM GB/s against the algorithmic bytes, and the CPU baseline timed
on the box's host cores in the same run (the reference's compiled cphoc.c when oracle/_ref holds it, else the python
oracle; the python oracle of the evaluator).  One JSON line per component.
    python tools/side_bench.py [--batch 64]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vitxt_gqa_b200 import featurize, metrics as M, synth  # noqa: E402
from vitxt_gqa_b200.pythia_api import SampleList, registry  # noqa: E402


def dev_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        z = json.load(open(p))
        for k in ("hbm_gbs", "hbm_copy_gbs", "hbm_GBps"):
            if k in z:
                return float(z[k]), "measured"
    return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    B, O, T, V = args.batch, 960, 12, 5000
    peak, src = peak_gbs()

    # ---- PHOC featuriser: B x 960 OCR tokens -> context_feature_1 [B, 960, 604] fp32
    toks = [synth.make_ocr_tokens(O, seed=100 + b) for b in range(B)]
    flat = [t for s in toks for t in s]
    data, offs = featurize.pack_tokens(flat)
    d_data, d_off = data.cuda(), offs.cuda()
    out = torch.empty(len(flat), featurize.PHOC_DIM, device="cuda")
    L = featurize._lib.get_lib()
    st = torch.cuda.current_stream().cuda_stream
    ms_kernel = dev_ms(lambda: L.phoc_build(d_data.data_ptr(), d_off.data_ptr(), len(flat), len(flat), out.data_ptr(),
                                           out.stride(0), st))
    ms_api = dev_ms(lambda: featurize.phoc_batch(toks, O), iters=5)          # packing + H2D + launch, host included
    t0 = time.perf_counter()
    torch.cuda.synchronize()
    n_cpu = 4096
    from oracle import build_ref, phoc_oracle
    ref = build_ref.load()
    t0 = time.perf_counter()
    if ref is not None:
        for t in flat[:n_cpu]:
            ref.build_phoc(phoc_oracle.clean_token(t))
        kind = "reference (oracle/_ref/cphoc.so, 1 thread)"
    else:
        for t in flat[:n_cpu]:
            phoc_oracle.build_phoc(t)
        kind = "port (oracle/phoc_oracle.py, 1 thread)"
    cpu_s = time.perf_counter() - t0
    bytes_out = len(flat) * featurize.PHOC_DIM * 4 + int(data.numel()) + int(offs.numel()) * 4
    print(json.dumps({
        "component": "PHOC featuriser (t2s_phoc_build)", "tokens": len(flat), "kernel_ms": ms_kernel,
        "tokens_per_s": len(flat) / ms_kernel * 1e3, "api_ms_with_host_packing": ms_api,
        "roofline": {"bound": "hbm", "achieved": bytes_out / ms_kernel / 1e6, "peak": peak, "unit": "GB/s",
                     "frac": bytes_out / ms_kernel / 1e6 / peak, "peak_source": src,
                     "algorithmic_bytes": bytes_out},
        "cpu_baseline": {"value": n_cpu / cpu_s, "unit": "tokens/s", "cores": 1, "kind": kind,
                         "sample": "%d tokens, %.2f s" % (n_cpu, cpu_s)}}))

    # ---- frame sampling + OCR truncate / pad / pack: B videos of 130 frames, up to 20 detections per frame -> 64 x 15 slots
    import random
    import numpy as np
    rng = random.Random(5)
    F, Of, n_frames = 64, 15, 130
    short = [t for t in flat[:2048] if len(t.encode()) <= 64]
    videos = []
    for b in range(B):
        info = {}
        for f in range(1, n_frames + 1):
            dets = []
            for _ in range(rng.randint(0, 20)):
                x, y, w, h = rng.uniform(0, 1200), rng.uniform(0, 700), rng.uniform(5, 200), rng.uniform(5, 80)
                dets.append({"points": [x, y, x + w, y, x + w, y + h, x, y + h], "ocr": rng.choice(short), "ID": rng.randint(1, 400)})
            info[str(f)] = dets
        videos.append(dict(featurize.ocr_info_to_csr(info), n_frames=n_frames, width=1280, height=720))
    ms_pack_api = dev_ms(lambda: featurize.pack_ocr_frames(videos, F, Of), iters=10)      # concat + H2D + launch
    packed = featurize.pack_ocr_frames(videos, F, Of)
    from oracle import pack_oracle
    t0 = time.perf_counter()
    n_cpu_v = 8
    for v in videos[:n_cpu_v]:
        pack_oracle.pack_ocr_frames(v["det_points"], v["det_track"], v["det_tokens"], v["frame_ptr"], n_frames, n_frames,
                                    1280.0, 720.0, F, Of)
    cpu_pack = (time.perf_counter() - t0) / n_cpu_v
    bytes_pack = sum(t.numel() * t.element_size() for t in packed.values())
    print(json.dumps({
        "component": "frame sampling + OCR pad / pack (t2s_pack_ocr_frames)", "videos": B, "slots": F * Of,
        "api_ms_incl_concat_and_h2d": ms_pack_api, "output_bytes": bytes_pack,
        "cpu_baseline": {"value": cpu_pack * 1e3, "unit": "ms per video", "cores": 1, "kind": "port",
                         "sample": "%d videos through oracle/pack_oracle.py (the reference's python list loop restated)" % n_cpu_v}}))

    # ---- evaluation step: all six metrics of one batch
    case = synth.make_metrics_case(B=B, T=T, V=V, O=O, frame_topk=5, ocr_topk=5, n_boxes=320, seed=7)
    registry.register("vtextgqa_answer_processor", synth.SynthAnswerProcessor(case["vocab"]))
    ann = M.GroundAnnotations(case["records"])
    registry.register("ground_annotations", {"val": ann, "test": ann})
    sl, mo = synth.metrics_sample_list(case, SampleList)
    sl = sl.to("cuda")
    mo = {k: v.cuda() for k, v in mo.items()}
    sc = mo["pos_scores"]
    buf = torch.empty(B * T + B, dtype=torch.int32, device="cuda")
    ms_dec = dev_ms(lambda: L.answer_decode(sc.data_ptr(), sc.stride(1), B, T, sc.shape[2], V, 2, buf.data_ptr(),
                                           buf[B * T:].data_ptr(), st))

    def ground():
        ev = M.BatchEval(sl, mo)
        ev.grounding()
    ms_ground = dev_ms(ground, iters=10)                 # includes the small D2H of the per-sample counts
    names = ["textvqa_accuracy", "stvqa_anls", "IOU@0.3", "IOU@0.5", "GQA@0.3", "GQA@0.5"]
    t0 = time.perf_counter()
    for _ in range(5):
        vals = M.Metrics(names)(sl, mo)
        torch.cuda.synchronize()
    ms_all = (time.perf_counter() - t0) / 5 * 1e3
    # CPU baseline: what the reference does per batch for the grounding metrics (D2H + python evaluator at 2 thresholds
    # x (IOU, GQA) = 4 evaluator passes) and for the answer side (argmax on the device, python loop on the host)
    from oracle import metrics_oracle as MO
    by_id = {r["question_id"]: r for r in case["records"]}
    t0 = time.perf_counter()
    for _ in range(3):
        gf = mo["ground_frame"].cpu().tolist()
        gb = mo["ground_box"].cpu().tolist()
        entries = [{"pred_frame": gf[b], "pred_box": gb[b], "ocr_topk": 5, "st_gt": by_id[q]["spatial_temporal_gt"],
                    "video_fps": by_id[q]["fps"], "width": by_id[q]["width"], "height": by_id[q]["height"]}
                   for b, q in enumerate(case["question_id"])]
        for thr in (0.3, 0.5, 0.3, 0.5):
            MO.box_accuracy(entries, thr)
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    bytes_dec = B * T * sc.shape[2] * 4
    print(json.dumps({
        "component": "evaluation step (t2s_answer_decode + t2s_ground_metrics + host strings)", "batch": B,
        "answer_decode_ms": ms_dec, "ground_metrics_ms_incl_d2h": ms_ground, "all_six_metrics_wall_ms": ms_all,
        "values": {k: float(v) for k, v in vals.items()},
        "roofline": {"kernel": "answer_argmax", "bound": "hbm", "achieved": bytes_dec / ms_dec / 1e6, "peak": peak,
                     "unit": "GB/s", "frac": bytes_dec / ms_dec / 1e6 / peak, "peak_source": src,
                     "algorithmic_bytes": bytes_dec},
        "cpu_baseline": {"value": cpu_ms, "unit": "ms per batch (grounding evaluator only)", "cores": 1, "kind": "port",
                         "sample": "3 batches of %d: D2H + .tolist() + 4 evaluator passes (oracle/metrics_oracle.py)" % B}}))


if __name__ == "__main__":
    main()
