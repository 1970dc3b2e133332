#!/usr/bin/env python
"""Host vs device time of the eval forward (B=64, t2s_abinet), plain call and pipelined submit():
host enqueue time per forward (no synchronise inside the loop), device time per step, and the phase
timeline of pipelined steps relative to the first step's start.
    python tools/pipe_timing.py [steps]"""
import json
import os
import sys
import time

os.environ.setdefault("T2S_B200_PHASES", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vitxt_gqa_b200 import model as tmodel, synth  # noqa: E402
from vitxt_gqa_b200.pythia_api import SampleList, load_yaml_config, register_defaults  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = load_yaml_config("t2s_abinet.yml", {"model_attributes.t2s.text_bert_init_from_bert_base": False,
                             "model_attributes.t2s.metrics": []})
mcfg = cfg.model_attributes.t2s
d = synth.dims_from_config(mcfg, vocab=5000)
register_defaults(vocab_size=d.vocab, ocr_max_num=d.ocr)
m = tmodel.T2S(mcfg)
m.build()
m.init_losses_and_metrics()
m.load_state_dict(synth.make_state_dict(d, seed=0, variant="stress"))
m = m.cuda().eval()
sl = synth.to_sample_list(synth.make_inputs(d, 64, seed=1235, full_frames=True), SampleList).to("cuda")
out = {"overlap_sms": m.overlap_sms}
with torch.no_grad():
    for pipe in (False, True):
        for _ in range(3):
            (m.submit(sl).result() if pipe else m(sl))
        torch.cuda.synchronize()
        m._phase_events = []
        keep = []
        orig_mark = m._mark

        def mark(name, stream=None, _k=keep):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream if stream is not None else torch.cuda.current_stream())
            _k.append((name, ev))
        m._mark = mark
        m._phases_on = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host = []
        e0.record()
        pend = None
        for _ in range(K):
            t0 = time.perf_counter()
            if pipe:
                nxt = m.submit(sl)
                if pend is not None:
                    pend.result()
                pend = nxt
            else:
                m(sl)
            host.append((time.perf_counter() - t0) * 1e3)
        if pipe:
            pend.result()
        e1.record()
        torch.cuda.synchronize()
        m._mark = orig_mark
        first = keep[0][1]
        tl = [(n, round(first.elapsed_time(e), 2)) for n, e in keep]
        out["pipelined" if pipe else "plain"] = {
            "device_ms_per_step": round(e0.elapsed_time(e1) / K, 3),
            "host_enqueue_ms_per_step": [round(h, 2) for h in host],
            "timeline_ms": tl[: 7 * 4]}
print(json.dumps(out))
