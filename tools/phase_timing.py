#!/usr/bin/env python
"""Phase timeline of one eval forward (B=64, t2s_abinet): CUDA events at the phase boundaries of T2S.forward.
    T2S_B200_PHASES=1 [T2S_B200_OVERLAP_SMS=..] [T2S_B200_SKINNY=..] python tools/phase_timing.py"""
import json
import os
import sys

os.environ.setdefault("T2S_B200_PHASES", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vitxt_gqa_b200 import model as tmodel, synth  # noqa: E402
from vitxt_gqa_b200.pythia_api import SampleList, load_yaml_config, register_defaults  # noqa: E402

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
reps = []
with torch.no_grad():
    for i in range(8):
        m(sl)
        r = m.phase_report()
        if i >= 3:
            reps.append(r)
keys = list(reps[0])
print(json.dumps({"overlap_sms": m.overlap_sms, 
                  "ms_since_start": {k: round(sum(r[k] for r in reps) / len(reps), 3) for k in keys}}))
