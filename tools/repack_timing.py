#!/usr/bin/env python
"""How long the bf16 / hi|lo / transposed weight copies take to rebuild after an optimizer step (training step, B200).
    python tools/repack_timing.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vitxt_gqa_b200 import model as tmodel, synth  # noqa: E402
from vitxt_gqa_b200.pythia_api import load_yaml_config, register_defaults  # noqa: E402

cfg = load_yaml_config("t2s_clipocr.yml", {"model_attributes.t2s.text_bert_init_from_bert_base": False})
mcfg = cfg.model_attributes.t2s
mcfg["metrics"] = []
d = synth.dims_from_config(mcfg, vocab=5000)
register_defaults(vocab_size=d.vocab, ocr_max_num=d.ocr)
m = tmodel.T2S(mcfg)
m.build()
m.init_losses_and_metrics()
m.load_state_dict(synth.make_state_dict(d, seed=0, variant="stress"))
m = m.cuda().train()
eng = m.train_engine()
dev = eng.dev
ts = []
for i in range(6):
    for p in m.parameters():          # what an optimizer step does to the version counters
        p.data.mul_(1.0)
        p._version
    m._packed = None
    eng._wt = None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    t0 = time.perf_counter()
    e0.record()
    m._pack(dev)
    eng._wt_pack()
    e1.record()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    ts.append((e0.elapsed_time(e1), host * 1e3))
print("repack device ms / host ms per step:", [(round(a, 2), round(b, 2)) for a, b in ts])
