import ctypes, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitxt_gqa_b200 import lib as tlib
L = tlib.get_lib()
cd = L.cdll
cd.t2s_attn_trace_read.restype = ctypes.c_int
cd.t2s_attn_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
B, Ln, H = 64, 1044, 768
g = torch.Generator(device="cuda").manual_seed(0)
st = torch.cuda.current_stream().cuda_stream
def run(x3, nkeys):
    w = 6 * H if x3 else 3 * H
    qkv = (torch.randn(B * Ln, w, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    keys = torch.arange(Ln, device="cuda", dtype=torch.int32).repeat(B, 1).contiguous()
    nk = torch.full((B,), nkeys, device="cuda", dtype=torch.int32)
    out = torch.empty(B * Ln, 2 * H if x3 else H, device="cuda", dtype=torch.bfloat16)
    buf = (ctypes.c_ulonglong * 8192)()
    for i in range(3):
        cd.t2s_attn_trace_read(buf, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.attn_tc(qkv.data_ptr(), w, 3 * H if x3 else 0, B, Ln, H, 12, keys.data_ptr(), nk.data_ptr(), Ln, out.data_ptr(), 2 * H if x3 else H, st)
        e1.record(); torch.cuda.synchronize()
    n = cd.t2s_attn_trace_read(buf, 1)
    ev = sorted(((buf[i] & 0xffffffffff, buf[i] >> 48, (buf[i] >> 40) & 0xff) for i in range(n)))
    t0 = ev[0][0]
    print("== x3" if x3 else "== bf16", "nkeys", nkeys, "launch ms", e0.elapsed_time(e1), "events", n)
    for t, idv, warp in ev:
        print("%8d  id %2d warp %2d" % (t - t0, idv, warp))
run(False, 345)
run(True, 560)
