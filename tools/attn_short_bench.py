import torch, sys
sys.path.insert(0, '/root/repo')
from vitxt_gqa_b200 import lib as tlib
L = tlib.get_lib()
B, Ls, H = 64, 1044, 768
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B * Ls, 3 * H, device="cuda", generator=g)).to(torch.bfloat16)
out = torch.empty(B * Ls, H, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for nk in (50, 128, 345, 564, 1044):
    keys = torch.stack([torch.randperm(Ls, device="cuda")[:nk].sort().values for _ in range(B)]).int()
    kpad = torch.zeros(B, Ls, dtype=torch.int32, device="cuda"); kpad[:, :nk] = keys
    n = torch.full((B,), nk, dtype=torch.int32, device="cuda")
    res = {}
    for name, fn in (("tc", lambda: L.attn_tc(qkv.data_ptr(), 3 * H, 0, B, Ls, H, 12, kpad.data_ptr(), n.data_ptr(), Ls, out.data_ptr(), H, st)),
                     ("mma", lambda: L.attn_bf16(qkv.data_ptr(), 3 * H, B, Ls, H, 12, kpad.data_ptr(), n.data_ptr(), Ls, out.data_ptr(), H, st))):
        for _ in range(3): fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        ts.sort(); res[name] = ts[len(ts) // 2]
    print(nk, {k: round(v * 1e3, 1) for k, v in res.items()}, "us")
