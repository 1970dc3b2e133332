#!/usr/bin/env python
"""Times the fusion-transformer GEMM shapes of t2s_abinet (B=64) one by one through the C ABI.
CUDA events on the launching stream, L2 flushed between launches (256 MB memset), median of N.
    python tools/gemm_bench.py [--rows 66816] [--iters 20]
Prints one JSON line per shape with achieved TFLOP/s against the measured/fallback bf16 peak."""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vitxt_gqa_b200 import lib as tlib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=64 * 1044)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()
    L = tlib.get_lib()
    M, H = args.rows, 768
    dev = "cuda"
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    shapes = [  # name, N, K, flags, residual, x3
        ("qkv      N=2304 K=768 ", 2304, 768, 0, False, False),
        ("attn_out N=768  K=768  +res", 768, 768, 0, True, False),
        ("ffn_up   N=3072 K=768  gelu", 3072, 768, tlib.GEMM_GELU, False, False),
        ("ffn_down N=768  K=3072 +res", 768, 3072, 0, True, False),
        ("ptr_key  N=768  K=768 ", 768, 768, 0, False, False),
        ("x3 qkv   N=2304 K=768  f32out", 2304, 768, tlib.GEMM_OUT_F32, False, True),
        ("x3 attn_out N=768 K=768 f32out+res", 768, 768, tlib.GEMM_OUT_F32 | tlib.GEMM_RES_F32, True, True),
        ("x3 ffn_up N=3072 K=768 gelu split", 3072, 768, tlib.GEMM_GELU | tlib.GEMM_OUT_SPLIT, False, True),
        ("x3 ffn_down N=768 K=3072 f32out+res", 768, 3072, tlib.GEMM_OUT_F32 | tlib.GEMM_RES_F32, True, True),
        ("decode cls M=64 N=5000 K=768 f32out", 5000, 768, tlib.GEMM_OUT_F32, False, False),
        ("decode qkv M=64 N=2304 K=768", 2304, 768, 0, False, False),
    ]
    for name, N, K, flags, res, x3 in shapes:
        m = 64 if name.startswith("decode") else M
        kk = 2 * K if x3 else K
        A = (rn(m, kk) * 0.5).to(torch.bfloat16)
        W = (rn(N, kk) * 0.05).to(torch.bfloat16)
        bias = rn(N)
        f32 = bool(flags & tlib.GEMM_OUT_F32)
        split = bool(flags & tlib.GEMM_OUT_SPLIT)
        C = torch.empty(m, 2 * N if split else N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
        R = None
        if res:
            R = rn(m, N) if flags & tlib.GEMM_RES_F32 else rn(m, N).to(torch.bfloat16)
        fn = L.gemm_bf16x3 if x3 else L.gemm_bf16
        call = lambda: fn(A.data_ptr(), kk, W.data_ptr(), kk, bias.data_ptr(), R.data_ptr() if res else None, N,
                          C.data_ptr(), C.shape[1], m, N, K, flags, 0, st)
        for _ in range(3):
            call()
        times = []
        for _ in range(args.iters):
            if not args.no_flush:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call()
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = statistics.median(times)
        fl = 2.0 * m * N * K
        print(json.dumps({"shape": name, "M": m, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1),
                          "mma_tflops": round((3 if x3 else 1) * fl / ms / 1e9, 1)}))


if __name__ == "__main__":
    main()
