// K6 / K7: pointer-network scores, greedy argmax feedback, and the two training losses.
//
// Reference call sites replaced:
//   OcrPtrNet.forward matmul / sqrt(768) / + raw 0/1 mask   pythia/models/t2s.py:661-666 (Q1)
//   torch.cat([fixed_scores, dynamic_ocr_scores])            t2s.py:285 -- both heads write into one
//                                                            [B, T, V+O] fp32 buffer, no concat
//   pos_scores.argmax(-1) -> prev_inds[:, 1:]                t2s.py:353-354
//   POSBCEWithMaskLoss.forward                               pythia/modules/losses.py:329-343
//   InfoNCE.forward                                          losses.py:361-385
#include "common.cuh"
#include "../../include/t2s_b200.h"

namespace t2s {

// scores[b, t0+i, V+o] = (q[b,t0+i] . keyp[b,o]) / sqrt(H) + mask[b,o]
constexpr int PS_THREADS = 256, PS_ROWS = 64, PS_MAXQ = 16;

// Each warp scores PS_WROWS key rows at a time so 12 independent 16-byte loads per lane are in flight (one row
// per iteration left the kernel latency bound at ~1.2 TB/s); MAXQ = 1 is the greedy-decode instantiation.
constexpr int PS_WROWS = 4;

template <int MAXQ>
__global__ void __launch_bounds__(PS_THREADS)
ptr_score_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, int T, int t0, int nq,
                 const __nv_bfloat16* __restrict__ keyp, long long key_batch_stride, long long ldk, int O, int H,
                 const float* __restrict__ mask, long long mask_stride, float* __restrict__ scores, long long ld_scores,
                 int V, float denom) {
    extern __shared__ float qs[];     // [nq][H]
    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_wait();            // one-row launches of the greedy chain come in early (common.cuh)
    pdl_release();
    for (int i = tid; i < nq * H; i += PS_THREADS) {
        const int r = i / H, d = i % H;
        qs[i] = __bfloat162float(q[((long long)b * T + t0 + r) * ldq + d]);
    }
    __syncthreads();
    constexpr int NW = PS_THREADS / 32;
    const int o_end = min(O, (int)(blockIdx.x + 1) * PS_ROWS);
    const __nv_bfloat16* kb = keyp + (long long)b * key_batch_stride;
    for (int o0 = blockIdx.x * PS_ROWS + warp; o0 < o_end; o0 += NW * PS_WROWS) {
        uint4 kv[PS_WROWS][4];
#pragma unroll
        for (int j = 0; j < PS_WROWS; ++j) {
            const int o = o0 + j * NW;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int d = lane * 8 + c * 256;
                kv[j][c] = (o < o_end && d < H) ? *reinterpret_cast<const uint4*>(kb + (long long)o * ldk + d)
                                                : make_uint4(0u, 0u, 0u, 0u);
            }
        }
        float acc[PS_WROWS][MAXQ];
#pragma unroll
        for (int j = 0; j < PS_WROWS; ++j)
#pragma unroll
            for (int r = 0; r < MAXQ; ++r) acc[j][r] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int d = lane * 8 + c * 256;
            if (d < H) {
#pragma unroll
                for (int r = 0; r < MAXQ; ++r)
                    if (r < nq) {
                        const float4 q0 = *reinterpret_cast<const float4*>(qs + r * H + d);
                        const float4 q1 = *reinterpret_cast<const float4*>(qs + r * H + d + 4);
#pragma unroll
                        for (int j = 0; j < PS_WROWS; ++j) {
                            const uint4 k4 = kv[j][c];
                            float a = acc[j][r];
                            a = fmaf(q0.x, bf16lo(k4.x), a); a = fmaf(q0.y, bf16hi(k4.x), a);
                            a = fmaf(q0.z, bf16lo(k4.y), a); a = fmaf(q0.w, bf16hi(k4.y), a);
                            a = fmaf(q1.x, bf16lo(k4.z), a); a = fmaf(q1.y, bf16hi(k4.z), a);
                            a = fmaf(q1.z, bf16lo(k4.w), a); a = fmaf(q1.w, bf16hi(k4.w), a);
                            acc[j][r] = a;
                        }
                    }
            }
        }
#pragma unroll
        for (int j = 0; j < PS_WROWS; ++j) {
            const int o = o0 + j * NW;
            if (o >= o_end) break;        // warp-uniform
            const float m = mask[(long long)b * mask_stride + o];
#pragma unroll
            for (int r = 0; r < MAXQ; ++r)
                if (r < nq) {
                    const float sum = warp_sum(acc[j][r]);
                    if (lane == 0) scores[((long long)b * T + t0 + r) * ld_scores + V + o] = sum / denom + m;
                }
        }
    }
}

// argmax over a row of N fp32 scores (first maximum wins); writes prev_inds[b, t+1] when t+1 < T
__global__ void __launch_bounds__(256)
argmax_feedback_kernel(const float* __restrict__ scores, long long ld_scores, int T, int t0, int nt, int N,
                       long long* __restrict__ prev_inds, int ld_prev, long long* __restrict__ argmax_out) {
    __shared__ float sv[8];
    __shared__ int si[8];
    const int w = blockIdx.x, b = w / nt, t = t0 + w % nt;
    const float* row = scores + ((long long)b * T + t) * ld_scores;
    pdl_wait();
    pdl_release();
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float v = row[i];
        if (v > best) { best = v; bi = i; }      // strided ascending: first max per thread is kept
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
            if (sv[k] > best || (sv[k] == best && si[k] < bi)) { best = sv[k]; bi = si[k]; }
        if (bi == 0x7fffffff) bi = 0;      // row of NaN / -inf only: no element compared greater (t2s_answer_decode maps it to 0 too)
        if (argmax_out) argmax_out[(long long)b * T + t] = bi;
        if (prev_inds && t + 1 < T) prev_inds[(long long)b * ld_prev + t + 1] = bi;
    }
}

// Teacher-forced passes (2 <= nq <= 16 decoder rows at once: the `ref` / `neg` tail of the eval forward, every pass of
// the training step): scores^T[16 keys x 16 queries] per warp on the tensor cores (mma.sync m16n8k16, fp32 accumulate).
// The SIMT kernel above re-did the bf16 unpack of every key row once per query and ran at 0.4 TB/s (231 us per launch
// for the 94 MB of pointer keys at batch 64); here a key row is read once (4-byte fragment loads straight from global
// memory, L1 turns the four k-steps of a 128-byte line into one fetch) and costs two MMAs per 16 hidden dims.
// A = key rows (M = keys), B = the queries (N), staged once per CTA in shared memory with padded rows.
constexpr int PM_THREADS = 128, PM_KEYS = 64, PM_QPAD = 8;     // 4 warps x 16 keys; Q rows padded by 8 bf16 (bank spread)

__global__ void __launch_bounds__(PM_THREADS)
ptr_score_mma_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, int T, int t0, int nq,
                     const __nv_bfloat16* __restrict__ keyp, long long key_batch_stride, long long ldk, int O, int H,
                     const float* __restrict__ mask, long long mask_stride, float* __restrict__ scores,
                     long long ld_scores, int V, float inv_denom) {
    extern __shared__ __align__(16) uint8_t pm_raw[];
    __nv_bfloat16* qs = reinterpret_cast<__nv_bfloat16*>(pm_raw);         // [16][H + PM_QPAD], rows >= nq zero
    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
    const int ldqs = H + PM_QPAD;
    for (int i = tid; i < 16 * (H / 8); i += PM_THREADS) {
        const int r = i / (H / 8), c = (i % (H / 8)) * 8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < nq) v = *reinterpret_cast<const uint4*>(q + ((long long)b * T + t0 + r) * ldq + c);
        *reinterpret_cast<uint4*>(qs + r * ldqs + c) = v;
    }
    __syncthreads();
    const int o0 = blockIdx.x * PM_KEYS + warp * 16;
    if (o0 >= O) return;
    const __nv_bfloat16* kb = keyp + (long long)b * key_batch_stride;
    const int r0 = min(o0 + g, O - 1), r1 = min(o0 + g + 8, O - 1);      // clamped: rows past O are computed, not stored
    const __nv_bfloat16* k0 = kb + (long long)r0 * ldk + tq * 2;
    const __nv_bfloat16* k1 = kb + (long long)r1 * ldk + tq * 2;
    const __nv_bfloat16* q0 = qs + g * ldqs + tq * 2;                      // queries 0-7 (n tile 0), 8-15 (n tile 1)
    const __nv_bfloat16* q1 = qs + (g + 8) * ldqs + tq * 2;
    float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int k = 0; k < H; k += 16) {
        uint32_t a[4];
        a[0] = *reinterpret_cast<const uint32_t*>(k0 + k);
        a[1] = *reinterpret_cast<const uint32_t*>(k1 + k);
        a[2] = *reinterpret_cast<const uint32_t*>(k0 + k + 8);
        a[3] = *reinterpret_cast<const uint32_t*>(k1 + k + 8);
        const uint32_t b00 = *reinterpret_cast<const uint32_t*>(q0 + k), b01 = *reinterpret_cast<const uint32_t*>(q0 + k + 8);
        const uint32_t b10 = *reinterpret_cast<const uint32_t*>(q1 + k), b11 = *reinterpret_cast<const uint32_t*>(q1 + k + 8);
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c0[0]), "+f"(c0[1]), "+f"(c0[2]), "+f"(c0[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b00), "r"(b01));
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c1[0]), "+f"(c1[1]), "+f"(c1[2]), "+f"(c1[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b10), "r"(b11));
    }
    // c[j]: key row g (j < 2) or g + 8, query 2 tq + (j & 1) (+ 8 for the second n tile)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int o = o0 + g + (j >> 1) * 8;
        if (o >= O) continue;
        const float m = mask[(long long)b * mask_stride + o];
        const int qa = 2 * tq + (j & 1), qb = qa + 8;
        if (qa < nq) scores[((long long)b * T + t0 + qa) * ld_scores + V + o] = c0[j] * inv_denom + m;
        if (qb < nq) scores[((long long)b * T + t0 + qb) * ld_scores + V + o] = c1[j] * inv_denom + m;
    }
}

// ------------------------------------------------------------------------------- masked BCE-with-logits
__global__ void __launch_bounds__(256)
bce_partial_kernel(const float* __restrict__ scores, const float* __restrict__ targets, const float* __restrict__ loss_mask,
                   long long rows, int N, double* __restrict__ partial) {
    __shared__ float red[33];
    double acc = 0.0;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const float m = loss_mask[r];
        if (m == 0.f) continue;                       // uniform per block: whole CTA skips the row
        const float* x = scores + r * N;
        const float* z = targets + r * N;
        float s = 0.f;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const float xv = x[i], zv = z[i];
            // (1 - z) * x - log_sigmoid(x),  log_sigmoid(x) = min(x, 0) - log1p(exp(-|x|))
            s += ((1.f - zv) * xv - (fminf(xv, 0.f) - log1pf(expf(-fabsf(xv))))) * m;
        }
        s = block_sum(s, red);
        acc += (double)s;
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(256)
bce_final_kernel(const double* __restrict__ partial, int n, const float* __restrict__ loss_mask, long long rows,
                 float* __restrict__ out) {
    // one CTA: fixed-order strided partial sums + a shared-memory tree, so the result is run-to-run identical
    __shared__ double sd[256];
    __shared__ float sc[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    float c = 0.f;
    for (long long r = threadIdx.x; r < rows; r += 256) c += loss_mask[r];
    sd[threadIdx.x] = s;
    sc[threadIdx.x] = c;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) { sd[threadIdx.x] += sd[threadIdx.x + off]; sc[threadIdx.x] += sc[threadIdx.x + off]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)(sd[0] / (double)fmaxf(sc[0], 1.f));
}

// ------------------------------------------------------------------------------- InfoNCE (per-sample 2-way)
// stats[row] = { |ref|^2, |pos|^2, |neg|^2, ref.pos, ref.neg } for row = (b, t)
__global__ void __launch_bounds__(256)
nce_rowstats_kernel(const float* __restrict__ ref, const float* __restrict__ pos, const float* __restrict__ neg, int N,
                    float* __restrict__ stats) {
    __shared__ float red[33];
    const long long r = blockIdx.x;
    const float* a = ref + r * N;
    const float* p = pos + r * N;
    const float* n = neg + r * N;
    float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float av = a[i], pv = p[i], nv = n[i];
        s[0] = fmaf(av, av, s[0]); s[1] = fmaf(pv, pv, s[1]); s[2] = fmaf(nv, nv, s[2]);
        s[3] = fmaf(av, pv, s[3]); s[4] = fmaf(av, nv, s[4]);
    }
    for (int k = 0; k < 5; ++k) {
        const float v = block_sum(s[k], red);
        if (threadIdx.x == 0) stats[r * 5 + k] = v;
    }
}
__global__ void nce_final_kernel(const float* __restrict__ stats, int B, int T, float temperature, float* __restrict__ out) {
    __shared__ float red[33];
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float qq = 0.f, pp = 0.f, nn = 0.f, qp = 0.f, qn = 0.f;
        for (int t = 0; t < T; ++t) {
            const float* s = stats + ((long long)b * T + t) * 5;
            const float nr = fmaxf(sqrtf(s[0]), 1e-12f), np = fmaxf(sqrtf(s[1]), 1e-12f), ng = fmaxf(sqrtf(s[2]), 1e-12f);
            qq += s[0] / (nr * nr); pp += s[1] / (np * np); nn += s[2] / (ng * ng);
            qp += s[3] / (nr * np); qn += s[4] / (nr * ng);
        }
        const float cp = qp / (fmaxf(sqrtf(qq), 1e-8f) * fmaxf(sqrtf(pp), 1e-8f));
        const float cn = qn / (fmaxf(sqrtf(qq), 1e-8f) * fmaxf(sqrtf(nn), 1e-8f));
        const float lp = cp / temperature, ln = cn / temperature;
        const float mx = fmaxf(lp, ln);
        acc += (mx + logf(expf(lp - mx) + expf(ln - mx))) - lp;     // -log_softmax(logits)[0]
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[0] = acc / (float)B;
}

}  // namespace t2s

using namespace t2s;

extern "C" int t2s_ptr_score(const void* q, long long ldq, int B, int T, int t0, int nq, const void* keyp,
                             long long key_batch_stride, long long ldk, int O, int H, const float* mask,
                             long long mask_stride, float* scores, long long ld_scores, int V, void* stream) {
    if (nq < 1 || nq > PS_MAXQ || (H % 256) || H > 1024 || (ldk % 8) || t0 < 0 || t0 + nq > T) { set_error("ptr_score: bad arguments"); return T2S_ERR_SHAPE; }
    if (nq > 1 && (ldq % 8) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(keyp) & 3) == 0 &&
        (ldk % 2) == 0 && (key_batch_stride % 2) == 0) {
        // several decoder rows per sample: tensor-core kernel (the scalar kernel stays for the one-row greedy step)
        const size_t sm = (size_t)16 * (H + PM_QPAD) * sizeof(__nv_bfloat16);
        dim3 grid((O + PM_KEYS - 1) / PM_KEYS, B);
        ptr_score_mma_kernel<<<grid, PM_THREADS, sm, reinterpret_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const __nv_bfloat16*>(q), ldq, T, t0, nq, reinterpret_cast<const __nv_bfloat16*>(keyp),
            key_batch_stride, ldk, O, H, mask, mask_stride, scores, ld_scores, V, 1.0f / sqrtf((float)H));
        return launch_status("ptr_score");
    }
    const size_t smem = (size_t)nq * H * sizeof(float);
    static size_t attr = 48 * 1024;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(ptr_score_kernel<PS_MAXQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("ptr_score attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = smem;
    }
    dim3 grid((O + PS_ROWS - 1) / PS_ROWS, B);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const __nv_bfloat16* qp = reinterpret_cast<const __nv_bfloat16*>(q);
    const __nv_bfloat16* kp = reinterpret_cast<const __nv_bfloat16*>(keyp);
    if (nq == 1)
        launch_pdl(true, ptr_score_kernel<1>, grid, dim3(PS_THREADS), smem, st, qp, ldq, T, t0, nq, kp, key_batch_stride,
                   ldk, O, H, mask, mask_stride, scores, ld_scores, V, sqrtf((float)H));
    else
        ptr_score_kernel<PS_MAXQ><<<grid, PS_THREADS, smem, st>>>(qp, ldq, T, t0, nq, kp, key_batch_stride, ldk, O, H,
                                                                 mask, mask_stride, scores, ld_scores, V, sqrtf((float)H));
    return launch_status("ptr_score");
}

extern "C" int t2s_argmax_feedback(const float* scores, long long ld_scores, int B, int T, int t0, int nt, int N,
                                   long long* prev_inds, int ld_prev, long long* argmax_out, void* stream) {
    if (B <= 0 || nt <= 0 || t0 < 0 || t0 + nt > T) { set_error("argmax_feedback: bad arguments"); return T2S_ERR_SHAPE; }
    launch_pdl(B * nt <= PDL_MAX_ROWS, argmax_feedback_kernel, dim3(B * nt), dim3(256), 0,
               reinterpret_cast<cudaStream_t>(stream), scores, ld_scores, T, t0, nt, N, prev_inds, ld_prev, argmax_out);
    return launch_status("argmax_feedback");
}

extern "C" long long t2s_loss_workspace_bytes(int B, int T) {
    return (long long)(1024 * sizeof(double)) + (long long)B * T * 5 * sizeof(float);
}

extern "C" int t2s_pos_bce_loss(const float* scores, const float* targets, const float* loss_mask, int B, int T, int N,
                                void* workspace, float* out, void* stream) {
    if (B <= 0 || T <= 0 || N <= 0) { set_error("pos_bce_loss: bad shape"); return T2S_ERR_SHAPE; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long rows = (long long)B * T;
    const int grid = (int)(rows < 1024 ? rows : 1024);
    bce_partial_kernel<<<grid, 256, 0, st>>>(scores, targets, loss_mask, rows, N, reinterpret_cast<double*>(workspace));
    bce_final_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const double*>(workspace), grid, loss_mask, rows, out);
    return launch_status("pos_bce_loss");
}

/* row statistics {|ref|^2, |pos|^2, |neg|^2, ref.pos, ref.neg} shared by the InfoNCE forward and backward */
extern "C" int t2s_nce_rowstats(const float* ref, const float* pos, const float* neg, int rows, int N, float* stats,
                                void* stream) {
    if (rows <= 0 || N <= 0) { set_error("nce_rowstats: bad shape"); return T2S_ERR_SHAPE; }
    nce_rowstats_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ref, pos, neg, N, stats);
    return launch_status("nce_rowstats");
}

extern "C" int t2s_info_nce_loss(const float* ref, const float* pos, const float* neg, int B, int T, int N,
                                 float temperature, void* workspace, float* out, void* stream) {
    if (B <= 0 || T <= 0 || N <= 0) { set_error("info_nce_loss: bad shape"); return T2S_ERR_SHAPE; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    float* stats = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 1024 * sizeof(double));
    nce_rowstats_kernel<<<B * T, 256, 0, st>>>(ref, pos, neg, N, stats);
    nce_final_kernel<<<1, 256, 0, st>>>(stats, B, T, temperature, out);
    return launch_status("info_nce_loss");
}
