// K8: row-wise backward kernels of the training step.  All HBM-bound: one warp per 768-wide
// row with 16-byte accesses, per-thread column accumulators for the parameter gradients, one fp32 atomicAdd per
// column per CTA at the end.
//
// What autograd runs in the reference for loss.backward() (pythia/trainers/base_trainer.py:264) through
//   BertLayerNorm / BertSelfOutput / BertOutput (LayerNorm(x + residual))      -> t2s_ln_bwd
//   nn.Linear bias gradients                                                  -> fused column sums / t2s_colsum
//   gelu of BertIntermediate (forward, training mode keeps the pre-activation) -> t2s_gelu_rows
//   x + tanh(enc(x)) of QTV (models/t2s.py:430-432)                            -> tanh flag of t2s_ln_bwd
//   nn.Embedding gradients (word / position / token_type, frame / temporal / track ids)  -> t2s_embed_scatter_add
//   OcrPtrNet matmul (t2s.py:661-666)                                          -> t2s_ptr_score_bwd
//   PrevPredEmbeddings (t2s.py:690-723)                                        -> t2s_prev_embed_bwd
//   LN(linear_ocr_feat) + LN(linear_ocr_bbox(bbox)) (t2s.py:246-252)           -> t2s_ocr_finish_bwd
//   BertEmbeddings (via t2s.py:530)                                            -> t2s_bert_embed_bwd
//   POSBCEWithMaskLoss / InfoNCE (modules/losses.py:329-385)                   -> t2s_pos_bce_loss_bwd / t2s_info_nce_loss_bwd
//   Adam + clip_grad_norm_ (trainers/base_trainer.py:266-269, utils/general.py:32-40)    -> t2s_sumsq / t2s_adam_step
#include "common.cuh"
#include "../../include/t2s_b200.h"

namespace t2s {

constexpr int BW_THREADS = 256;      // 8 warps, one row each per iteration
constexpr int BW_MAXV = 6;           // H <= 768, H % 128 == 0 (the model is built for hidden 768)

struct BwRowMap {           // row = (r / per) * group + off + r % per   (per == 0: identity)
    int per, group, off;
    __device__ __forceinline__ long long operator()(int r) const {
        return per > 0 ? (long long)(r / per) * group + off + (r % per) : (long long)r;
    }
};

template <typename T>
__device__ __forceinline__ float4 bw_load4(const T* p);
template <>
__device__ __forceinline__ float4 bw_load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 bw_load4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    return make_float4(bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y));
}
__device__ __forceinline__ void bw_store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void bw_store4(__nv_bfloat16* p, float4 v) {
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = o;
}
__device__ __forceinline__ float sum4(float4 v) { return (v.x + v.y) + (v.z + v.w); }

// Row statistics + normalised row: xh = (x - mean) * rstd, in place.  Returns rstd.
__device__ __forceinline__ float bw_normalise(float4 (&x)[BW_MAXV], int nv, int H, float eps) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < BW_MAXV; ++i)
        if (i < nv) s += sum4(x[i]);
    const float mean = warp_sum(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < BW_MAXV; ++i)
        if (i < nv) {
            x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
            q += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
        }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + eps);
#pragma unroll
    for (int i = 0; i < BW_MAXV; ++i)
        if (i < nv) { x[i].x *= rstd; x[i].y *= rstd; x[i].z *= rstd; x[i].w *= rstd; }
    return rstd;
}

// LayerNorm backward of one row held in registers.  xh = normalised input (in), dy = upstream gradient (in),
// returns dx in `dy`; accumulates dgamma / dbeta into ag / ab.
__device__ __forceinline__ void bw_ln_row(const float4 (&xh)[BW_MAXV], float4 (&dy)[BW_MAXV], int nv, int H, float rstd,
                                          const float* __restrict__ gamma, int lane, float4 (&ag)[BW_MAXV],
                                          float4 (&ab)[BW_MAXV]) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < BW_MAXV; ++i)
        if (i < nv) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4));
            ag[i].x = fmaf(dy[i].x, xh[i].x, ag[i].x); ag[i].y = fmaf(dy[i].y, xh[i].y, ag[i].y);
            ag[i].z = fmaf(dy[i].z, xh[i].z, ag[i].z); ag[i].w = fmaf(dy[i].w, xh[i].w, ag[i].w);
            ab[i].x += dy[i].x; ab[i].y += dy[i].y; ab[i].z += dy[i].z; ab[i].w += dy[i].w;
            dy[i].x *= g.x; dy[i].y *= g.y; dy[i].z *= g.z; dy[i].w *= g.w;      // g_hat = dy * gamma
            s1 += sum4(dy[i]);
            s2 += (dy[i].x * xh[i].x + dy[i].y * xh[i].y) + (dy[i].z * xh[i].z + dy[i].w * xh[i].w);
        }
    const float m1 = warp_sum(s1) / (float)H, m2 = warp_sum(s2) / (float)H;
#pragma unroll
    for (int i = 0; i < BW_MAXV; ++i)
        if (i < nv) {
            dy[i].x = rstd * (dy[i].x - m1 - xh[i].x * m2); dy[i].y = rstd * (dy[i].y - m1 - xh[i].y * m2);
            dy[i].z = rstd * (dy[i].z - m1 - xh[i].z * m2); dy[i].w = rstd * (dy[i].w - m1 - xh[i].w * m2);
        }
}

// Cross-warp reduction of per-thread column accumulators, then one atomicAdd per column per CTA.
// `red` = [BW_THREADS / 32][H] floats of shared memory.
__device__ __forceinline__ void bw_flush_cols(const float4 (&acc)[BW_MAXV], int nv, int H, float* red, float* __restrict__ dst) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < BW_MAXV; ++i)
        if (i < nv) *reinterpret_cast<float4*>(red + warp * H + (i * 32 + lane) * 4) = acc[i];
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += BW_THREADS) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < BW_THREADS / 32; ++w) s += red[w * H + c];
        atomicAdd(dst + c, s);
    }
}

__device__ __forceinline__ void zero_acc(float4 (&a)[BW_MAXV]) {
#pragma unroll
    for (int i = 0; i < BW_MAXV; ++i) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------- LayerNorm backward
// y = LN(h) (h = pre-LayerNorm sum, saved by the forward); optional out = base + tanh(y) (QTV): dy <- dy (1 - tanh^2 y).
// dh -> `dh` (row-compact, also the gradient of the residual branch and of the preceding Linear's output);
// dgamma / dbeta / dbias (column sums of dh, the preceding Linear's bias gradient) are accumulated with atomics.
template <typename TH, typename TD, typename TO>
__global__ void __launch_bounds__(BW_THREADS)
ln_bwd_kernel(const TH* __restrict__ h, long long ldh, const TD* __restrict__ dy, long long lddy, BwRowMap dy_map,
              const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int rows, int H, int tanh_out,
              TO* __restrict__ dh, long long lddh, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ dbias, TO* __restrict__ dh_drop, DropCfg drop) {
    extern __shared__ float red[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = H / 128;
    float4 ag[BW_MAXV], ab[BW_MAXV], ac[BW_MAXV];
    zero_acc(ag); zero_acc(ab); zero_acc(ac);
    for (int row = blockIdx.x * (BW_THREADS / 32) + warp; row < rows; row += gridDim.x * (BW_THREADS / 32)) {
        float4 xh[BW_MAXV], d[BW_MAXV];
        const long long drow = dy_map(row);
#pragma unroll
        for (int i = 0; i < BW_MAXV; ++i)
            if (i < nv) {
                const int e = (i * 32 + lane) * 4;
                xh[i] = bw_load4<TH>(h + (long long)row * ldh + e);
                d[i] = bw_load4<TD>(dy + drow * lddy + e);
            }
        const float rstd = bw_normalise(xh, nv, H, eps);
        if (tanh_out) {
#pragma unroll
            for (int i = 0; i < BW_MAXV; ++i)
                if (i < nv) {
                    const int e = (i * 32 + lane) * 4;
                    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + e));
                    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + e));
                    float t;
                    t = tanhf(fmaf(xh[i].x, g.x, b.x)); d[i].x *= 1.f - t * t;
                    t = tanhf(fmaf(xh[i].y, g.y, b.y)); d[i].y *= 1.f - t * t;
                    t = tanhf(fmaf(xh[i].z, g.z, b.z)); d[i].z *= 1.f - t * t;
                    t = tanhf(fmaf(xh[i].w, g.w, b.w)); d[i].w *= 1.f - t * t;
                }
        }
        bw_ln_row(xh, d, nv, H, rstd, gamma, lane, ag, ab);
#pragma unroll
        for (int i = 0; i < BW_MAXV; ++i)
            if (i < nv) {
                const int e = (i * 32 + lane) * 4;
                bw_store4(dh + (long long)row * lddh + e, d[i]);
                if (dh_drop) {
                    // the forward was LN(dropout(Linear(.)) + residual): `dh` is the residual branch's gradient, the
                    // Linear's output (and its bias) get the masked one
                    const float4 m = drop_mask4(drop, row, H, e);
                    d[i].x *= m.x; d[i].y *= m.y; d[i].z *= m.z; d[i].w *= m.w;
                    bw_store4(dh_drop + (long long)row * lddh + e, d[i]);
                }
                ac[i].x += d[i].x; ac[i].y += d[i].y; ac[i].z += d[i].z; ac[i].w += d[i].w;
            }
    }
    bw_flush_cols(ag, nv, H, red, dgamma);
    bw_flush_cols(ab, nv, H, red, dbeta);
    if (dbias) bw_flush_cols(ac, nv, H, red, dbias);
}

// ------------------------------------------------------------------------------- column sums (bias gradients)
// dst[c] += sum_r x[r, c].  bf16 path: CTA = 8 warps x 256 columns; a warp reads 512 contiguous bytes of a row per
// load (16 B per lane), four rows in flight; per-thread fp32 partials for its 8 columns, cross-warp reduction in
// shared memory, one atomicAdd per column per CTA.  (A thread-per-column loop with one load in flight ran at
// ~1.5 TB/s-equivalent latency, 200 us per call; this streams at HBM rate.)
constexpr int CS_ROWS_PER_CTA = 256;
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int rows, int N, float* __restrict__ dst) {
    __shared__ float red[8][256 + 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * 256 + lane * 8;
    const int r_begin = blockIdx.y * CS_ROWS_PER_CTA, r_end = min(rows, r_begin + CS_ROWS_PER_CTA);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c0 + 8 <= N) {
        for (int r = r_begin + warp; r < r_end; r += 32) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = r + 8 * u;
                v[u] = rr < r_end ? *reinterpret_cast<const uint4*>(x + (long long)rr * ldx + c0) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc[0] += bf16lo(v[u].x); acc[1] += bf16hi(v[u].x); acc[2] += bf16lo(v[u].y); acc[3] += bf16hi(v[u].y);
                acc[4] += bf16lo(v[u].z); acc[5] += bf16hi(v[u].z); acc[6] += bf16lo(v[u].w); acc[7] += bf16hi(v[u].w);
            }
        }
    } else {
        for (int r = r_begin + warp; r < r_end; r += 8)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (c0 + j < N) acc[j] += __bfloat162float(x[(long long)r * ldx + c0 + j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < N) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        atomicAdd(dst + c, t);
    }
}
// fp32 rows (small: loss gradients over the decoder rows): thread owns a column, rows strided over blockIdx.y
__global__ void __launch_bounds__(256)
colsum_f32_kernel(const float* __restrict__ x, long long ldx, int rows, int N, float* __restrict__ dst) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= N) return;
    float s = 0.f;
    for (int r = blockIdx.y; r < rows; r += gridDim.y) s += x[(long long)r * ldx + c];
    atomicAdd(dst + c, s);
}

// ------------------------------------------------------------------------------- operand copies of the weights
// After an optimizer step every GEMM operand derived from the fp32 parameters is stale: the bf16 copies (answer
// transformer, heads), the bf16 hi|lo splits (grounding chain), the transposed bf16 copies the dgrad GEMMs read.  The
// host used to rebuild them with ~300 small tensor ops (3.3 ms per step, host bound); this kernel rewrites them all in
// place from a job table: job = one parameter matrix -> one destination.  One CTA per 32 x 32 tile of a source matrix.
struct RepackJob {          // 48 bytes; mirrored by vitxt_gqa_b200/train.py (struct format "qqqiiiiii4x")
    long long src_off;      // first element of the fp32 [rows, cols] matrix in the flat parameter buffer
    long long dst;          // destination pointer
    long long ld_dst;       // row pitch of the destination in elements
    int rows, cols;
    int mode;               // 0 bf16 copy, 1 bf16 hi|lo split (lo at column k_pad), 2 bf16 transposed, 3 fp32 copy
    int k_pad;              // mode 1: columns [cols, k_pad) of both halves are zero filled
    int tile0;              // index of this job's first tile in the launch
    int tiles_x;            // tiles along the columns (of max(cols, k_pad))
};

__global__ void __launch_bounds__(256)
repack_kernel(const float* __restrict__ flat, const RepackJob* __restrict__ jobs, int n_jobs) {
    __shared__ float tile[32][33];
    // binary search: the job whose tile range holds blockIdx.x
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].tile0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const RepackJob j = jobs[lo];
    const int t = blockIdx.x - j.tile0;
    const int r0 = (t / j.tiles_x) * 32, c0 = (t % j.tiles_x) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    const float* src = flat + j.src_off;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + tx;
        tile[ty + 8 * i][tx] = (r < j.rows && c < j.cols) ? src[(long long)r * j.cols + c] : 0.f;
    }
    __syncthreads();
    if (j.mode == 2) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(j.dst);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = c0 + ty + 8 * i, r = r0 + tx;          // destination row = source column
            if (c < j.cols && r < j.rows) dst[(long long)c * j.ld_dst + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + tx;
        if (r >= j.rows) continue;
        const float v = tile[ty + 8 * i][tx];
        if (j.mode == 0) {
            if (c < j.cols) reinterpret_cast<__nv_bfloat16*>(j.dst)[(long long)r * j.ld_dst + c] = __float2bfloat16_rn(v);
        } else if (j.mode == 1) {
            if (c < j.k_pad) {
                __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(j.dst) + (long long)r * j.ld_dst;
                const __nv_bfloat16 h = __float2bfloat16_rn(v);              // v == 0 in the padding columns
                d[c] = h;
                d[j.k_pad + c] = __float2bfloat16_rn(v - __bfloat162float(h));
            }
        } else {
            if (c < j.cols) reinterpret_cast<float*>(j.dst)[(long long)r * j.ld_dst + c] = v;
        }
    }
}

// ------------------------------------------------------------------------------- row sums / casts
// out[map(r), :] (+)= a[r] + b[r] + c[r]   (inputs bf16, fp32 output); used to gather the gradient of the joint
// embedding from the three grounding variants
__global__ void __launch_bounds__(256)
rows_add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                const __nv_bfloat16* __restrict__ c, long long ldi, int rows, int H, float* __restrict__ out,
                long long ldo, BwRowMap map, int accumulate) {
    const long long n4 = (long long)rows * (H / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / (H / 4)), col = (int)(i % (H / 4)) * 4;
        float4 v = bw_load4<__nv_bfloat16>(a + (long long)row * ldi + col);
        if (b) { const float4 t = bw_load4<__nv_bfloat16>(b + (long long)row * ldi + col); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
        if (c) { const float4 t = bw_load4<__nv_bfloat16>(c + (long long)row * ldi + col); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
        float* o = out + map(row) * ldo + col;
        if (accumulate) { const float4 t = *reinterpret_cast<const float4*>(o); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
        *reinterpret_cast<float4*>(o) = v;
    }
}

// GELU of saved pre-activations: u (bf16) -> bf16, or u (fp32) -> bf16 hi|lo (lo at column lo_off)
template <typename T>
__global__ void __launch_bounds__(256)
gelu_rows_kernel(const T* __restrict__ u, long long ldu, int rows, int N, __nv_bfloat16* __restrict__ out, long long ldo,
                 int lo_off) {
    const long long n4 = (long long)rows * (N / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / (N / 4)), col = (int)(i % (N / 4)) * 4;
        float4 v = bw_load4<T>(u + (long long)row * ldu + col);
        v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
        __nv_bfloat16* o = out + (long long)row * ldo + col;
        uint2 hi;
        hi.x = pack_bf16x2(v.x, v.y);
        hi.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(o) = hi;
        if (lo_off > 0) {
            uint2 lo;
            lo.x = pack_bf16x2(v.x - bf16lo(hi.x), v.y - bf16hi(hi.x));
            lo.y = pack_bf16x2(v.z - bf16lo(hi.y), v.w - bf16hi(hi.y));
            *reinterpret_cast<uint2*>(o + lo_off) = lo;
        }
    }
}

// ------------------------------------------------------------------------------- embedding gradients
// table[ids[r], 0..d) += src[r, c0 .. c0+d)   (nn.Embedding backward; rows with ids[r] == pad_id are skipped when
// pad_id >= 0, as padding_idx does for word_embeddings)
template <typename T>
__global__ void __launch_bounds__(256)
embed_scatter_add_kernel(const T* __restrict__ src, long long lds, int c0, int d, const long long* __restrict__ ids,
                         int rows, long long pad_id, float* __restrict__ table, long long ldt) {
    const long long n = (long long)rows * d;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / d), c = (int)(i % d);
        const long long id = ids[row];
        if (id == pad_id) continue;
        float v;
        if constexpr (sizeof(T) == 2) v = __bfloat162float(src[(long long)row * lds + c0 + c]);
        else v = src[(long long)row * lds + c0 + c];
        atomicAdd(table + id * ldt + c, v);
    }
}

// ------------------------------------------------------------------------------- pointer-network score backward
// scores[b,t,V+o] = q[b,t].k[b,o] / sqrt(H) + mask   =>   dq[b,t] = sum_o dS[b,t,V+o] k[b,o] / sqrt(H),
// dk[b,o] = sum_t dS[b,t,V+o] q[b,t] / sqrt(H).  One CTA per (sample, 32-key slab) for dk, per sample for dq.
constexpr int PB_T = 16;
__global__ void __launch_bounds__(256)
ptr_score_bwd_dk_kernel(const float* __restrict__ dS, long long ld_s, int T, int V, const __nv_bfloat16* __restrict__ q,
                        long long ldq, int O, int H, __nv_bfloat16* __restrict__ dk, long long dk_batch_stride,
                        long long lddk, float inv) {
    extern __shared__ float qs[];          // [T][H]
    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < T * H; i += 256) qs[i] = __bfloat162float(q[((long long)b * T + i / H) * ldq + i % H]) * inv;
    __syncthreads();
    for (int o = blockIdx.x * 32 + warp; o < min(O, (int)(blockIdx.x + 1) * 32); o += 8) {
        float g[PB_T];
#pragma unroll
        for (int t = 0; t < PB_T; ++t) g[t] = t < T ? dS[((long long)b * T + t) * ld_s + V + o] : 0.f;
        for (int d = lane * 2; d < H; d += 64) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int t = 0; t < PB_T; ++t)
                if (t < T) { a0 = fmaf(g[t], qs[t * H + d], a0); a1 = fmaf(g[t], qs[t * H + d + 1], a1); }
            *reinterpret_cast<uint32_t*>(dk + (long long)b * dk_batch_stride + (long long)o * lddk + d) = pack_bf16x2(a0, a1);
        }
    }
}
__global__ void __launch_bounds__(256)
ptr_score_bwd_dq_kernel(const float* __restrict__ dS, long long ld_s, int T, int V, const __nv_bfloat16* __restrict__ k,
                        long long k_batch_stride, long long ldk, int O, int H, __nv_bfloat16* __restrict__ dq,
                        long long lddq, float inv) {
    // CTA = (decoder row t, sample b).  Warp w takes keys o = w, w + 8, ...; a lane holds 8 columns per 256-column
    // slab (16-byte loads of the key rows, four keys in flight); partial rows are reduced across warps in shared memory.
    extern __shared__ float gs[];          // [O] scaled dS row, then [8][H] partials
    float* part = gs + ((O + 3) & ~3);
    const int b = blockIdx.y, t = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int o = tid; o < O; o += 256) gs[o] = dS[((long long)b * T + t) * ld_s + V + o] * inv;
    __syncthreads();
    const __nv_bfloat16* kb = k + (long long)b * k_batch_stride;
    for (int d0 = 0; d0 < H; d0 += 256) {
        const int d = d0 + lane * 8;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (d + 8 <= H) {
            for (int o = warp; o < O; o += 32) {
                uint4 v[4];
                float g[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int oo = o + 8 * u;
                    const bool ok = oo < O;
                    v[u] = ok ? *reinterpret_cast<const uint4*>(kb + (long long)oo * ldk + d) : make_uint4(0u, 0u, 0u, 0u);
                    g[u] = ok ? gs[oo] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    acc[0] = fmaf(g[u], bf16lo(v[u].x), acc[0]); acc[1] = fmaf(g[u], bf16hi(v[u].x), acc[1]);
                    acc[2] = fmaf(g[u], bf16lo(v[u].y), acc[2]); acc[3] = fmaf(g[u], bf16hi(v[u].y), acc[3]);
                    acc[4] = fmaf(g[u], bf16lo(v[u].z), acc[4]); acc[5] = fmaf(g[u], bf16hi(v[u].z), acc[5]);
                    acc[6] = fmaf(g[u], bf16lo(v[u].w), acc[6]); acc[7] = fmaf(g[u], bf16hi(v[u].w), acc[7]);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) part[warp * H + d + j] = acc[j];
        }
    }
    __syncthreads();
    for (int d = tid * 2; d < H; d += 512) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a0 += part[w * H + d]; a1 += part[w * H + d + 1]; }
        *reinterpret_cast<uint32_t*>(dq + ((long long)b * T + t) * lddq + d) = pack_bf16x2(a0, a1);
    }
}

// ------------------------------------------------------------------------------- PrevPredEmbeddings backward
// dec[b,t] = LN_src(src row) + LN_emb(pos[t] + type[is_ocr]); one warp per decoder row.
__global__ void __launch_bounds__(BW_THREADS)
prev_embed_bwd_kernel(const __nv_bfloat16* __restrict__ dx, long long lddx, const long long* __restrict__ prev_inds,
                      int ld_prev, int B, int T, int V, int H, const float* __restrict__ ans_w,
                      const float* __restrict__ ocr_emb, long long ocr_batch_stride, long long ld_ocr,
                      const float* __restrict__ pos_emb, const float* __restrict__ type_emb,
                      const float* __restrict__ ans_g, const float* __restrict__ ocr_g, const float* __restrict__ emb_g,
                      float eps, float* __restrict__ d_ans_w, float* __restrict__ d_ocr_emb,
                      float* __restrict__ d_pos, float* __restrict__ d_type, float* __restrict__ d_ans_g,
                      float* __restrict__ d_ans_b, float* __restrict__ d_ocr_g, float* __restrict__ d_ocr_b,
                      float* __restrict__ d_emb_g, float* __restrict__ d_emb_b, int n_ocr) {
    extern __shared__ float red[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = H / 128;
    float4 g_ans[BW_MAXV], b_ans[BW_MAXV], g_ocr[BW_MAXV], b_ocr[BW_MAXV], g_emb[BW_MAXV], b_emb[BW_MAXV];
    zero_acc(g_ans); zero_acc(b_ans); zero_acc(g_ocr); zero_acc(b_ocr); zero_acc(g_emb); zero_acc(b_emb);
    for (int w = blockIdx.x * (BW_THREADS / 32) + warp; w < B * T; w += gridDim.x * (BW_THREADS / 32)) {
        const int b = w / T, t = w % T;
        const long long idx = clamp_index(prev_inds[(long long)b * ld_prev + t], (long long)V + n_ocr);   // as prev_embed
        const bool is_ocr = idx >= V;
        const float* src = is_ocr ? ocr_emb + (long long)b * ocr_batch_stride + (idx - V) * ld_ocr : ans_w + idx * H;
        const float* ty = type_emb + (is_ocr ? H : 0);
        float4 r[BW_MAXV], e[BW_MAXV], d1[BW_MAXV], d2[BW_MAXV];
#pragma unroll
        for (int i = 0; i < BW_MAXV; ++i)
            if (i < nv) {
                const int c = (i * 32 + lane) * 4;
                r[i] = *reinterpret_cast<const float4*>(src + c);
                const float4 p = *reinterpret_cast<const float4*>(pos_emb + (long long)t * H + c);
                const float4 q = *reinterpret_cast<const float4*>(ty + c);
                e[i] = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
                d1[i] = bw_load4<__nv_bfloat16>(dx + (long long)w * lddx + c);
                d2[i] = d1[i];
            }
        const float rs1 = bw_normalise(r, nv, H, eps);
        const float rs2 = bw_normalise(e, nv, H, eps);
        if (is_ocr) bw_ln_row(r, d1, nv, H, rs1, ocr_g, lane, g_ocr, b_ocr);
        else bw_ln_row(r, d1, nv, H, rs1, ans_g, lane, g_ans, b_ans);
        bw_ln_row(e, d2, nv, H, rs2, emb_g, lane, g_emb, b_emb);
        float* dsrc = is_ocr ? d_ocr_emb + (long long)b * ocr_batch_stride + (idx - V) * ld_ocr : d_ans_w + idx * H;
        float* dty = d_type + (is_ocr ? H : 0);
#pragma unroll
        for (int i = 0; i < BW_MAXV; ++i)
            if (i < nv) {
                const int c = (i * 32 + lane) * 4;
                atomicAdd(dsrc + c, d1[i].x); atomicAdd(dsrc + c + 1, d1[i].y);
                atomicAdd(dsrc + c + 2, d1[i].z); atomicAdd(dsrc + c + 3, d1[i].w);
                float* dp = d_pos + (long long)t * H + c;
                atomicAdd(dp, d2[i].x); atomicAdd(dp + 1, d2[i].y); atomicAdd(dp + 2, d2[i].z); atomicAdd(dp + 3, d2[i].w);
                atomicAdd(dty + c, d2[i].x); atomicAdd(dty + c + 1, d2[i].y);
                atomicAdd(dty + c + 2, d2[i].z); atomicAdd(dty + c + 3, d2[i].w);
            }
    }
    bw_flush_cols(g_ans, nv, H, red, d_ans_g); bw_flush_cols(b_ans, nv, H, red, d_ans_b);
    bw_flush_cols(g_ocr, nv, H, red, d_ocr_g); bw_flush_cols(b_ocr, nv, H, red, d_ocr_b);
    bw_flush_cols(g_emb, nv, H, red, d_emb_g); bw_flush_cols(b_emb, nv, H, red, d_emb_b);
}

// ------------------------------------------------------------------------------- OCR encoder tail backward
// out = LN1(h) + LN2(c), c = W2.bbox + b2 (K = 4).  dout rows are gathered from the joint gradient (dy_map).
// Writes dh (bf16, gradient of linear_ocr_feat_to_mmt_in's output); accumulates LN1 / LN2 parameter gradients, the
// bias of linear_ocr_feat (column sums of dh), and dW2 [H,4], db2.
__global__ void __launch_bounds__(BW_THREADS)
ocr_finish_bwd_kernel(const float* __restrict__ h, long long ldh, const float* __restrict__ bbox,
                      const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ g1,
                      const float* __restrict__ g2, float eps, int rows, int H, const float* __restrict__ dout,
                      long long ldd, BwRowMap dy_map, __nv_bfloat16* __restrict__ dh, long long lddh,
                      float* __restrict__ dc_out, long long lddc,
                      float* __restrict__ dg1, float* __restrict__ db1, float* __restrict__ dg2, float* __restrict__ db2ln,
                      float* __restrict__ dbias1, float* __restrict__ db2) {
    extern __shared__ float red[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = H / 128;
    float4 ag1[BW_MAXV], ab1[BW_MAXV], ag2[BW_MAXV], ab2[BW_MAXV], ac1[BW_MAXV], ac2[BW_MAXV];
    zero_acc(ag1); zero_acc(ab1); zero_acc(ag2); zero_acc(ab2); zero_acc(ac1); zero_acc(ac2);
    for (int row = blockIdx.x * (BW_THREADS / 32) + warp; row < rows; row += gridDim.x * (BW_THREADS / 32)) {
        const float4 bx = *reinterpret_cast<const float4*>(bbox + (long long)row * 4);
        float4 a[BW_MAXV], c[BW_MAXV], d1[BW_MAXV], d2[BW_MAXV];
        const long long drow = dy_map(row);
#pragma unroll
        for (int i = 0; i < BW_MAXV; ++i)
            if (i < nv) {
                const int e = (i * 32 + lane) * 4;
                a[i] = *reinterpret_cast<const float4*>(h + (long long)row * ldh + e);
                float r[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 w = __ldg(reinterpret_cast<const float4*>(w2 + (long long)(e + j) * 4));
                    r[j] = fmaf(bx.w, w.w, fmaf(bx.z, w.z, fmaf(bx.y, w.y, bx.x * w.x))) + __ldg(b2 + e + j);
                }
                c[i] = make_float4(r[0], r[1], r[2], r[3]);
                d1[i] = *reinterpret_cast<const float4*>(dout + drow * ldd + e);
                d2[i] = d1[i];
            }
        const float rs1 = bw_normalise(a, nv, H, eps);
        const float rs2 = bw_normalise(c, nv, H, eps);
        bw_ln_row(a, d1, nv, H, rs1, g1, lane, ag1, ab1);
        bw_ln_row(c, d2, nv, H, rs2, g2, lane, ag2, ab2);
#pragma unroll
        for (int i = 0; i < BW_MAXV; ++i)
            if (i < nv) {
                const int e = (i * 32 + lane) * 4;
                ac1[i].x += d1[i].x; ac1[i].y += d1[i].y; ac1[i].z += d1[i].z; ac1[i].w += d1[i].w;
                ac2[i].x += d2[i].x; ac2[i].y += d2[i].y; ac2[i].z += d2[i].z; ac2[i].w += d2[i].w;
                bw_store4(dh + (long long)row * lddh + e, d1[i]);
                bw_store4(dc_out + (long long)row * lddc + e, d2[i]);      // gradient of linear_ocr_bbox's output
            }
    }
    bw_flush_cols(ag1, nv, H, red, dg1); bw_flush_cols(ab1, nv, H, red, db1);
    bw_flush_cols(ag2, nv, H, red, dg2); bw_flush_cols(ab2, nv, H, red, db2ln);
    bw_flush_cols(ac1, nv, H, red, dbias1); bw_flush_cols(ac2, nv, H, red, db2);
}

// dW[e, k] += sum_r dc[r, e] x[r, k] for a K = 4 input (linear_ocr_bbox_to_mmt_in): thread owns column e
__global__ void __launch_bounds__(256)
k4_wgrad_kernel(const float* __restrict__ dc, long long lddc, const float* __restrict__ x, int rows, int H,
                float* __restrict__ dw) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= H) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        const float g = dc[(long long)r * lddc + e];
        const float4 bx = __ldg(reinterpret_cast<const float4*>(x + (long long)r * 4));
        acc.x = fmaf(g, bx.x, acc.x); acc.y = fmaf(g, bx.y, acc.y); acc.z = fmaf(g, bx.z, acc.z); acc.w = fmaf(g, bx.w, acc.w);
    }
    float* p = dw + (long long)e * 4;
    atomicAdd(p, acc.x); atomicAdd(p + 1, acc.y); atomicAdd(p + 2, acc.z); atomicAdd(p + 3, acc.w);
}

// ------------------------------------------------------------------------------- BertEmbeddings backward
// e = word[id] + pos[p] + type[0]; y = LN(e).  dy rows gathered from the first layer's input gradient (bf16).
__global__ void __launch_bounds__(BW_THREADS)
bert_embed_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const long long* __restrict__ ids, int rows,
                      int L, int H, const float* __restrict__ word, const float* __restrict__ pos,
                      const float* __restrict__ type0, const float* __restrict__ gamma, float eps,
                      float* __restrict__ d_word, float* __restrict__ d_pos, float* __restrict__ d_type,
                      float* __restrict__ dgamma, float* __restrict__ dbeta) {
    extern __shared__ float red[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = H / 128;
    float4 ag[BW_MAXV], ab[BW_MAXV], at[BW_MAXV];
    zero_acc(ag); zero_acc(ab); zero_acc(at);
    for (int row = blockIdx.x * (BW_THREADS / 32) + warp; row < rows; row += gridDim.x * (BW_THREADS / 32)) {
        const long long id = ids[row];
        const int p = row % L;
        float4 x[BW_MAXV], d[BW_MAXV];
#pragma unroll
        for (int i = 0; i < BW_MAXV; ++i)
            if (i < nv) {
                const int e = (i * 32 + lane) * 4;
                const float4 w = *reinterpret_cast<const float4*>(word + id * H + e);
                const float4 q = *reinterpret_cast<const float4*>(pos + (long long)p * H + e);
                const float4 t = *reinterpret_cast<const float4*>(type0 + e);
                x[i] = make_float4((w.x + q.x) + t.x, (w.y + q.y) + t.y, (w.z + q.z) + t.z, (w.w + q.w) + t.w);
                d[i] = bw_load4<__nv_bfloat16>(dy + (long long)row * lddy + e);
            }
        const float rstd = bw_normalise(x, nv, H, eps);
        bw_ln_row(x, d, nv, H, rstd, gamma, lane, ag, ab);
#pragma unroll
        for (int i = 0; i < BW_MAXV; ++i)
            if (i < nv) {
                const int e = (i * 32 + lane) * 4;
                at[i].x += d[i].x; at[i].y += d[i].y; at[i].z += d[i].z; at[i].w += d[i].w;
                if (id != 0) {      // padding_idx = 0 of word_embeddings receives no gradient
                    float* dw = d_word + id * H + e;
                    atomicAdd(dw, d[i].x); atomicAdd(dw + 1, d[i].y); atomicAdd(dw + 2, d[i].z); atomicAdd(dw + 3, d[i].w);
                }
                float* dp = d_pos + (long long)p * H + e;
                atomicAdd(dp, d[i].x); atomicAdd(dp + 1, d[i].y); atomicAdd(dp + 2, d[i].z); atomicAdd(dp + 3, d[i].w);
            }
    }
    bw_flush_cols(ag, nv, H, red, dgamma);
    bw_flush_cols(ab, nv, H, red, dbeta);
    bw_flush_cols(at, nv, H, red, d_type);      // token_type row 0
}

// ------------------------------------------------------------------------------- loss backward
// d/dscores of w * sum(BCEWithLogits(x, z) * mask) / max(sum(mask), 1): (sigmoid(x) - z) * mask * w / count
__global__ void __launch_bounds__(256)
bce_bwd_kernel(const float* __restrict__ scores, const float* __restrict__ targets, const float* __restrict__ loss_mask,
               long long rows, int N, const float* __restrict__ grad_out, float* __restrict__ dscores, int accumulate) {
    __shared__ float s_scale;
    if (threadIdx.x == 0) {
        float c = 0.f;
        for (long long r = 0; r < rows; ++r) c += loss_mask[r];
        s_scale = grad_out[0] / fmaxf(c, 1.f);
    }
    __syncthreads();
    const float scale = s_scale;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const float m = loss_mask[r] * scale;
        const float* x = scores + r * N;
        const float* z = targets + r * N;
        float* o = dscores + r * N;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const float g = (1.f / (1.f + expf(-x[i])) - z[i]) * m;
            o[i] = accumulate ? o[i] + g : g;
        }
    }
}

// InfoNCE backward.  With a = ref, p = pos, n = neg rows (b,t) and the forward's row statistics
// stats[row] = {|a|^2, |p|^2, |n|^2, a.p, a.n}: the unit rows are A = a/|a| etc, the flattened [T*N] vectors have
// norms sqrt(QQ) etc (sums over t of |A_t|^2 = number of non-zero rows), cos_p = (sum_t A_t.P_t) / (sqrt(QQ) sqrt(PP)).
// loss_b = -log softmax([cos_p, cos_n] / tau)[0].  Per sample: dL/dcos_p = (s_p - 1) / tau, dL/dcos_n = s_n / tau.
// coef[b] = {dcp, dcn, QP, QN, QQ, PP, NN} is prepared by nce_coef_kernel.
__global__ void nce_coef_kernel(const float* __restrict__ stats, int B, int T, float temperature,
                                const float* __restrict__ grad_out, float* __restrict__ coef) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float qq = 0.f, pp = 0.f, nn = 0.f, qp = 0.f, qn = 0.f;
    for (int t = 0; t < T; ++t) {
        const float* s = stats + ((long long)b * T + t) * 5;
        const float nr = fmaxf(sqrtf(s[0]), 1e-12f), np = fmaxf(sqrtf(s[1]), 1e-12f), ng = fmaxf(sqrtf(s[2]), 1e-12f);
        qq += s[0] / (nr * nr); pp += s[1] / (np * np); nn += s[2] / (ng * ng);
        qp += s[3] / (nr * np); qn += s[4] / (nr * ng);
    }
    const float nq = fmaxf(sqrtf(qq), 1e-8f), npp = fmaxf(sqrtf(pp), 1e-8f), nnn = fmaxf(sqrtf(nn), 1e-8f);
    const float cp = qp / (nq * npp), cn = qn / (nq * nnn);
    const float lp = cp / temperature, ln = cn / temperature;
    const float mx = fmaxf(lp, ln);
    const float ep = expf(lp - mx), en = expf(ln - mx);
    const float sp = ep / (ep + en), sn = en / (ep + en);
    const float g = grad_out[0] / (float)B;
    float* c = coef + (long long)b * 8;
    c[0] = g * (sp - 1.f) / temperature;       // dL/dcos_p
    c[1] = g * sn / temperature;               // dL/dcos_n
    c[2] = cp; c[3] = cn; c[4] = nq; c[5] = npp; c[6] = nnn; c[7] = 0.f;
}
// One CTA per row (b,t).  Flat unit vectors: Aflat = concat_t A_t; cos_p = Aflat.Pflat / (nq np).
//   dcos_p/dA_t = P_t / (nq np) - cos_p A_t / nq^2,   dcos_p/dP_t = A_t / (nq np) - cos_p P_t / np^2   (same for n)
// then through the row normalisation A_t = a_t / |a_t|:  da_t = (G - (G.A_t) A_t) / |a_t|.
__global__ void __launch_bounds__(256)
nce_bwd_kernel(const float* __restrict__ ref, const float* __restrict__ pos, const float* __restrict__ neg, int T, int N,
               const float* __restrict__ stats, const float* __restrict__ coef, float* __restrict__ dref,
               float* __restrict__ dpos, float* __restrict__ dneg, int accumulate) {
    __shared__ float redbuf[33];
    const long long r = blockIdx.x;
    const int b = (int)(r / T);
    const float* s = stats + r * 5;
    const float* c = coef + (long long)b * 8;
    const float na = fmaxf(sqrtf(s[0]), 1e-12f), np = fmaxf(sqrtf(s[1]), 1e-12f), nn = fmaxf(sqrtf(s[2]), 1e-12f);
    const float dcp = c[0], dcn = c[1], cp = c[2], cn = c[3], nq = c[4], npp = c[5], nnn = c[6];
    const float* a = ref + r * N;
    const float* p = pos + r * N;
    const float* n = neg + r * N;
    // G_a = dcp (P/(nq np) - cp A/nq^2) + dcn (Nn/(nq nn) - cn A/nq^2);  G_p = dcp (A/(nq np) - cp P/np^2);  G_n likewise
    const float kap = dcp / (nq * npp), kan = dcn / (nq * nnn), kaa = -(dcp * cp + dcn * cn) / (nq * nq);
    const float kpp = -dcp * cp / (npp * npp), knn = -dcn * cn / (nnn * nnn);
    // projections G.A etc from the row statistics: A.A = s0/na^2, A.P = s3/(na np), ...
    const float AA = s[0] / (na * na), PP = s[1] / (np * np), NN = s[2] / (nn * nn);
    const float AP = s[3] / (na * np), AN = s[4] / (na * nn);
    const float ga_A = kap * AP + kan * AN + kaa * AA;        // G_a . A
    const float gp_P = kap * AP + kpp * PP;                   // G_p . P
    const float gn_N = kan * AN + knn * NN;                   // G_n . N
    (void)redbuf;
    float* da = dref + r * N;
    float* dp = dpos + r * N;
    float* dn = dneg + r * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float A = a[i] / na, P = p[i] / np, Nn = n[i] / nn;
        const float Ga = kap * P + kan * Nn + kaa * A;
        const float Gp = kap * A + kpp * P;
        const float Gn = kan * A + knn * Nn;
        const float va = (Ga - ga_A * A) / na, vp = (Gp - gp_P * P) / np, vn = (Gn - gn_N * Nn) / nn;
        if (accumulate) { da[i] += va; dp[i] += vp; dn[i] += vn; }
        else { da[i] = va; dp[i] = vp; dn[i] = vn; }
    }
}

// ------------------------------------------------------------------------------- optimizer
// partial[blockIdx] = sum of squares of a flat fp32 buffer (clip_grad_norm_, fixed-order two-stage reduction)
__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float* __restrict__ g, long long n, double* __restrict__ partial) {
    __shared__ float red[33];
    float s = 0.f;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) s = fmaf(g[i], g[i], s);
    s = block_sum(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = (double)s;
}
__global__ void __launch_bounds__(256)
sumsq_final_kernel(const double* __restrict__ partial, int n, float* __restrict__ out) {
    __shared__ double sd[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    sd[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) sd[threadIdx.x] += sd[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)sd[0];
}
// torch.optim.Adam (weight_decay 0, amsgrad off) over a flat fp32 parameter range, with the gradient scaled by
// clip = min(1, max_norm / (sqrt(sumsq) + 1e-6)) as clip_grad_norm_ does (max_norm <= 0: no clipping) and by
// `grad_scale` (1 / world size after a summing all-reduce).
__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long n, float lr, float beta1, float beta2, float eps, float bc1, float bc2,
                 const float* __restrict__ sumsq, float max_norm, float grad_scale) {
    float clip = grad_scale;
    if (max_norm > 0.f && sumsq) {
        const float norm = sqrtf(sumsq[0]) * grad_scale;
        clip *= fminf(1.f, max_norm / (norm + 1e-6f));
    }
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float gi = g[i] * clip;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
        p[i] -= (lr / bc1) * (mi / denom);
    }
}

static inline bool bw_h_ok(int H) { return H % 128 == 0 && H / 128 <= BW_MAXV && H > 0; }
static inline int bw_row_grid(int rows) {
    int g = (rows + BW_THREADS / 32 - 1) / (BW_THREADS / 32);
    const int cap = num_sms() * 4;
    return g < cap ? (g < 1 ? 1 : g) : cap;
}
static inline int flat_grid(long long n) {
    long long g = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace t2s

using namespace t2s;

extern "C" int t2s_nce_rowstats(const float* ref, const float* pos, const float* neg, int rows, int N, float* stats,
                                void* stream);      // scores_loss.cu

static int ln_bwd_entry(const void* h, int h_bf16, long long ldh, const void* dy, int dy_bf16, long long lddy,
                        int dy_rows_per_group, int dy_group_rows, int dy_row_off, const float* gamma,
                        const float* beta, float eps, int rows, int H, int tanh_out, void* dh, int dh_bf16,
                        long long lddh, float* dgamma, float* dbeta, float* dbias, void* dh_drop, DropCfg drop,
                        void* stream) {
    if (!bw_h_ok(H) || rows <= 0) { set_error("ln_bwd: H %d must be a multiple of 128 <= 1024", H); return T2S_ERR_SHAPE; }
    if (!dgamma || !dbeta || !dh) { set_error("ln_bwd: missing output"); return T2S_ERR_ARG; }
    BwRowMap map{dy_rows_per_group, dy_group_rows, dy_row_off};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = bw_row_grid(rows);
    const size_t smem = (size_t)(BW_THREADS / 32) * H * sizeof(float);
#define T2S_LNB(TH, TD, TO)                                                                                          \
    ln_bwd_kernel<TH, TD, TO><<<grid, BW_THREADS, smem, st>>>(reinterpret_cast<const TH*>(h), ldh,                  \
        reinterpret_cast<const TD*>(dy), lddy, map, gamma, beta, eps, rows, H, tanh_out, reinterpret_cast<TO*>(dh), \
        lddh, dgamma, dbeta, dbias, reinterpret_cast<TO*>(dh_drop), drop)
    typedef __nv_bfloat16 bf;
    const int key = (h_bf16 ? 4 : 0) | (dy_bf16 ? 2 : 0) | (dh_bf16 ? 1 : 0);
    switch (key) {
        case 7: T2S_LNB(bf, bf, bf); break;
        case 6: T2S_LNB(bf, bf, float); break;
        case 5: T2S_LNB(bf, float, bf); break;
        case 4: T2S_LNB(bf, float, float); break;
        case 3: T2S_LNB(float, bf, bf); break;
        case 2: T2S_LNB(float, bf, float); break;
        case 1: T2S_LNB(float, float, bf); break;
        default: T2S_LNB(float, float, float); break;
    }
#undef T2S_LNB
    return launch_status("ln_bwd");
}

extern "C" int t2s_ln_bwd(const void* h, int h_bf16, long long ldh, const void* dy, int dy_bf16, long long lddy,
                          int dy_rows_per_group, int dy_group_rows, int dy_row_off, const float* gamma,
                          const float* beta, float eps, int rows, int H, int tanh_out, void* dh, int dh_bf16,
                          long long lddh, float* dgamma, float* dbeta, float* dbias, void* stream) {
    return ln_bwd_entry(h, h_bf16, ldh, dy, dy_bf16, lddy, dy_rows_per_group, dy_group_rows, dy_row_off, gamma, beta, eps,
                        rows, H, tanh_out, dh, dh_bf16, lddh, dgamma, dbeta, dbias, nullptr, DropCfg{0, 0, 0, 0, 1.f}, stream);
}

/* backward of t2s_add_ln_dropout: dh = gradient of the pre-LayerNorm sum (the residual branch), dh_drop = dh * mask
 * (same site / seed as the forward) = gradient of the Linear's output; dbias sums dh_drop */
extern "C" int t2s_ln_bwd_dropout(const void* h, int h_bf16, long long ldh, const void* dy, int dy_bf16, long long lddy,
                                  int dy_rows_per_group, int dy_group_rows, int dy_row_off, const float* gamma,
                                  const float* beta, float eps, int rows, int H, int tanh_out, void* dh, int dh_bf16,
                                  long long lddh, float* dgamma, float* dbeta, float* dbias, void* dh_drop, float p,
                                  unsigned long long seed, unsigned site, void* stream) {
    if (p < 0.f || p >= 1.f || !dh_drop) { set_error("ln_bwd_dropout: bad arguments"); return T2S_ERR_ARG; }
    return ln_bwd_entry(h, h_bf16, ldh, dy, dy_bf16, lddy, dy_rows_per_group, dy_group_rows, dy_row_off, gamma, beta, eps,
                        rows, H, tanh_out, dh, dh_bf16, lddh, dgamma, dbeta, dbias, dh_drop, make_drop(p, seed, site), stream);
}

extern "C" int t2s_colsum(const void* x, int x_bf16, long long ldx, int rows, int N, float* dst, void* stream) {
    if (rows <= 0 || N <= 0) { set_error("colsum: bad shape"); return T2S_ERR_SHAPE; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (x_bf16) {
        if ((ldx % 8) || (reinterpret_cast<uintptr_t>(x) & 15)) { set_error("colsum: bf16 rows need 16-byte alignment"); return T2S_ERR_ALIGN; }
        dim3 grid((N + 255) / 256, (rows + CS_ROWS_PER_CTA - 1) / CS_ROWS_PER_CTA);
        colsum_bf16_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, rows, N, dst);
    } else {
        int gy = (rows + 63) / 64;
        if (gy > 64) gy = 64;
        colsum_f32_kernel<<<dim3((N + 255) / 256, gy), 256, 0, st>>>(reinterpret_cast<const float*>(x), ldx, rows, N, dst);
    }
    return launch_status("colsum");
}

extern "C" int t2s_rows_add(const void* a, const void* b, const void* c, long long ldi, int rows, int H, float* out,
                            long long ldo, int rows_per_group, int out_group_rows, int out_row_off, int accumulate,
                            void* stream) {
    if (rows <= 0 || (H % 4) || (ldi % 4) || (ldo % 4) || !a) { set_error("rows_add: bad shape"); return T2S_ERR_SHAPE; }
    rows_add_kernel<<<flat_grid((long long)rows * (H / 4)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b),
        reinterpret_cast<const __nv_bfloat16*>(c), ldi, rows, H, out, ldo,
        BwRowMap{rows_per_group, out_group_rows, out_row_off}, accumulate);
    return launch_status("rows_add");
}

extern "C" int t2s_gelu_rows(const void* u, int u_bf16, long long ldu, int rows, int N, void* out, long long ldo,
                             int lo_off, void* stream) {
    if (rows <= 0 || (N % 4) || (ldu % 4) || (ldo % 4) || (lo_off % 4)) { set_error("gelu_rows: bad shape"); return T2S_ERR_SHAPE; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = flat_grid((long long)rows * (N / 4));
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    if (u_bf16) gelu_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(u), ldu, rows, N, o, ldo, lo_off);
    else gelu_rows_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(u), ldu, rows, N, o, ldo, lo_off);
    return launch_status("gelu_rows");
}

extern "C" int t2s_embed_scatter_add(const void* src, int src_bf16, long long lds, int c0, int d, const long long* ids,
                                     int rows, long long pad_id, float* table, long long ldt, void* stream) {
    if (rows <= 0 || d <= 0) { set_error("embed_scatter_add: bad shape"); return T2S_ERR_SHAPE; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = flat_grid((long long)rows * d);
    if (src_bf16) embed_scatter_add_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src), lds, c0, d, ids, rows, pad_id, table, ldt);
    else embed_scatter_add_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(src), lds, c0, d, ids, rows, pad_id, table, ldt);
    return launch_status("embed_scatter_add");
}

extern "C" int t2s_ptr_score_bwd(const float* dscores, long long ld_scores, int B, int T, int V, const void* q,
                                 long long ldq, const void* keyp, long long key_batch_stride, long long ldk, int O, int H,
                                 void* dq, long long lddq, void* dkeyp, long long dkey_batch_stride, long long lddk,
                                 void* stream) {
    if (T > PB_T || T <= 0 || (H % 2) || B <= 0 || O <= 0) { set_error("ptr_score_bwd: bad shape (T <= 16)"); return T2S_ERR_SHAPE; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const float inv = 1.0f / sqrtf((float)H);
    const size_t smem_k = (size_t)T * H * sizeof(float);
    static size_t attr = 48 * 1024;
    if (smem_k > attr) {
        cudaError_t e = cudaFuncSetAttribute(ptr_score_bwd_dk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_k);
        if (e != cudaSuccess) { set_error("ptr_score_bwd attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = smem_k;
    }
    ptr_score_bwd_dk_kernel<<<dim3((O + 31) / 32, B), 256, smem_k, st>>>(
        dscores, ld_scores, T, V, reinterpret_cast<const __nv_bfloat16*>(q), ldq, O, H,
        reinterpret_cast<__nv_bfloat16*>(dkeyp), dkey_batch_stride, lddk, inv);
    if ((H % 256) || (ldk % 8)) { set_error("ptr_score_bwd: H %% 256 and ldk %% 8"); return T2S_ERR_SHAPE; }
    const size_t smem_q = ((size_t)((O + 3) & ~3) + 8 * (size_t)H) * sizeof(float);
    static size_t attr_q = 48 * 1024;
    if (smem_q > attr_q) {
        cudaError_t e = cudaFuncSetAttribute(ptr_score_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q);
        if (e != cudaSuccess) { set_error("ptr_score_bwd attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_q = smem_q;
    }
    ptr_score_bwd_dq_kernel<<<dim3(T, B), 256, smem_q, st>>>(
        dscores, ld_scores, T, V, reinterpret_cast<const __nv_bfloat16*>(keyp), key_batch_stride, ldk, O, H,
        reinterpret_cast<__nv_bfloat16*>(dq), lddq, inv);
    return launch_status("ptr_score_bwd");
}

extern "C" int t2s_prev_embed_bwd(const void* dx, long long lddx, const long long* prev_inds, int ld_prev, int B, int T,
                                  int V, int H, const float* ans_w, const float* ocr_emb, long long ocr_batch_stride,
                                  long long ld_ocr, const float* pos_emb, const float* type_emb, const float* ans_g,
                                  const float* ocr_g, const float* emb_g, float eps, float* d_ans_w, float* d_ocr_emb,
                                  float* d_pos, float* d_type, float* d_ans_g, float* d_ans_b, float* d_ocr_g,
                                  float* d_ocr_b, float* d_emb_g, float* d_emb_b, int n_ocr, void* stream) {
    if (!bw_h_ok(H) || B <= 0 || T <= 0 || n_ocr < 0) { set_error("prev_embed_bwd: bad shape"); return T2S_ERR_SHAPE; }
    const size_t smem = (size_t)(BW_THREADS / 32) * H * sizeof(float);
    prev_embed_bwd_kernel<<<bw_row_grid(B * T), BW_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(dx), lddx, prev_inds, ld_prev, B, T, V, H, ans_w, ocr_emb, ocr_batch_stride,
        ld_ocr, pos_emb, type_emb, ans_g, ocr_g, emb_g, eps, d_ans_w, d_ocr_emb, d_pos, d_type, d_ans_g, d_ans_b, d_ocr_g,
        d_ocr_b, d_emb_g, d_emb_b, n_ocr);
    return launch_status("prev_embed_bwd");
}

extern "C" int t2s_ocr_finish_bwd(const float* h, long long ldh, const float* bbox, const float* w2, const float* b2,
                                  const float* g1, const float* g2, float eps, int rows, int H, const float* dout,
                                  long long ldd, int dy_rows_per_group, int dy_group_rows, int dy_row_off, void* dh,
                                  long long lddh, float* dc_ws, long long lddc, float* dg1, float* db1, float* dg2,
                                  float* db2ln, float* dbias1, float* dw2, float* db2, void* stream) {
    if (!bw_h_ok(H) || rows <= 0) { set_error("ocr_finish_bwd: bad shape"); return T2S_ERR_SHAPE; }
    const size_t smem = (size_t)(BW_THREADS / 32) * H * sizeof(float);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    ocr_finish_bwd_kernel<<<bw_row_grid(rows), BW_THREADS, smem, st>>>(
        h, ldh, bbox, w2, b2, g1, g2, eps, rows, H, dout, ldd, BwRowMap{dy_rows_per_group, dy_group_rows, dy_row_off},
        reinterpret_cast<__nv_bfloat16*>(dh), lddh, dc_ws, lddc, dg1, db1, dg2, db2ln, dbias1, db2);
    int gy = (rows + 127) / 128;
    if (gy > 256) gy = 256;
    k4_wgrad_kernel<<<dim3((H + 255) / 256, gy), 256, 0, st>>>(dc_ws, lddc, bbox, rows, H, dw2);
    return launch_status("ocr_finish_bwd");
}

extern "C" int t2s_bert_embed_bwd(const void* dy, long long lddy, const long long* ids, int rows, int L, int H,
                                  const float* word, const float* pos, const float* type0, const float* gamma, float eps,
                                  float* d_word, float* d_pos, float* d_type, float* dgamma, float* dbeta, void* stream) {
    if (!bw_h_ok(H) || rows <= 0) { set_error("bert_embed_bwd: bad shape"); return T2S_ERR_SHAPE; }
    const size_t smem = (size_t)(BW_THREADS / 32) * H * sizeof(float);
    bert_embed_bwd_kernel<<<bw_row_grid(rows), BW_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(dy), lddy, ids, rows, L, H, word, pos, type0, gamma, eps, d_word, d_pos,
        d_type, dgamma, dbeta);
    return launch_status("bert_embed_bwd");
}

extern "C" int t2s_pos_bce_loss_bwd(const float* scores, const float* targets, const float* loss_mask, int B, int T, int N,
                                    const float* grad_out, float* dscores, int accumulate, void* stream) {
    if (B <= 0 || T <= 0 || N <= 0) { set_error("pos_bce_loss_bwd: bad shape"); return T2S_ERR_SHAPE; }
    const long long rows = (long long)B * T;
    bce_bwd_kernel<<<(int)(rows < 2048 ? rows : 2048), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        scores, targets, loss_mask, rows, N, grad_out, dscores, accumulate);
    return launch_status("pos_bce_loss_bwd");
}

extern "C" int t2s_info_nce_loss_bwd(const float* ref, const float* pos, const float* neg, int B, int T, int N,
                                     float temperature, void* workspace, const float* grad_out, float* dref, float* dpos,
                                     float* dneg, int accumulate, void* stream) {
    if (B <= 0 || T <= 0 || N <= 0) { set_error("info_nce_loss_bwd: bad shape"); return T2S_ERR_SHAPE; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // workspace layout of t2s_info_nce_loss: 1024 doubles, then B*T*5 floats of row statistics (recomputed here so the
    // call does not depend on the forward's workspace still being alive), then B*8 floats of coefficients
    float* stats = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 1024 * sizeof(double));
    float* coef = stats + (long long)B * T * 5;
    int rc = t2s_nce_rowstats(ref, pos, neg, B * T, N, stats, stream);
    if (rc) return rc;
    nce_coef_kernel<<<(B + 127) / 128, 128, 0, st>>>(stats, B, T, temperature, grad_out, coef);
    nce_bwd_kernel<<<B * T, 256, 0, st>>>(ref, pos, neg, T, N, stats, coef, dref, dpos, dneg, accumulate);
    return launch_status("info_nce_loss_bwd");
}

extern "C" long long t2s_loss_bwd_workspace_bytes(int B, int T) {
    return (long long)(1024 * sizeof(double)) + (long long)B * T * 5 * sizeof(float) + (long long)B * 8 * sizeof(float);
}

extern "C" int t2s_repack_weights(const float* flat_param, const void* jobs, int n_jobs, int n_tiles, void* stream) {
    if (!flat_param || !jobs || n_jobs <= 0 || n_tiles <= 0) { set_error("repack_weights: bad arguments"); return T2S_ERR_ARG; }
    static_assert(sizeof(RepackJob) == 48, "RepackJob layout is mirrored on the host");
    repack_kernel<<<n_tiles, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(flat_param, reinterpret_cast<const RepackJob*>(jobs), n_jobs);
    return launch_status("repack_weights");
}

extern "C" int t2s_sumsq(const float* g, long long n, void* workspace, float* out, void* stream) {
    if (n <= 0) { set_error("sumsq: empty"); return T2S_ERR_SHAPE; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long gl = (n + 255) / 256;
    const int grid = (int)(gl < 1024 ? gl : 1024);
    sumsq_partial_kernel<<<grid, 256, 0, st>>>(g, n, reinterpret_cast<double*>(workspace));
    sumsq_final_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const double*>(workspace), grid, out);
    return launch_status("sumsq");
}

extern "C" int t2s_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                             float beta2, float eps, int step, const float* sumsq, float max_norm, float grad_scale,
                             void* stream) {
    if (n <= 0 || step < 1) { set_error("adam_step: bad arguments"); return T2S_ERR_SHAPE; }
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    adam_step_kernel<<<flat_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps,
                                                                                      bc1, bc2, sumsq, max_norm, grad_scale);
    return launch_status("adam_step");
}
