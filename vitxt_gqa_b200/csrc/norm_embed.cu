// K3 / K4: embedding gathers, L2-normalise + concat, and (residual +) LayerNorm rows.
// All kernels are HBM-bound streaming kernels: one warp per 768-wide row, 16-byte vector
// accesses, row statistics by warp shuffle, nothing staged in shared memory.
//
// Reference call sites replaced:
//   t2s_bert_embed_ln   BertEmbeddings (word+position+token_type -> LN), via pythia/models/t2s.py:530
//   t2s_feat_concat     F.normalize + nn.Embedding + torch.cat of t2s.py:195-207 and 223-244
//   t2s_add_ln          BertSelfOutput/BertOutput LayerNorm(x + residual); obj_feat_layer_norm
//                       (t2s.py:209-213); QTV's x + tanh(enc(x)) (t2s.py:430-432) as an epilogue
//   t2s_ocr_finish      LN(linear_ocr_feat) + LN(linear_ocr_bbox(bbox)) of t2s.py:246-252
//   t2s_prev_embed      PrevPredEmbeddings.forward (t2s.py:690-723), gathering only the rows used
#include "common.cuh"
#include "../../include/t2s_b200.h"

namespace t2s {

constexpr int NE_THREADS = 128;      // 4 rows per CTA
constexpr int NE_MAXV = 8;           // H <= 1024 (H % 128 == 0)

struct RowMap {           // out_row = (r / per) * group + off + r % per
    int per, group, off;
    __device__ __forceinline__ long long operator()(int r) const {
        return per > 0 ? (long long)(r / per) * group + off + (r % per) : (long long)r;
    }
};

template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    return make_float4(bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y));
}
__device__ __forceinline__ void store4_bf16(__nv_bfloat16* p, float4 v) {
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = o;
}
__device__ __forceinline__ void store4_any(__nv_bfloat16* p, float4 v) { store4_bf16(p, v); }
__device__ __forceinline__ void store4_any(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// bf16 hi|lo split of 4 fp32 values: hi = bf16(v) at p, lo = bf16(v - hi) at p + lo_off
__device__ __forceinline__ void store4_split(__nv_bfloat16* p, long long lo_off, float4 v) {
    uint2 o, l;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    l.x = pack_bf16x2(v.x - bf16lo(o.x), v.y - bf16hi(o.x));
    l.y = pack_bf16x2(v.z - bf16lo(o.y), v.w - bf16hi(o.y));
    *reinterpret_cast<uint2*>(p) = o;
    *reinterpret_cast<uint2*>(p + lo_off) = l;
}

// LayerNorm of one row held as nv float4 per lane (element index = (i*32 + lane)*4).
__device__ __forceinline__ void warp_layernorm(float4 (&x)[NE_MAXV], int nv, int H, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float eps, int lane) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NE_MAXV; ++i)
        if (i < nv) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    const float mean = warp_sum(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NE_MAXV; ++i)
        if (i < nv) {
            const float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    const float var = warp_sum(q) / (float)H;
    const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
    for (int i = 0; i < NE_MAXV; ++i)
        if (i < nv) {
            const int e = (i * 32 + lane) * 4;
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + e));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + e));
            x[i].x = (x[i].x - mean) * rstd * g.x + bb.x;
            x[i].y = (x[i].y - mean) * rstd * g.y + bb.y;
            x[i].z = (x[i].z - mean) * rstd * g.z + bb.z;
            x[i].w = (x[i].w - mean) * rstd * g.w + bb.w;
        }
}

// ------------------------------------------------------------------------------- BertEmbeddings
__global__ void __launch_bounds__(NE_THREADS)
bert_embed_ln_kernel(const long long* __restrict__ ids, int rows, int L, int H, const float* __restrict__ word,
                     const float* __restrict__ pos, const float* __restrict__ type0, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, float* __restrict__ out, long long ldo) {
    const int row = blockIdx.x * (NE_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int nv = H / 128;
    const long long id = ids[row];
    const int p = row % L;
    float4 x[NE_MAXV];
#pragma unroll
    for (int i = 0; i < NE_MAXV; ++i)
        if (i < nv) {
            const int e = (i * 32 + lane) * 4;
            const float4 w = *reinterpret_cast<const float4*>(word + id * H + e);
            const float4 q = *reinterpret_cast<const float4*>(pos + (long long)p * H + e);
            const float4 t = *reinterpret_cast<const float4*>(type0 + e);
            x[i] = make_float4((w.x + q.x) + t.x, (w.y + q.y) + t.y, (w.z + q.z) + t.z, (w.w + q.w) + t.w);
        }
    warp_layernorm(x, nv, H, gamma, beta, eps, lane);
#pragma unroll
    for (int i = 0; i < NE_MAXV; ++i)
        if (i < nv) *reinterpret_cast<float4*>(out + (long long)row * ldo + (i * 32 + lane) * 4) = x[i];
}

// ------------------------------------------------------------------------------- normalise + concat
// out[r] = [ f0/max(|f0|,1e-12) | f1/max(|f1|,1e-12) | tab0[id0[r]] | tab1[id1[r]] | 0-pad ]  (fp32)
__device__ __forceinline__ void put_split(__nv_bfloat16* o, int lo_off, int c, float v) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    o[c] = hi;
    o[lo_off + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// out16 != null: write the row as bf16 hi|lo (lo at column k_pad) instead of fp32 -- the operand format of
// t2s_gemm_bf16x3, saving the fp32 round trip through HBM
__global__ void __launch_bounds__(NE_THREADS)
feat_concat_kernel(const float* __restrict__ f0, int d0, const float* __restrict__ f1, int d1,
                   const long long* __restrict__ id0, const float* __restrict__ tab0,
                   const long long* __restrict__ id1, const float* __restrict__ tab1, int id_dim, int rows,
                   float* __restrict__ out, long long ldo, int k_pad, __nv_bfloat16* __restrict__ out16, long long ldo16) {
    const int row = blockIdx.x * (NE_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* o = out ? out + (long long)row * ldo : nullptr;
    __nv_bfloat16* o16 = out16 ? out16 + (long long)row * ldo16 : nullptr;
    int col = 0;
    for (int seg = 0; seg < 2; ++seg) {
        const float* f = seg == 0 ? f0 : f1;
        const int d = seg == 0 ? d0 : d1;
        if (!f) continue;
        const float* src = f + (long long)row * d;
        float ss = 0.f;
        for (int e = lane; e < d; e += 32) { const float v = src[e]; ss = fmaf(v, v, ss); }
        const float denom = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);     // F.normalize: x / max(||x||, eps)
        for (int e = lane; e < d; e += 32) {
            const float v = src[e] / denom;
            if (o) o[col + e] = v;
            if (o16) put_split(o16, k_pad, col + e, v);
        }
        col += d;
    }
    for (int seg = 0; seg < 2; ++seg) {
        const long long* idp = seg == 0 ? id0 : id1;
        const float* tab = seg == 0 ? tab0 : tab1;
        if (!idp) continue;
        const float* src = tab + idp[row] * id_dim;
        for (int e = lane; e < id_dim; e += 32) {
            if (o) o[col + e] = src[e];
            if (o16) put_split(o16, k_pad, col + e, src[e]);
        }
        col += id_dim;
    }
    for (int e = col + lane; e < k_pad; e += 32) {
        if (o) o[e] = 0.f;
        if (o16) { o16[e] = __float2bfloat16_rn(0.f); o16[k_pad + e] = __float2bfloat16_rn(0.f); }
    }
}

// ------------------------------------------------------------------------------- (residual +) LayerNorm
// DROP (training step): y = LN(dropout(x) + res) as BertSelfOutput / BertOutput compute it (Linear -> dropout ->
// LN(. + input)); the pre-LayerNorm sum is written back to `h_out` (may alias x) because the backward's LayerNorm
// needs it.  x and h_out are not __restrict__ in that form.
// RPW rows per warp.  A bf16 row is 1.5 KB and ncu shows the 66 816-row bf16 launches at 3.4 TB/s against 5.3 - 6.1 TB/s
// for the fp32 rows of the same kernel, so loading two rows before the first reduction was tried (RPW = 2): add_ln went
// from 1.87 to 3.34 ms per eval step (96 registers, five CTAs per SM).  RPW stays 1; the template is kept for the record.
// The opposite direction -- __launch_bounds__(128, 10): 48 registers instead of 60, ten CTAs per SM instead of eight, ~50 B of
// spills -- measured 2.12 against 1.95 ms per step: not kept either.
template <typename TX, typename TR, bool SPLIT, bool DROP, int RPW>
__global__ void __launch_bounds__(NE_THREADS)
add_ln_kernel(const TX* x, long long ldx, const TR* __restrict__ res, long long ldr,
              const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int rows, int H,
              const float* __restrict__ tanh_base, long long ld_base, float* __restrict__ out32, long long ldo32,
              __nv_bfloat16* __restrict__ out16, long long ldo16, RowMap map, TX* h_out, DropCfg drop) {
    const int row0 = (blockIdx.x * (NE_THREADS / 32) + (threadIdx.x >> 5)) * RPW, lane = threadIdx.x & 31;
    pdl_wait();            // decode-chain launches come in early (common.cuh)
    pdl_release();
    if (row0 >= rows) return;
    const int nv = H / 128;
    float4 v[RPW][NE_MAXV];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int row = row0 + r;
        if (row >= rows) break;
#pragma unroll
        for (int i = 0; i < NE_MAXV; ++i)
            if (i < nv) {
                const int e = (i * 32 + lane) * 4;
                v[r][i] = load4<TX>(x + (long long)row * ldx + e);
                if (DROP && drop.thr) {
                    const float4 m = drop_mask4(drop, row, H, e);
                    v[r][i].x *= m.x; v[r][i].y *= m.y; v[r][i].z *= m.z; v[r][i].w *= m.w;
                }
                if (res) {
                    const float4 q = load4<TR>(res + (long long)row * ldr + e);
                    v[r][i].x += q.x; v[r][i].y += q.y; v[r][i].z += q.z; v[r][i].w += q.w;
                }
                if (DROP && h_out) store4_any(h_out + (long long)row * ldx + e, v[r][i]);
            }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int row = row0 + r;
        if (row >= rows) break;
        warp_layernorm(v[r], nv, H, gamma, beta, eps, lane);
        const long long orow = map(row);
#pragma unroll
        for (int i = 0; i < NE_MAXV; ++i)
            if (i < nv) {
                const int e = (i * 32 + lane) * 4;
                if (tanh_base) {
                    const float4 b = *reinterpret_cast<const float4*>(tanh_base + orow * ld_base + e);
                    v[r][i].x = b.x + tanhf(v[r][i].x); v[r][i].y = b.y + tanhf(v[r][i].y);
                    v[r][i].z = b.z + tanhf(v[r][i].z); v[r][i].w = b.w + tanhf(v[r][i].w);
                }
                if (out32) *reinterpret_cast<float4*>(out32 + orow * ldo32 + e) = v[r][i];
                if (out16) {
                    if (SPLIT) store4_split(out16 + orow * ldo16 + e, H, v[r][i]);
                    else store4_bf16(out16 + orow * ldo16 + e, v[r][i]);
                }
            }
    }
}

// ------------------------------------------------------------------------------- OCR: LN(h) + LN(W2.bbox + b2)
__global__ void __launch_bounds__(NE_THREADS)
ocr_finish_kernel(const float* __restrict__ h, long long ldh, const float* __restrict__ bbox,
                  const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ g1,
                  const float* __restrict__ be1, const float* __restrict__ g2, const float* __restrict__ be2,
                  float eps, int rows, int H, float* __restrict__ out, long long ldo, RowMap map) {
    const int row = blockIdx.x * (NE_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int nv = H / 128;
    const float4 bx = *reinterpret_cast<const float4*>(bbox + (long long)row * 4);
    float4 a[NE_MAXV], c[NE_MAXV];
#pragma unroll
    for (int i = 0; i < NE_MAXV; ++i)
        if (i < nv) {
            const int e = (i * 32 + lane) * 4;
            a[i] = *reinterpret_cast<const float4*>(h + (long long)row * ldh + e);
            float r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(w2 + (long long)(e + j) * 4));
                // same association order as a K=4 dot product accumulated left to right, then + bias
                r[j] = fmaf(bx.w, w.w, fmaf(bx.z, w.z, fmaf(bx.y, w.y, bx.x * w.x))) + __ldg(b2 + e + j);
            }
            c[i] = make_float4(r[0], r[1], r[2], r[3]);
        }
    warp_layernorm(a, nv, H, g1, be1, eps, lane);
    warp_layernorm(c, nv, H, g2, be2, eps, lane);
    const long long orow = map(row);
#pragma unroll
    for (int i = 0; i < NE_MAXV; ++i)
        if (i < nv) {
            const int e = (i * 32 + lane) * 4;
            *reinterpret_cast<float4*>(out + orow * ldo + e) =
                make_float4(a[i].x + c[i].x, a[i].y + c[i].y, a[i].z + c[i].z, a[i].w + c[i].w);
        }
}

// ------------------------------------------------------------------------------- PrevPredEmbeddings
// dec[b, t] = LN_ans(Wcls[idx]) or LN_ocr(ocr_emb[b, idx - V])  +  LN_emb(pos[t] + type[idx >= V])
__global__ void __launch_bounds__(NE_THREADS)
prev_embed_kernel(const long long* __restrict__ prev_inds, int ld_prev, int B, int t0, int nt, int V, int H,
                  const float* __restrict__ ans_w, const float* __restrict__ ocr_emb, long long ocr_batch_stride,
                  long long ld_ocr, const float* __restrict__ pos_emb, const float* __restrict__ type_emb,
                  const float* __restrict__ ans_g, const float* __restrict__ ans_b, const float* __restrict__ ocr_g,
                  const float* __restrict__ ocr_b, const float* __restrict__ emb_g, const float* __restrict__ emb_b,
                  float eps, __nv_bfloat16* __restrict__ out16, float* __restrict__ out32, long long ldo, int T,
                  int n_ocr) {
    const int w = blockIdx.x * (NE_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    pdl_wait();
    pdl_release();
    if (w >= B * nt) return;
    const int b = w / nt, t = t0 + w % nt;
    const int nv = H / 128;
    // indices outside [0, V + n_ocr) cannot come from an argmax over the score row; clamp instead of reading out of bounds
    const long long idx = clamp_index(prev_inds[(long long)b * ld_prev + t], (long long)V + n_ocr);
    const bool is_ocr = idx >= V;
    const float* src = is_ocr ? ocr_emb + (long long)b * ocr_batch_stride + (idx - V) * ld_ocr : ans_w + idx * H;
    const float* ty = type_emb + (is_ocr ? H : 0);
    float4 r[NE_MAXV], e[NE_MAXV];
#pragma unroll
    for (int i = 0; i < NE_MAXV; ++i)
        if (i < nv) {
            const int c = (i * 32 + lane) * 4;
            r[i] = *reinterpret_cast<const float4*>(src + c);
            const float4 p = *reinterpret_cast<const float4*>(pos_emb + (long long)t * H + c);
            const float4 q = *reinterpret_cast<const float4*>(ty + c);
            e[i] = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
        }
    warp_layernorm(r, nv, H, is_ocr ? ocr_g : ans_g, is_ocr ? ocr_b : ans_b, eps, lane);
    warp_layernorm(e, nv, H, emb_g, emb_b, eps, lane);
    const long long orow = (long long)b * T + t;
#pragma unroll
    for (int i = 0; i < NE_MAXV; ++i)
        if (i < nv) {
            const int c = (i * 32 + lane) * 4;
            const float4 v = make_float4(r[i].x + e[i].x, r[i].y + e[i].y, r[i].z + e[i].z, r[i].w + e[i].w);
            if (out16) store4_bf16(out16 + orow * ldo + c, v);
            if (out32) *reinterpret_cast<float4*>(out32 + orow * ldo + c) = v;
        }
}

// ------------------------------------------------------------------------------- fp32 -> bf16 rows
__global__ void cast_rows_bf16_kernel(const float* __restrict__ x, long long ldx, int rows, int H,
                                      __nv_bfloat16* __restrict__ out, long long ldo, RowMap map) {
    const long long n4 = (long long)rows * (H / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / (H / 4)), c = (int)(i % (H / 4)) * 4;
        const float4 v = *reinterpret_cast<const float4*>(x + (long long)row * ldx + c);
        store4_bf16(out + map(row) * ldo + c, v);
    }
}

// fp32 rows -> bf16 hi|lo rows (operand format of t2s_gemm_bf16x3); zero-fills the K..lo_off padding
__global__ void split_bf16_kernel(const float* __restrict__ x, long long ldx, int rows, int K, int lo_off,
                                  __nv_bfloat16* __restrict__ out, long long ldo, RowMap in_map) {
    const int c4n = lo_off / 4;
    const long long n4 = (long long)rows * c4n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / c4n), c = (int)(i % c4n) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const long long irow = in_map(row);
        if (c + 4 <= K) v = *reinterpret_cast<const float4*>(x + irow * ldx + c);
        else if (c < K) {
            const float* p = x + irow * ldx + c;
            v.x = p[0];
            if (c + 1 < K) v.y = p[1];
            if (c + 2 < K) v.z = p[2];
        }
        store4_split(out + (long long)row * ldo + c, lo_off, v);
    }
}

// in place: x[map(r), :] *= mask(r, :) * scale  (nn.Dropout on embedding rows: BertEmbeddings, obj_drop / ocr_drop of
// t2s.py:95,118,214,253, PrevPredEmbeddings.emb_dropout t2s.py:720 -- and, in the backward, the same mask on the
// gradient rows).  `r` is the compact row index: forward and backward address a site the same way.
template <typename T>
__global__ void __launch_bounds__(256)
dropout_rows_kernel(T* __restrict__ x, long long ldx, int rows, int H, RowMap map, DropCfg drop) {
    const long long n4 = (long long)rows * (H / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / (H / 4)), c = (int)(i % (H / 4)) * 4;
        T* p = x + map(row) * ldx + c;
        float4 v = load4<T>(p);
        const float4 m = drop_mask4(drop, row, H, c);
        v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
        store4_any(p, v);
    }
}
// test / inspection helper: the multipliers (0 or scale) of a site, as fp32
__global__ void __launch_bounds__(256)
dropout_mask_kernel(float* __restrict__ out, long long n_pairs, int mode, int H, int n_query, int n_key, DropCfg drop) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pairs; i += (long long)gridDim.x * blockDim.x) {
        float m0, m1;
        if (mode == 0) {                       // [rows, H]: pair i = row * H/2 + col/2
            drop_pair(drop, (uint32_t)i, drop.site << 20, m0, m1);
            out[2 * i] = m0;
            out[2 * i + 1] = m1;
        } else {                               // [BH, n_query, n_key] attention probabilities
            const int kp = (n_key + 1) / 2;
            const int bh = (int)(i / ((long long)n_query * kp));
            const int q = (int)((i / kp) % n_query), jp = (int)(i % kp);
            drop_pair(drop, drop_attn_x(q, 2 * jp), drop_attn_y(drop, bh), m0, m1);
            float* o = out + ((long long)bh * n_query + q) * n_key;
            o[2 * jp] = m0;
            if (2 * jp + 1 < n_key) o[2 * jp + 1] = m1;
        }
    }
}

static inline int rows_grid(int rows) { return (rows + NE_THREADS / 32 - 1) / (NE_THREADS / 32); }
static inline bool h_ok(int H) { return H % 128 == 0 && H / 128 <= NE_MAXV && H > 0; }

}  // namespace t2s

using namespace t2s;

extern "C" int t2s_bert_embed_ln(const long long* ids, int rows, int L, int H, const float* word, const float* pos,
                                 const float* type0, const float* gamma, const float* beta, float eps, float* out,
                                 long long ldo, void* stream) {
    if (!h_ok(H) || rows <= 0) { set_error("bert_embed_ln: H %d must be a multiple of 128 <= 1024", H); return T2S_ERR_SHAPE; }
    bert_embed_ln_kernel<<<rows_grid(rows), NE_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        ids, rows, L, H, word, pos, type0, gamma, beta, eps, out, ldo);
    return launch_status("bert_embed_ln");
}

extern "C" int t2s_feat_concat(const float* f0, int d0, const float* f1, int d1, const long long* id0, const float* tab0,
                               const long long* id1, const float* tab1, int id_dim, int rows, float* out, long long ldo,
                               int k_pad, void* out_split, long long ldo_split, void* stream) {
    const int k = (f0 ? d0 : 0) + (f1 ? d1 : 0) + (id0 ? id_dim : 0) + (id1 ? id_dim : 0);
    if (rows <= 0 || k > k_pad || (out && k_pad > ldo) || (out_split && 2LL * k_pad > ldo_split) || (!out && !out_split)) {
        set_error("feat_concat: bad widths k %d k_pad %d ldo %lld ldo_split %lld", k, k_pad, ldo, ldo_split);
        return T2S_ERR_SHAPE;
    }
    feat_concat_kernel<<<rows_grid(rows), NE_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        f0, d0, f1, d1, id0, tab0, id1, tab1, id_dim, rows, out, ldo, k_pad, reinterpret_cast<__nv_bfloat16*>(out_split),
        ldo_split);
    return launch_status("feat_concat");
}

static int add_ln_entry(bool split, const void* x, int x_bf16, long long ldx, const void* res, int res_bf16,
                        long long ldr, const float* gamma, const float* beta, float eps, int rows, int H,
                        const float* tanh_base, long long ld_base, float* out32, long long ldo32, void* out16,
                        long long ldo16, int rows_per_group, int out_group_rows, int out_row_off, void* stream,
                        bool train = false, void* h_out = nullptr, DropCfg drop = DropCfg{0, 0, 0, 0, 1.f}) {
    if (!h_ok(H) || rows <= 0) { set_error("add_ln: H %d must be a multiple of 128 <= 1024", H); return T2S_ERR_SHAPE; }
    if (!out32 && !out16) { set_error("add_ln: no output"); return T2S_ERR_ARG; }
    if (train && (H % 2)) { set_error("add_ln_dropout: H must be even"); return T2S_ERR_SHAPE; }
    if (split && (!out16 || ldo16 < 2LL * H)) { set_error("add_ln_split: needs out16 with row pitch >= 2H"); return T2S_ERR_ARG; }
    RowMap map{rows_per_group, out_group_rows, out_row_off};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(out16);
#define T2S_LN_LAUNCH(TX, TR, SP)                                                                                 \
    do {                                                                                                          \
        constexpr int RPW = 1;     /* two bf16 rows per warp: measured 1.8x SLOWER on the 66 816-row launches */  \
        const int grid = rows_grid((rows + RPW - 1) / RPW);                                                       \
        if (train)                                                                                                \
            add_ln_kernel<TX, TR, SP, true, RPW><<<grid, NE_THREADS, 0, st>>>(                                    \
                reinterpret_cast<const TX*>(x), ldx, reinterpret_cast<const TR*>(res), ldr, gamma, beta, eps,     \
                rows, H, tanh_base, ld_base, out32, ldo32, o16, ldo16, map, reinterpret_cast<TX*>(h_out), drop);  \
        else                                                                                                      \
            launch_pdl(rows <= PDL_MAX_ROWS, add_ln_kernel<TX, TR, SP, false, RPW>, dim3(grid), dim3(NE_THREADS), \
                       0, st, reinterpret_cast<const TX*>(x), ldx, reinterpret_cast<const TR*>(res), ldr, gamma,  \
                       beta, eps, rows, H, tanh_base, ld_base, out32, ldo32, o16, ldo16, map,                     \
                       static_cast<TX*>(nullptr), drop);                                                          \
    } while (0)
    if (split) {
        if (x_bf16) { set_error("add_ln_split: fp32 input only"); return T2S_ERR_ARG; }
        if (res_bf16 && res) T2S_LN_LAUNCH(float, __nv_bfloat16, true);
        else T2S_LN_LAUNCH(float, float, true);
    } else if (x_bf16) {
        if (res_bf16 || !res) T2S_LN_LAUNCH(__nv_bfloat16, __nv_bfloat16, false);
        else T2S_LN_LAUNCH(__nv_bfloat16, float, false);
    } else {
        if (res_bf16 && res) T2S_LN_LAUNCH(float, __nv_bfloat16, false);
        else T2S_LN_LAUNCH(float, float, false);
    }
#undef T2S_LN_LAUNCH
    return launch_status("add_ln");
}

extern "C" int t2s_add_ln(const void* x, int x_bf16, long long ldx, const void* res, int res_bf16, long long ldr,
                          const float* gamma, const float* beta, float eps, int rows, int H, const float* tanh_base,
                          long long ld_base, float* out32, long long ldo32, void* out16, long long ldo16,
                          int rows_per_group, int out_group_rows, int out_row_off, void* stream) {
    return add_ln_entry(false, x, x_bf16, ldx, res, res_bf16, ldr, gamma, beta, eps, rows, H, tanh_base, ld_base, out32,
                        ldo32, out16, ldo16, rows_per_group, out_group_rows, out_row_off, stream);
}

extern "C" int t2s_add_ln_split(const void* x, int x_bf16, long long ldx, const void* res, int res_bf16, long long ldr,
                                const float* gamma, const float* beta, float eps, int rows, int H,
                                const float* tanh_base, long long ld_base, float* out32, long long ldo32, void* out16,
                                long long ldo16, int rows_per_group, int out_group_rows, int out_row_off, void* stream) {
    return add_ln_entry(true, x, x_bf16, ldx, res, res_bf16, ldr, gamma, beta, eps, rows, H, tanh_base, ld_base, out32,
                        ldo32, out16, ldo16, rows_per_group, out_group_rows, out_row_off, stream);
}

extern "C" int t2s_add_ln_dropout(const void* x, int x_bf16, long long ldx, const void* res, int res_bf16, long long ldr,
                                  const float* gamma, const float* beta, float eps, int rows, int H,
                                  const float* tanh_base, long long ld_base, float* out32, long long ldo32, void* out16,
                                  long long ldo16, int split, int rows_per_group, int out_group_rows, int out_row_off,
                                  void* h_out, float p, unsigned long long seed, unsigned site, void* stream) {
    if (p < 0.f || p >= 1.f) { set_error("add_ln_dropout: p %f outside [0, 1)", p); return T2S_ERR_ARG; }
    return add_ln_entry(split != 0, x, x_bf16, ldx, res, res_bf16, ldr, gamma, beta, eps, rows, H, tanh_base, ld_base,
                        out32, ldo32, out16, ldo16, rows_per_group, out_group_rows, out_row_off, stream, true, h_out,
                        make_drop(p, seed, site));
}

extern "C" int t2s_dropout_rows(void* x, int x_bf16, long long ldx, int rows, int H, int rows_per_group, int group_rows,
                                int row_off, float p, unsigned long long seed, unsigned site, void* stream) {
    if (rows <= 0 || H <= 0 || (H % 4) || (ldx % 4) || p < 0.f || p >= 1.f) { set_error("dropout_rows: bad arguments"); return T2S_ERR_ARG; }
    if (p == 0.f) return T2S_OK;
    const long long n4 = (long long)rows * (H / 4);
    int grid = (int)((n4 + 255) / 256);
    if (grid > num_sms() * 16) grid = num_sms() * 16;
    RowMap map{rows_per_group, group_rows, row_off};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (x_bf16) dropout_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(x), ldx, rows, H, map, make_drop(p, seed, site));
    else dropout_rows_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<float*>(x), ldx, rows, H, map, make_drop(p, seed, site));
    return launch_status("dropout_rows");
}

extern "C" int t2s_dropout_mask(float* out, int mode, int rows_or_bh, int H, int n_query, int n_key, float p,
                                unsigned long long seed, unsigned site, void* stream) {
    if (!out || rows_or_bh <= 0 || p < 0.f || p >= 1.f || (mode == 0 && (H <= 0 || (H % 2))) ||
        (mode != 0 && (n_query <= 0 || n_key <= 0 || n_query > 65535))) { set_error("dropout_mask: bad arguments"); return T2S_ERR_ARG; }
    const long long n_pairs = mode == 0 ? (long long)rows_or_bh * (H / 2) : (long long)rows_or_bh * n_query * ((n_key + 1) / 2);
    int grid = (int)((n_pairs + 255) / 256);
    if (grid > num_sms() * 16) grid = num_sms() * 16;
    dropout_mask_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, n_pairs, mode, H, n_query, n_key,
                                                                                  make_drop(p, seed, site));
    return launch_status("dropout_mask");
}

extern "C" int t2s_split_bf16(const float* x, long long ldx, int rows, int K, int lo_off, void* out, long long ldo,
                              int rows_per_group, int in_group_rows, int in_row_off, void* stream) {
    if (rows <= 0 || K <= 0 || lo_off < K || (lo_off % 8) || ldo < 2LL * lo_off || (ldx % 4) || (ldo % 8)) {
        set_error("split_bf16: bad shape (K %d lo_off %d ldx %lld ldo %lld)", K, lo_off, ldx, ldo);
        return T2S_ERR_SHAPE;
    }
    const long long n4 = (long long)rows * (lo_off / 4);
    int grid = (int)((n4 + 255) / 256);
    if (grid > num_sms() * 16) grid = num_sms() * 16;
    split_bf16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, ldx, rows, K, lo_off, reinterpret_cast<__nv_bfloat16*>(out), ldo,
        RowMap{rows_per_group, in_group_rows, in_row_off});
    return launch_status("split_bf16");
}

extern "C" int t2s_ocr_finish(const float* h, long long ldh, const float* bbox, const float* w2, const float* b2,
                              const float* g1, const float* be1, const float* g2, const float* be2, float eps, int rows,
                              int H, float* out, long long ldo, int rows_per_group, int out_group_rows, int out_row_off,
                              void* stream) {
    if (!h_ok(H) || rows <= 0) { set_error("ocr_finish: H %d must be a multiple of 128 <= 1024", H); return T2S_ERR_SHAPE; }
    RowMap map{rows_per_group, out_group_rows, out_row_off};
    ocr_finish_kernel<<<rows_grid(rows), NE_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        h, ldh, bbox, w2, b2, g1, be1, g2, be2, eps, rows, H, out, ldo, map);
    return launch_status("ocr_finish");
}

extern "C" int t2s_prev_embed(const long long* prev_inds, int ld_prev, int B, int t0, int nt, int T, int V, int H,
                              const float* ans_w, const float* ocr_emb, long long ocr_batch_stride, long long ld_ocr,
                              const float* pos_emb, const float* type_emb, const float* ans_g, const float* ans_b,
                              const float* ocr_g, const float* ocr_b, const float* emb_g, const float* emb_b, float eps,
                              void* out16, float* out32, long long ldo, int n_ocr, void* stream) {
    if (!h_ok(H) || B <= 0 || nt <= 0 || t0 < 0 || t0 + nt > T || n_ocr < 0 || V <= 0) { set_error("prev_embed: bad arguments"); return T2S_ERR_SHAPE; }
    launch_pdl(B * nt <= PDL_MAX_ROWS, prev_embed_kernel, dim3(rows_grid(B * nt)), dim3(NE_THREADS), 0,
               reinterpret_cast<cudaStream_t>(stream), prev_inds, ld_prev, B, t0, nt, V, H, ans_w, ocr_emb,
               ocr_batch_stride, ld_ocr, pos_emb, type_emb, ans_g, ans_b, ocr_g, ocr_b, emb_g, emb_b, eps,
               reinterpret_cast<__nv_bfloat16*>(out16), out32, ldo, T, n_ocr);
    return launch_status("prev_embed");
}

extern "C" int t2s_cast_rows_bf16(const float* x, long long ldx, int rows, int H, void* out, long long ldo,
                                  int rows_per_group, int out_group_rows, int out_row_off, void* stream) {
    if (H % 4 || rows <= 0) { set_error("cast_rows_bf16: H %% 4"); return T2S_ERR_SHAPE; }
    RowMap map{rows_per_group, out_group_rows, out_row_off};
    const long long n4 = (long long)rows * (H / 4);
    int grid = (int)((n4 + 255) / 256);
    if (grid > num_sms() * 16) grid = num_sms() * 16;
    cast_rows_bf16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, ldx, rows, H, reinterpret_cast<__nv_bfloat16*>(out), ldo, map);
    return launch_status("cast_rows_bf16");
}
