// K2: masked multi-head attention without materialising masks or probabilities.
//
// Replaces BertSelfAttention's matmul / +mask / softmax / matmul chain
// (pytorch_transformers modeling_bert, invoked from reference pythia/models/t2s.py:423,538,622)
// and the [B,1,L,L] fp32 masks the reference builds at t2s.py:413-419 and 609-618.
//
// Mask semantics: the reference adds -10000.0 to masked keys; exp(-10000 - max) is exactly 0
// in fp32 and every query has at least one valid key (question tokens), so *skipping* the
// masked keys is arithmetically identical.  All kernels therefore take a compacted per-sample
// key list (`key_idx[b, 0..n_keys[b])` = row indices of the valid keys inside the sample).
//
//   t2s_attn_f32   fp32 flash-style SIMT kernel (grounding chain: TextBert, QTV)
//   t2s_attn_x3    fp32-class flash-style kernel on mma.sync with bf16 hi|lo operands (three products per
//                  contraction): TextBert / QTV of the grounding chain
//   t2s_attn_bf16  bf16 flash-style kernel on mma.sync m16n8k16, K/V staged in swizzled
//                  shared memory by cp.async (encoder rows of the answer transformer)
//   t2s_attn_dec   small-sequence kernel for the <=16 decoder rows: K/V of the valid encoder
//                  keys + causal decoder keys, scores staged in shared memory, warp-shuffle
//                  softmax (t2s.py:574-579,609-615 prefix-LM mask)
#include "common.cuh"
#include <stdlib.h>
#include "../../include/t2s_b200.h"

namespace t2s {

constexpr int DH = 64;  // head size (hidden 768 / 12 heads), fixed by the reference configs

// =============================================================================== fp32 SIMT
constexpr int AF_BQ = 64, AF_BK = 64, AF_PITCH = 68, AF_THREADS = 256;
constexpr int AF_SMEM = 4 * AF_BQ * AF_PITCH * 4;

__global__ void __launch_bounds__(AF_THREADS, 2)
attn_f32_kernel(const float* __restrict__ qkv, long long ld, int L, int H, const int* __restrict__ key_idx,
                const int* __restrict__ n_keys, int key_stride, float* __restrict__ out, long long ldo,
                __nv_bfloat16* __restrict__ out_split, long long ldos, float scale) {
    extern __shared__ __align__(16) float sm[];
    float(*Qs)[AF_PITCH] = reinterpret_cast<float(*)[AF_PITCH]>(sm);
    float(*Ks)[AF_PITCH] = reinterpret_cast<float(*)[AF_PITCH]>(sm + AF_BQ * AF_PITCH);
    float(*Vs)[AF_PITCH] = reinterpret_cast<float(*)[AF_PITCH]>(sm + 2 * AF_BQ * AF_PITCH);
    float(*Ps)[AF_PITCH] = reinterpret_cast<float(*)[AF_PITCH]>(sm + 3 * AF_BQ * AF_PITCH);

    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AF_BQ;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int nk = n_keys[b];
    const int* kidx = key_idx + (long long)b * key_stride;
    const float* base = qkv + (long long)b * L * ld;

    // Q tile (pre-scaled): 64 rows x 16 float4
    for (int i = tid; i < AF_BQ * 16; i += AF_THREADS) {
        const int r = i >> 4, c4 = (i & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < L) v = *reinterpret_cast<const float4*>(base + (long long)(q0 + r) * ld + h * DH + c4);
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        *reinterpret_cast<float4*>(&Qs[r][c4]) = v;
    }

    float m_i[4], l_i[4], o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m_i[i] = -INFINITY;
        l_i[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    }

    for (int k0 = 0; k0 < nk; k0 += AF_BK) {
        __syncthreads();   // previous tile fully consumed (also covers the Q store on the first pass)
        for (int i = tid; i < AF_BK * 16; i += AF_THREADS) {
            const int r = i >> 4, c4 = (i & 15) * 4;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (k0 + r < nk) {
                const float* rowp = base + (long long)kidx[k0 + r] * ld + h * DH + c4;
                kv = *reinterpret_cast<const float4*>(rowp + H);
                vv = *reinterpret_cast<const float4*>(rowp + 2 * H);
            }
            *reinterpret_cast<float4*>(&Ks[r][c4]) = kv;
            *reinterpret_cast<float4*>(&Vs[r][c4]) = vv;
        }
        __syncthreads();

        // S[i][j] for rows ty*4+i, key columns tx+16j
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
        for (int d = 0; d < DH; d += 4) {
            float4 qv[4], kv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(&Qs[ty * 4 + i][d]);
#pragma unroll
            for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4*>(&Ks[tx + 16 * j][d]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[i][j] = fmaf(qv[i].x, kv[j].x, s[i][j]);
                    s[i][j] = fmaf(qv[i].y, kv[j].y, s[i][j]);
                    s[i][j] = fmaf(qv[i].z, kv[j].z, s[i][j]);
                    s[i][j] = fmaf(qv[i].w, kv[j].w, s[i][j]);
                }
        }
        // online softmax; a row's 64 columns live on the 16 lanes sharing `ty`
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (k0 + tx + 16 * j >= nk) s[i][j] = -INFINITY;
                mx = fmaxf(mx, s[i][j]);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float m_new = fmaxf(m_i[i], mx);
            const float corr = expf(m_i[i] - m_new);     // exp(-inf) = 0 on the first tile
            float rs = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = expf(s[i][j] - m_new);
                Ps[ty * 4 + i][tx + 16 * j] = p;
                rs += p;
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
            l_i[i] = l_i[i] * corr + rs;
            m_i[i] = m_new;
#pragma unroll
            for (int j = 0; j < 4; ++j) o[i][j] *= corr;
        }
        __syncthreads();
        // O[rows ty*4+i][cols tx*4..+3] += P . V
#pragma unroll 4
        for (int c = 0; c < AF_BK; c += 4) {
            float4 pv[4], vv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pv[i] = *reinterpret_cast<const float4*>(&Ps[ty * 4 + i][c]);
#pragma unroll
            for (int j = 0; j < 4; ++j) vv[j] = *reinterpret_cast<const float4*>(&Vs[c + j][tx * 4]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float p[4] = {pv[i].x, pv[i].y, pv[i].z, pv[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    o[i][0] = fmaf(p[j], vv[j].x, o[i][0]);
                    o[i][1] = fmaf(p[j], vv[j].y, o[i][1]);
                    o[i][2] = fmaf(p[j], vv[j].z, o[i][2]);
                    o[i][3] = fmaf(p[j], vv[j].w, o[i][3]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = q0 + ty * 4 + i;
        if (r < L) {
            const float inv = 1.0f / l_i[i];
            float4 v = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
            if (out) *reinterpret_cast<float4*>(out + ((long long)b * L + r) * ldo + h * DH + tx * 4) = v;
            if (out_split) {     // bf16 hi|lo (lo at column H): operand format of t2s_gemm_bf16x3
                uint2 hi, lo;
                hi.x = pack_bf16x2(v.x, v.y); hi.y = pack_bf16x2(v.z, v.w);
                lo.x = pack_bf16x2(v.x - bf16lo(hi.x), v.y - bf16hi(hi.x));
                lo.y = pack_bf16x2(v.z - bf16lo(hi.y), v.w - bf16hi(hi.y));
                __nv_bfloat16* op = out_split + ((long long)b * L + r) * ldos + h * DH + tx * 4;
                *reinterpret_cast<uint2*>(op) = hi;
                *reinterpret_cast<uint2*>(op + H) = lo;
            }
        }
    }
}

// =============================================================================== bf16 mma.sync
constexpr int AB_BQ = 64, AB_BK = 64, AB_THREADS = 128;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// tile element (row, 16-byte chunk) -> byte offset with the chunk index XOR-swizzled by row
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__global__ void __launch_bounds__(AB_THREADS)
attn_bf16_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld, int L, int H, const int* __restrict__ key_idx,
                 const int* __restrict__ n_keys, int key_stride, __nv_bfloat16* __restrict__ out, long long ldo,
                 float scale_log2) {
    __shared__ __align__(128) uint8_t Qs[AB_BQ * 128];
    __shared__ __align__(128) uint8_t Ks[2][AB_BK * 128];
    __shared__ __align__(128) uint8_t Vs[2][AB_BK * 128];
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AB_BQ;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nk = n_keys[b];
    const int* kidx = key_idx + (long long)b * key_stride;
    const __nv_bfloat16* base = qkv + (long long)b * L * ld;
    const int ntiles = (nk + AB_BK - 1) / AB_BK;

    // Q tile: 64 rows x 8 chunks
    for (int i = tid; i < AB_BQ * 8; i += AB_THREADS) {
        const int r = i >> 3, c = i & 7;
        const bool ok = q0 + r < L;
        const __nv_bfloat16* src = base + (long long)(ok ? q0 + r : 0) * ld + h * DH + c * 8;
        cp_async16(Qs + sw_off(r, c), src, ok);
    }
    auto load_kv = [&](int t, int buf) {
        for (int i = tid; i < AB_BK * 8; i += AB_THREADS) {
            const int r = i >> 3, c = i & 7;
            const bool ok = t * AB_BK + r < nk;
            const long long row = ok ? kidx[t * AB_BK + r] : 0;
            const __nv_bfloat16* src = base + row * ld + h * DH + c * 8;
            cp_async16(Ks[buf] + sw_off(r, c), src + H, ok);
            cp_async16(Vs[buf] + sw_off(r, c), src + 2 * H, ok);
        }
    };
    if (ntiles > 0) load_kv(0, 0);
    cp_async_commit();

    float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[n][j] = 0.f;
    uint32_t qf[4][4];

    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) load_kv(t + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (t == 0) {
            // A fragments of Q for this warp's 16 rows: 4 k-steps of 16
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int row = warp * 16 + (lane & 15);
                const int chunk = ks * 2 + (lane >> 4);
                ldmatrix_x4(qf[ks], smem_u32(Qs + sw_off(row, chunk)));
            }
        }
        // S = Q K^T : 8 n-tiles (8 keys each) x 4 k-steps
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[n][j] = 0.f;
#pragma unroll
        for (int np = 0; np < 4; ++np) {         // pairs of n-tiles (16 keys)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t kb[4];
                // matrices: (keys np*16+0..7, d ks*16+0..7), (same keys, d +8), (keys +8, d), (keys +8, d +8)
                const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int chunk = ks * 2 + ((lane >> 3) & 1);
                ldmatrix_x4(kb, smem_u32(Ks[buf] + sw_off(row, chunk)));
                mma_bf16_16816(s[np * 2], qf[ks], kb[0], kb[1]);
                mma_bf16_16816(s[np * 2 + 1], qf[ks], kb[2], kb[3]);
            }
        }
        // online softmax in the exp2 domain; thread holds rows g=lane/4 (c0,c1) and g+8 (c2,c3)
        const int kbase = t * AB_BK + (lane & 3) * 2;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key = kbase + n * 8 + (j & 1);
                float v = s[n][j] * scale_log2;
                if (key >= nk) v = -INFINITY;
                s[n][j] = v;
                mx[j >> 1] = fmaxf(mx[j >> 1], v);
            }
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_i[r], mx[r]);
            corr[r] = exp2f(m_i[r] - m_new);
            m_i[r] = m_new;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pf[4][4];   // P as A fragments: k-step = 16 keys = two n-tiles
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float p0 = exp2f(s[n][0] - m_i[0]), p1 = exp2f(s[n][1] - m_i[0]);
            const float p2 = exp2f(s[n][2] - m_i[1]), p3 = exp2f(s[n][3] - m_i[1]);
            rs[0] += p0 + p1;
            rs[1] += p2 + p3;
            const int ks = n >> 1, hi = n & 1;
            pf[ks][hi * 2 + 0] = pack_bf16x2(p0, p1);
            pf[ks][hi * 2 + 1] = pack_bf16x2(p2, p3);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_i[r] = l_i[r] * corr[r] + rs[r];   // per-thread partial; lanes reduced at the end
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            o[n][0] *= corr[0]; o[n][1] *= corr[0];
            o[n][2] *= corr[1]; o[n][3] *= corr[1];
        }
        // O += P V : k = keys (4 k-steps), n = d (8 n-tiles); V is [key][d] -> transposed ldmatrix
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {     // pairs of d n-tiles (16 dims)
                uint32_t vb[4];
                // matrices: (keys ks*16+0..7, d dp*16+0..7), (keys +8, d), (keys, d+8), (keys+8, d+8)
                const int row = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                const int chunk = dp * 2 + (lane >> 4);
                ldmatrix_x4_trans(vb, smem_u32(Vs[buf] + sw_off(row, chunk)));
                mma_bf16_16816(o[dp * 2], pf[ks], vb[0], vb[1]);
                mma_bf16_16816(o[dp * 2 + 1], pf[ks], vb[2], vb[3]);
            }
        }
        __syncthreads();   // everyone done with `buf` before it is refilled two iterations later
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
        l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
    }
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = q0 + warp * 16 + g + r * 8;
        if (row < L) {
            const float inv = l_i[r] > 0.f ? 1.0f / l_i[r] : 0.f;
            __nv_bfloat16* op = out + ((long long)b * L + row) * ldo + h * DH + tq * 2;
#pragma unroll
            for (int n = 0; n < 8; ++n)
                *reinterpret_cast<uint32_t*>(op + n * 8) = pack_bf16x2(o[n][r * 2] * inv, o[n][r * 2 + 1] * inv);
        }
    }
}

// =============================================================================== bf16x3 mma.sync
// fp32-class attention for the grounding chain (TextBert / QTV) on the tensor pipe.  q|k|v arrive as
// bf16 hi|lo pairs (hi at column c, lo at column lo_off + c: the OUT_SPLIT format of t2s_gemm_bf16x3);
// S = Qh.Kh + Qh.Kl + Ql.Kh and O = Ph.Vh + Ph.Vl + Pl.Vh with P split in registers, everything else
// (scale, online softmax, accumulators) in fp32.  Per-product error ~2^-17 instead of 2^-8.
constexpr int AX_SMEM = (2 * AB_BQ + 8 * AB_BK) * 128;     // Q hi/lo + double-buffered K hi/lo, V hi/lo = 80 KB

__global__ void __launch_bounds__(AB_THREADS)
attn_x3_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld, int lo_off, int L, int H,
               const int* __restrict__ key_idx, const int* __restrict__ n_keys, int key_stride,
               __nv_bfloat16* __restrict__ out, long long ldo, float scale_log2) {
    extern __shared__ __align__(128) uint8_t xsm[];
    uint8_t* Qh = xsm;
    uint8_t* Ql = Qh + AB_BQ * 128;
    uint8_t* KV = Ql + AB_BQ * 128;            // [buf][Kh, Kl, Vh, Vl][64 rows x 128 B]
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AB_BQ;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nk = n_keys[b];
    const int* kidx = key_idx + (long long)b * key_stride;
    const __nv_bfloat16* base = qkv + (long long)b * L * ld;
    const int ntiles = (nk + AB_BK - 1) / AB_BK;

    for (int i = tid; i < AB_BQ * 8; i += AB_THREADS) {
        const int r = i >> 3, c = i & 7;
        const bool ok = q0 + r < L;
        const __nv_bfloat16* src = base + (long long)(ok ? q0 + r : 0) * ld + h * DH + c * 8;
        cp_async16(Qh + sw_off(r, c), src, ok);
        cp_async16(Ql + sw_off(r, c), src + lo_off, ok);
    }
    auto load_kv = [&](int t, int buf) {
        uint8_t* dst = KV + buf * (4 * AB_BK * 128);
        for (int i = tid; i < AB_BK * 8; i += AB_THREADS) {
            const int r = i >> 3, c = i & 7;
            const bool ok = t * AB_BK + r < nk;
            const long long row = ok ? kidx[t * AB_BK + r] : 0;
            const __nv_bfloat16* src = base + row * ld + h * DH + c * 8;
            const uint32_t o = sw_off(r, c);
            cp_async16(dst + o, src + H, ok);
            cp_async16(dst + AB_BK * 128 + o, src + H + lo_off, ok);
            cp_async16(dst + 2 * AB_BK * 128 + o, src + 2 * H, ok);
            cp_async16(dst + 3 * AB_BK * 128 + o, src + 2 * H + lo_off, ok);
        }
    };
    if (ntiles > 0) load_kv(0, 0);
    cp_async_commit();

    float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[n][j] = 0.f;
    uint32_t qh[4][4], ql[4][4];

    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) load_kv(t + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const uint8_t* Kh = KV + buf * (4 * AB_BK * 128);
        const uint8_t* Kl = Kh + AB_BK * 128;
        const uint8_t* Vh = Kl + AB_BK * 128;
        const uint8_t* Vl = Vh + AB_BK * 128;
        if (t == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int row = warp * 16 + (lane & 15);
                const int chunk = ks * 2 + (lane >> 4);
                ldmatrix_x4(qh[ks], smem_u32(Qh + sw_off(row, chunk)));
                ldmatrix_x4(ql[ks], smem_u32(Ql + sw_off(row, chunk)));
            }
        }
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[n][j] = 0.f;
#pragma unroll
        for (int np = 0; np < 4; ++np) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t kh[4], kl[4];
                const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int chunk = ks * 2 + ((lane >> 3) & 1);
                const uint32_t off = sw_off(row, chunk);
                ldmatrix_x4(kh, smem_u32(Kh + off));
                ldmatrix_x4(kl, smem_u32(Kl + off));
                // small terms first, the dominant hi.hi product last
                mma_bf16_16816(s[np * 2], ql[ks], kh[0], kh[1]);
                mma_bf16_16816(s[np * 2 + 1], ql[ks], kh[2], kh[3]);
                mma_bf16_16816(s[np * 2], qh[ks], kl[0], kl[1]);
                mma_bf16_16816(s[np * 2 + 1], qh[ks], kl[2], kl[3]);
                mma_bf16_16816(s[np * 2], qh[ks], kh[0], kh[1]);
                mma_bf16_16816(s[np * 2 + 1], qh[ks], kh[2], kh[3]);
            }
        }
        const int kbase = t * AB_BK + (lane & 3) * 2;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key = kbase + n * 8 + (j & 1);
                float v = s[n][j] * scale_log2;
                if (key >= nk) v = -INFINITY;
                s[n][j] = v;
                mx[j >> 1] = fmaxf(mx[j >> 1], v);
            }
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_i[r], mx[r]);
            corr[r] = exp2f(m_i[r] - m_new);
            m_i[r] = m_new;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t ph[4][4], pl[4][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float p0 = exp2f(s[n][0] - m_i[0]), p1 = exp2f(s[n][1] - m_i[0]);
            const float p2 = exp2f(s[n][2] - m_i[1]), p3 = exp2f(s[n][3] - m_i[1]);
            rs[0] += p0 + p1;
            rs[1] += p2 + p3;
            const int ks = n >> 1, hi = n & 1;
            const uint32_t h01 = pack_bf16x2(p0, p1), h23 = pack_bf16x2(p2, p3);
            ph[ks][hi * 2 + 0] = h01;
            ph[ks][hi * 2 + 1] = h23;
            pl[ks][hi * 2 + 0] = pack_bf16x2(p0 - bf16lo(h01), p1 - bf16hi(h01));
            pl[ks][hi * 2 + 1] = pack_bf16x2(p2 - bf16lo(h23), p3 - bf16hi(h23));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_i[r] = l_i[r] * corr[r] + rs[r];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            o[n][0] *= corr[0]; o[n][1] *= corr[0];
            o[n][2] *= corr[1]; o[n][3] *= corr[1];
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                uint32_t vh[4], vl[4];
                const int row = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                const int chunk = dp * 2 + (lane >> 4);
                const uint32_t off = sw_off(row, chunk);
                ldmatrix_x4_trans(vh, smem_u32(Vh + off));
                ldmatrix_x4_trans(vl, smem_u32(Vl + off));
                mma_bf16_16816(o[dp * 2], pl[ks], vh[0], vh[1]);
                mma_bf16_16816(o[dp * 2 + 1], pl[ks], vh[2], vh[3]);
                mma_bf16_16816(o[dp * 2], ph[ks], vl[0], vl[1]);
                mma_bf16_16816(o[dp * 2 + 1], ph[ks], vl[2], vl[3]);
                mma_bf16_16816(o[dp * 2], ph[ks], vh[0], vh[1]);
                mma_bf16_16816(o[dp * 2 + 1], ph[ks], vh[2], vh[3]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
        l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
    }
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = q0 + warp * 16 + g + r * 8;
        if (row < L) {
            const float inv = l_i[r] > 0.f ? 1.0f / l_i[r] : 0.f;
            __nv_bfloat16* op = out + ((long long)b * L + row) * ldo + h * DH + tq * 2;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float a = o[n][r * 2] * inv, c = o[n][r * 2 + 1] * inv;
                const uint32_t hi = pack_bf16x2(a, c);
                *reinterpret_cast<uint32_t*>(op + n * 8) = hi;
                *reinterpret_cast<uint32_t*>(op + n * 8 + H) = pack_bf16x2(a - bf16lo(hi), c - bf16hi(hi));
            }
        }
    }
}

// =============================================================================== decoder rows
constexpr int AD_THREADS = 256, AD_MAXQ = 16, AD_WARPS = AD_THREADS / 32;

// Queries: decoder positions t0 .. t0+nq-1 of sample b (rows of `qkv_dec`, [B, T, 3H] bf16).
// Keys: the n_keys[b] valid encoder rows of `qkv_enc` ([B, L_enc, 3H]) followed by decoder
// positions 0 .. t0+nq-1 (causal: query i sees decoder key j iff j <= t0+i).
//
// One CTA per (head, sample).  HBM/L2-bound: every K and V row of the head (128 B each) is read once
// per chunk of QC queries, 8 lanes x 16 B per row so that a warp load covers four full 128-byte rows;
// the key loops are unrolled four deep to keep 16 lines in flight per warp.  Scores are staged in
// shared memory, softmax is one warp per query row (shuffle reductions), the P.V partial sums of
// the 32 key slices are reduced by shuffle + shared memory.
template <int QC>
__global__ void __launch_bounds__(AD_THREADS)
attn_dec_kernel(const __nv_bfloat16* __restrict__ qkv_enc, long long ld_enc, int L_enc,
                const __nv_bfloat16* __restrict__ qkv_dec, long long ld_dec, int T, int H,
                const int* __restrict__ key_idx, const int* __restrict__ n_keys, int key_stride,
                int t0, int nq, __nv_bfloat16* __restrict__ out, long long ldo, float scale, int max_keys,
                DropCfg drop, float* __restrict__ lse_out) {
    extern __shared__ __align__(16) float dsm[];
    float* Qs = dsm;                                   // [AD_MAXQ][64], pre-scaled
    float* Ss = Qs + AD_MAXQ * DH;                     // [QC][max_keys]
    float* Os = Ss + QC * max_keys;                    // [AD_WARPS][QC][64]
    int* s_idx = reinterpret_cast<int*>(Os + AD_WARPS * QC * DH);   // [max_keys] row offsets of the encoder keys
    const int b = blockIdx.y, h = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = lane & 7, sub = lane >> 3;           // 16-byte chunk of the 128-byte row / row within the warp load
    pdl_wait();            // one-row launches of the greedy chain come in early (common.cuh)
    pdl_release();
    const int n_enc = n_keys[b];
    const int nk = n_enc + t0 + nq;
    const int* kidx = key_idx + (long long)b * key_stride;
    const __nv_bfloat16* enc = qkv_enc + (long long)b * L_enc * ld_enc + h * DH;
    const __nv_bfloat16* dec = qkv_dec + (long long)b * T * ld_dec + h * DH;

    for (int i = tid; i < nq * DH; i += AD_THREADS) {
        const int q = i / DH, d = i % DH;
        Qs[q * DH + d] = __bfloat162float(dec[(long long)(t0 + q) * ld_dec + d]) * scale;
    }
    for (int k = tid; k < n_enc; k += AD_THREADS) s_idx[k] = kidx[k];
    __syncthreads();

    auto row_ptr = [&](int k) -> const __nv_bfloat16* {
        return k < n_enc ? enc + (long long)s_idx[k] * ld_enc : dec + (long long)(k - n_enc) * ld_dec;
    };

    for (int q0 = 0; q0 < nq; q0 += QC) {
        // ---- scores: S[q][k] = q . K[k]
        float qf[QC][8];
#pragma unroll
        for (int q = 0; q < QC; ++q)
#pragma unroll
            for (int e = 0; e < 8; ++e) qf[q][e] = (q0 + q < nq) ? Qs[(q0 + q) * DH + c * 8 + e] : 0.f;
        for (int k0 = warp * 4; k0 < nk; k0 += AD_WARPS * 4 * 4) {     // warp-uniform trip count (shuffles inside)
            const int kb = k0 + sub;
            uint4 kv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = kb + u * AD_WARPS * 4;
                kv[u] = k < nk ? *reinterpret_cast<const uint4*>(row_ptr(k) + H + c * 8) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = kb + u * AD_WARPS * 4;
                const float kf[8] = {bf16lo(kv[u].x), bf16hi(kv[u].x), bf16lo(kv[u].y), bf16hi(kv[u].y),
                                     bf16lo(kv[u].z), bf16hi(kv[u].z), bf16lo(kv[u].w), bf16hi(kv[u].w)};
#pragma unroll
                for (int q = 0; q < QC; ++q) {
                    float a = 0.f;
#pragma unroll
                    for (int e = 0; e < 8; ++e) a = fmaf(qf[q][e], kf[e], a);
                    a += __shfl_xor_sync(0xffffffffu, a, 1);
                    a += __shfl_xor_sync(0xffffffffu, a, 2);
                    a += __shfl_xor_sync(0xffffffffu, a, 4);
                    if (c == 0 && k < nk) {
                        const int j = k - n_enc;          // decoder key position (causal)
                        Ss[q * max_keys + k] = (j > t0 + q0 + q) ? -INFINITY : a;
                    }
                }
            }
        }
        __syncthreads();
        // ---- softmax: one warp per query row
        for (int q = warp; q < QC; q += AD_WARPS) {
            if (q0 + q >= nq) continue;
            float* srow = Ss + q * max_keys;
            float mx = -INFINITY;
            for (int k = lane; k < nk; k += 32) mx = fmaxf(mx, srow[k]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (int k = lane; k < nk; k += 32) {
                const float p = expf(srow[k] - mx);
                srow[k] = p;
                sum += p;
            }
            sum = warp_sum(sum);
            const float inv = 1.0f / sum;
            // training step: the row's log2-sum-exp in the {lse2, D} layout of t2s_attn_bwd's statistics buffer
            if (lse_out && lane == 0)
                lse_out[(((long long)b * gridDim.x + h) * (L_enc + T) + L_enc + t0 + q0 + q) * 2] =
                    (mx + logf(sum)) * 1.4426950408889634f;
            if (drop.thr) {
                // attention_probs dropout (training step): query = decoder position L_enc + t of the virtual sequence,
                // key k = position in [compacted encoder keys; decoder keys] -- as t2s_attn_bwd_dropout recomputes it
                const uint32_t y = drop_attn_y(drop, b * gridDim.x + h);
                const int qi = L_enc + t0 + q0 + q;
                for (int k = lane; k < nk; k += 32) {
                    const uint32_t hsh = drop_hash(drop.s0, drop.s1, drop_attn_x(qi, k), y);
                    const uint32_t u = (k & 1) ? (hsh >> 16) : (hsh & 0xffffu);
                    srow[k] = u >= drop.thr ? srow[k] * inv * drop.scale : 0.f;
                }
            } else {
                for (int k = lane; k < nk; k += 32) srow[k] *= inv;
            }
        }
        __syncthreads();
        // ---- O[q] = sum_k P[q][k] V[k]; this thread: dims c*8..c*8+7 of key slice (warp*4+sub) mod 32
        float acc[QC][8];
#pragma unroll
        for (int q = 0; q < QC; ++q)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[q][e] = 0.f;
        for (int k0 = warp * 4; k0 < nk; k0 += AD_WARPS * 4 * 4) {
            const int kb = k0 + sub;
            uint4 vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = kb + u * AD_WARPS * 4;
                vv[u] = k < nk ? *reinterpret_cast<const uint4*>(row_ptr(k) + 2 * H + c * 8) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = kb + u * AD_WARPS * 4;
                if (k < nk) {
                    const float vf[8] = {bf16lo(vv[u].x), bf16hi(vv[u].x), bf16lo(vv[u].y), bf16hi(vv[u].y),
                                         bf16lo(vv[u].z), bf16hi(vv[u].z), bf16lo(vv[u].w), bf16hi(vv[u].w)};
#pragma unroll
                    for (int q = 0; q < QC; ++q) {
                        const float p = Ss[q * max_keys + k];
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[q][e] = fmaf(p, vf[e], acc[q][e]);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < QC; ++q)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float a = acc[q][e];
                a += __shfl_xor_sync(0xffffffffu, a, 8);
                a += __shfl_xor_sync(0xffffffffu, a, 16);
                if (sub == 0) Os[(warp * QC + q) * DH + c * 8 + e] = a;
            }
        __syncthreads();
        for (int i = tid; i < QC * DH; i += AD_THREADS) {
            const int q = i / DH, d = i % DH;
            if (q0 + q < nq) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < AD_WARPS; ++w) v += Os[(w * QC + q) * DH + d];
                out[((long long)b * T + t0 + q0 + q) * ldo + h * DH + d] = __float2bfloat16_rn(v);
            }
        }
        __syncthreads();     // Ss / Os are reused by the next query chunk
    }
}

}  // namespace t2s

using namespace t2s;

extern "C" int t2s_attn_f32(const float* qkv, long long ld, int B, int L, int H, int heads, const int* key_idx,
                            const int* n_keys, int key_stride, float* out, long long ldo, void* out_split,
                            long long ldo_split, void* stream) {
    if ((!out && !out_split) || (out_split && (ldo_split < 2LL * H || (ldo_split % 4)))) {
        set_error("attn_f32: needs an output (split pitch >= 2H)");
        return T2S_ERR_ARG;
    }
    if (H != heads * DH || (ld % 4) || (ldo % 4)) { set_error("attn_f32: head size must be 64 (H %d heads %d)", H, heads); return T2S_ERR_SHAPE; }
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM);
        if (e != cudaSuccess) { set_error("attn_f32 attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    dim3 grid((L + AF_BQ - 1) / AF_BQ, heads, B);
    attn_f32_kernel<<<grid, AF_THREADS, AF_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(
        qkv, ld, L, H, key_idx, n_keys, key_stride, out, ldo, reinterpret_cast<__nv_bfloat16*>(out_split), ldo_split,
        0.125f);
    return launch_status("attn_f32");
}

extern "C" int t2s_attn_bf16(const void* qkv, long long ld, int B, int L, int H, int heads, const int* key_idx,
                             const int* n_keys, int key_stride, void* out, long long ldo, void* stream) {
    if (H != heads * DH || (ld % 8) || (ldo % 2)) { set_error("attn_bf16: head size must be 64 (H %d heads %d)", H, heads); return T2S_ERR_SHAPE; }
    dim3 grid((L + AB_BQ - 1) / AB_BQ, heads, B);
    attn_bf16_kernel<<<grid, AB_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv), ld, L, H, key_idx, n_keys, key_stride,
        reinterpret_cast<__nv_bfloat16*>(out), ldo, 0.125f * 1.4426950408889634f);
    return launch_status("attn_bf16");
}

extern "C" int t2s_attn_x3(const void* qkv, long long ld, int lo_off, int B, int L, int H, int heads,
                           const int* key_idx, const int* n_keys, int key_stride, void* out_split, long long ldo,
                           void* stream) {
    if (H != heads * DH || (ld % 8) || (lo_off % 8) || lo_off < 3 * H || ld < lo_off + 3 * H || ldo < 2LL * H || (ldo % 2)) {
        set_error("attn_x3: bad layout (H %d heads %d ld %lld lo_off %d ldo %lld)", H, heads, ld, lo_off, ldo);
        return T2S_ERR_SHAPE;
    }
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AX_SMEM);
        if (e != cudaSuccess) { set_error("attn_x3 attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    dim3 grid((L + AB_BQ - 1) / AB_BQ, heads, B);
    attn_x3_kernel<<<grid, AB_THREADS, AX_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv), ld, lo_off, L, H, key_idx, n_keys, key_stride,
        reinterpret_cast<__nv_bfloat16*>(out_split), ldo, 0.125f * 1.4426950408889634f);
    return launch_status("attn_x3");
}

namespace t2s {
// csrc/attn_bwd.cu: mma.sync flash kernel for several decoder rows per sample
int launch_attn_dec_mma(const void* qkv_enc, long long ld_enc, int L_enc, const void* qkv_dec, long long ld_dec, int T,
                        int B, int H, int heads, const int* key_idx, const int* n_keys, int key_stride, int t0, int nq,
                        void* out, long long ldo, cudaStream_t st, DropCfg drop, float* lse_out);
}

static int attn_dec_entry(const void* qkv_enc, long long ld_enc, int L_enc, const void* qkv_dec, long long ld_dec,
                          int T, int B, int H, int heads, const int* key_idx, const int* n_keys, int key_stride,
                          int t0, int nq, void* out, long long ldo, void* stream, DropCfg drop, float* lse_out = nullptr) {
    if (H != heads * DH || nq < 1 || nq > AD_MAXQ || t0 < 0 || t0 + nq > T || (ld_enc % 8) || (ld_dec % 8)) {
        set_error("attn_dec: bad arguments (H %d heads %d t0 %d nq %d T %d)", H, heads, t0, nq, T);
        return T2S_ERR_SHAPE;
    }
    static const int use_mma = []() { const char* e = getenv("T2S_ATTN_DEC_MMA"); return (e && e[0] == '0') ? 0 : 1; }();
    if (nq > 1 && use_mma && (ldo % 2) == 0)
        // several decoder rows of a sample (teacher-forced passes): tensor-core kernel, K / V read once
        return launch_attn_dec_mma(qkv_enc, ld_enc, L_enc, qkv_dec, ld_dec, T, B, H, heads, key_idx, n_keys, key_stride, t0,
                                   nq, out, ldo, reinterpret_cast<cudaStream_t>(stream), drop, lse_out);
    const int max_keys = ((L_enc + T + 3) / 4) * 4;
    // query chunk: every K / V row is read once per chunk of QC queries.  QC = 12 (all teacher-forced rows in one
    // pass) was measured SLOWER than three passes of QC = 4 on B200 (195 vs 95 us per launch at 1056 keys: 185
    // registers and 84 KB of score rows leave one CTA per SM), so it is only used when asked for (T2S_ATTN_DEC_QC=12)
    static const int qc_big = []() { const char* e = getenv("T2S_ATTN_DEC_QC"); return e && atoi(e) == 12 ? 12 : 4; }();
    int qc = nq == 1 ? 1 : (nq <= 4 ? 4 : qc_big);
    auto smem_for = [&](int q) { return (AD_MAXQ * DH + q * max_keys + AD_WARPS * q * DH + max_keys) * 4; };
    // long videos (stress sweep: up to 256 frames x 60 OCR slots = 15.6 k keys): the score rows of a 4-query chunk no
    // longer fit shared memory -> one query per chunk (K / V are re-read per query; the rows stay L2-resident)
    while (qc > 1 && smem_for(qc) > 200 * 1024) qc = qc == 12 ? 4 : 1;
    const int slot = qc == 1 ? 0 : (qc == 4 ? 1 : 2);
    const int smem = smem_for(qc);
    if (smem > 200 * 1024) { set_error("attn_dec: %d keys exceed the shared-memory score buffer", max_keys); return T2S_ERR_SHAPE; }
    static int attr_bytes[3] = {48 * 1024, 48 * 1024, 48 * 1024};
    if (smem > attr_bytes[slot]) {
        cudaError_t e = qc == 1
            ? cudaFuncSetAttribute(attn_dec_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
            : qc == 4 ? cudaFuncSetAttribute(attn_dec_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                      : cudaFuncSetAttribute(attn_dec_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("attn_dec attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_bytes[slot] = smem;
    }
    dim3 grid(heads, B);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const __nv_bfloat16* pe = reinterpret_cast<const __nv_bfloat16*>(qkv_enc);
    const __nv_bfloat16* pd = reinterpret_cast<const __nv_bfloat16*>(qkv_dec);
    __nv_bfloat16* po = reinterpret_cast<__nv_bfloat16*>(out);
    if (qc == 1)
        launch_pdl(true, attn_dec_kernel<1>, grid, dim3(AD_THREADS), (size_t)smem, st, pe, ld_enc, L_enc, pd, ld_dec, T, H,
                   key_idx, n_keys, key_stride, t0, nq, po, ldo, 0.125f, max_keys, drop, lse_out);
    else if (qc == 4)
        attn_dec_kernel<4><<<grid, AD_THREADS, smem, st>>>(pe, ld_enc, L_enc, pd, ld_dec, T, H, key_idx, n_keys,
                                                           key_stride, t0, nq, po, ldo, 0.125f, max_keys, drop, lse_out);
    else
        attn_dec_kernel<12><<<grid, AD_THREADS, smem, st>>>(pe, ld_enc, L_enc, pd, ld_dec, T, H, key_idx, n_keys,
                                                            key_stride, t0, nq, po, ldo, 0.125f, max_keys, drop, lse_out);
    return launch_status("attn_dec");
}

extern "C" int t2s_attn_dec(const void* qkv_enc, long long ld_enc, int L_enc, const void* qkv_dec, long long ld_dec,
                            int T, int B, int H, int heads, const int* key_idx, const int* n_keys, int key_stride,
                            int t0, int nq, void* out, long long ldo, void* stream) {
    return attn_dec_entry(qkv_enc, ld_enc, L_enc, qkv_dec, ld_dec, T, B, H, heads, key_idx, n_keys, key_stride, t0, nq,
                          out, ldo, stream, DropCfg{0, 0, 0, 0, 1.f});
}

/* t2s_attn_dec of the training step (decoder rows = positions L_enc + t of the virtual sequence): attention_probs
 * dropout (p may be 0) and, when lse_out != null, the rows' log2-sum-exp at
 * lse_out[((b * heads + h) * (L_enc + T) + L_enc + t) * 2] */
extern "C" int t2s_attn_dec_dropout(const void* qkv_enc, long long ld_enc, int L_enc, const void* qkv_dec,
                                    long long ld_dec, int T, int B, int H, int heads, const int* key_idx,
                                    const int* n_keys, int key_stride, int t0, int nq, void* out, long long ldo, float p,
                                    unsigned long long seed, unsigned site, float* lse_out, void* stream) {
    if (p < 0.f || p >= 1.f || L_enc + T > 65535) { set_error("attn_dec_dropout: p in [0, 1), L < 65536"); return T2S_ERR_ARG; }
    return attn_dec_entry(qkv_enc, ld_enc, L_enc, qkv_dec, ld_dec, T, B, H, heads, key_idx, n_keys, key_stride, t0, nq,
                          out, ldo, stream, make_drop(p, seed, site), lse_out);
}
