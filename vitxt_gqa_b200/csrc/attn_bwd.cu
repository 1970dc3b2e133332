// K8c: backward of the masked multi-head attention (head size 64), recompute style -- no probabilities are saved.
//
// Replaces what autograd runs for BertSelfAttention's matmul / +mask / softmax / matmul chain (pytorch_transformers
// modeling_bert, reached from reference pythia/models/t2s.py:423,538,622) under loss.backward()
// (pythia/trainers/base_trainer.py:264), for both mask shapes of the reference: the key-padding mask of TextBert / QTV
// (t2s.py:413-419,533-534) and the prefix-LM mask of the answer transformer (t2s.py:609-618).
//
// Per (sample, head) the problem is a "virtual sequence": queries i in [0, Le + T) -- Le encoder rows then T decoder
// rows (T may be 0) -- and keys j in [0, nk + T): the nk valid encoder keys of the compacted key list, then the
// decoder rows.  allowed(i, j) = j < nk  ||  (i >= Le && j - nk <= i - Le).  Masked keys are skipped exactly as in the
// forward kernels (exp(-10000 - max) == 0 in fp32).
//
// Kernels, all bf16 mma.sync m16n8k16 with fp32 accumulation, 64 x 64 tiles, K/V (or Q/dO) tiles double buffered with
// cp.async into XOR-swizzled shared memory:
//   attn_bwd_stats   lse2[i] = log2 sum_j exp2(s_ij) (s in the exp2 domain) and D[i] = dO_i . O_i
//   attn_bwd_dkv     per 64-key tile, loop over query tiles: S^T = K Q^T, dP^T = V dO^T, P^T = exp2(S^T - lse),
//                    dS^T = P^T (dP^T - D) / 8; dV += P^T dO, dK += dS^T Q -- and, from the same dS^T (transposed through
//                    shared memory), this key tile's share of dQ = dS K, added into an fp32 buffer with red.global.add
//                    (the key tiles of a query row run in different CTAs); attn_bwd_dq_store converts it to bf16.
//                    Five products per tile pair instead of the seven of a separate dQ pass (S and dP were recomputed).
//   attn_bwd_dq      the separate, atomic-free dQ pass (per 64-query tile, loop over key tiles); kept for
//                    T2S_ATTN_BWD_FUSED=0 (run-to-run identical results; the fp32 adds of the fused form commute only
//                    up to rounding).
// dK / dV rows are written by exactly one CTA (rows that are not in a key list keep the zeros the entry point memsets).
#include "common.cuh"
#include <stdlib.h>
#include "../../include/t2s_b200.h"

namespace t2s {

constexpr int XB = 64;               // tile edge (queries and keys)
constexpr int XB_THREADS = 128;      // 4 warps x 16 rows
constexpr int XDH = 64;

__device__ __forceinline__ void xb_ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void xb_ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void xb_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t xb_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

struct AttnBwdArgs {
    const __nv_bfloat16* qkv_enc; long long ld_enc;        // [B*Le, >= 3H]: q | k | v at columns 0, H, 2H
    const __nv_bfloat16* qkv_dec; long long ld_dec;        // [B*T, >= 3H] (null when T == 0)
    const __nv_bfloat16* o_enc; long long ldo_enc;         // forward output (context) rows
    const __nv_bfloat16* o_dec; long long ldo_dec;
    const __nv_bfloat16* do_enc; long long ldg_enc;        // gradient of the context rows
    const __nv_bfloat16* do_dec; long long ldg_dec;
    __nv_bfloat16* dqkv_enc; long long ldq_enc;            // outputs: dq | dk | dv, same column layout
    __nv_bfloat16* dqkv_dec; long long ldq_dec;
    const int* key_idx; const int* n_keys; int key_stride;
    float* stats;                                           // [B, heads, Le + T, 2] = {lse2, D}
    float* dq32;                                            // fused form: [B * (Le + T), H] fp32 dQ accumulator (zeroed)
    int have_lse;                                           // stats[..][0] was written by the forward kernels
    int Le, T, H, heads;
    float scale_log2, scale;
    DropCfg drop;                                           // attention_probs dropout of the forward (thr == 0: none)
};

// row resolvers of the virtual sequence (sample b); column offset `col` in elements
__device__ __forceinline__ const __nv_bfloat16* xb_qrow(const AttnBwdArgs& a, const __nv_bfloat16* enc, long long ld_e,
                                                        const __nv_bfloat16* dec, long long ld_d, int b, int i) {
    return i < a.Le ? enc + ((long long)b * a.Le + i) * ld_e : dec + ((long long)b * a.T + (i - a.Le)) * ld_d;
}
__device__ __forceinline__ long long xb_krow_index(const AttnBwdArgs& a, const int* kidx, int nk, int j, bool& is_dec) {
    is_dec = j >= nk;
    return is_dec ? (long long)(j - nk) : (long long)kidx[j];
}

// load a [64 x 64] bf16 tile of query-side rows i0.. (zero-filled past n_rows) from the (enc | dec) pair of buffers
__device__ __forceinline__ void xb_load_qtile(uint8_t* dst, const AttnBwdArgs& a, const __nv_bfloat16* enc, long long ld_e,
                                              const __nv_bfloat16* dec, long long ld_d, int b, int i0, int n_rows, int col) {
    for (int t = threadIdx.x; t < XB * 8; t += XB_THREADS) {
        const int r = t >> 3, c = t & 7;
        const bool ok = i0 + r < n_rows;
        const __nv_bfloat16* src = xb_qrow(a, enc, ld_e, dec, ld_d, b, ok ? i0 + r : 0) + col + c * 8;
        cp_async16(dst + xb_off(r, c), src, ok);
    }
}
// load a [64 x 64] tile of key-side rows j0.. of the virtual key list
__device__ __forceinline__ void xb_load_ktile(uint8_t* dst, const AttnBwdArgs& a, const __nv_bfloat16* enc, long long ld_e,
                                              const __nv_bfloat16* dec, long long ld_d, int b, const int* kidx, int nk,
                                              int j0, int n_keys_v, int col) {
    for (int t = threadIdx.x; t < XB * 8; t += XB_THREADS) {
        const int r = t >> 3, c = t & 7;
        const bool ok = j0 + r < n_keys_v;
        bool is_dec = false;
        const long long row = ok ? xb_krow_index(a, kidx, nk, j0 + r, is_dec) : 0;
        const __nv_bfloat16* src = (is_dec ? dec + ((long long)b * a.T + row) * ld_d : enc + ((long long)b * a.Le + row) * ld_e) + col + c * 8;
        cp_async16(dst + xb_off(r, c), src, ok);
    }
}
__device__ __forceinline__ bool xb_allowed(int i, int j, int Le, int nk) { return j < nk || (i >= Le && (j - nk) <= (i - Le)); }

// A fragments (4 k-steps over the 64 head dims) of this warp's 16 rows of a [64 x 64] tile
__device__ __forceinline__ void xb_load_afrags(uint32_t (&f)[4][4], const uint8_t* tile, int warp, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
        xb_ldmatrix_x4(f[ks], smem_u32(tile + xb_off(warp * 16 + (lane & 15), ks * 2 + (lane >> 4))));
}
// C[16 x 64] = A[16 x 64(d)] . Bt, Bt tile stored [n][d] (n = 64 rows of the tile): c[n8][4]
__device__ __forceinline__ void xb_mma_nt(float (&c)[8][4], const uint32_t (&af)[4][4], const uint8_t* tile, int lane) {
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[n][j] = 0.f;
#pragma unroll
    for (int np = 0; np < 4; ++np)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t bf[4];
            const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
            const int chunk = ks * 2 + ((lane >> 3) & 1);
            xb_ldmatrix_x4(bf, smem_u32(tile + xb_off(row, chunk)));
            xb_mma(c[np * 2], af[ks], bf[0], bf[1]);
            xb_mma(c[np * 2 + 1], af[ks], bf[2], bf[3]);
        }
}
// acc[16 x 64(d)] += P[16 x 64(k)] . B, B tile stored [k][d]: pf = A fragments over the 64 k rows
__device__ __forceinline__ void xb_mma_nn(float (&acc)[8][4], const uint32_t (&pf)[4][4], const uint8_t* tile, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
            uint32_t bf[4];
            const int row = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
            const int chunk = dp * 2 + (lane >> 4);
            xb_ldmatrix_x4_trans(bf, smem_u32(tile + xb_off(row, chunk)));
            xb_mma(acc[dp * 2], pf[ks], bf[0], bf[1]);
            xb_mma(acc[dp * 2 + 1], pf[ks], bf[2], bf[3]);
        }
}
// C tile (fp32, [16 x 64]) -> A fragments for the next product
__device__ __forceinline__ void xb_c_to_a(uint32_t (&pf)[4][4], const float (&c)[8][4]) {
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const int ks = n >> 1, hi = n & 1;
        pf[ks][hi * 2 + 0] = pack_bf16x2(c[n][0], c[n][1]);
        pf[ks][hi * 2 + 1] = pack_bf16x2(c[n][2], c[n][3]);
    }
}

// ------------------------------------------------------------------------------- stats: lse2 and D per query
__global__ void __launch_bounds__(XB_THREADS)
attn_bwd_stats_kernel(AttnBwdArgs a) {
    __shared__ __align__(128) uint8_t Qs[XB * 128];
    __shared__ __align__(128) uint8_t Ks[2][XB * 128];
    const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * XB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int nq = a.Le + a.T;
    const int nk = a.n_keys[b];
    const int* kidx = a.key_idx + (long long)b * a.key_stride;
    // only tiles that contain decoder rows can see decoder keys
    const int nkv = (i0 + XB > a.Le) ? nk + a.T : nk;
    const int ntiles = (nkv + XB - 1) / XB;
    const int col = h * XDH;

    float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
    uint32_t qf[4][4];
    if (!a.have_lse) {
        xb_load_qtile(Qs, a, a.qkv_enc, a.ld_enc, a.qkv_dec, a.ld_dec, b, i0, nq, col);
        xb_load_ktile(Ks[0], a, a.qkv_enc, a.ld_enc, a.qkv_dec, a.ld_dec, b, kidx, nk, 0, nkv, a.H + col);
        cp_async_commit();
    }
    for (int t = 0; t < (a.have_lse ? 0 : ntiles); ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles)
            xb_load_ktile(Ks[buf ^ 1], a, a.qkv_enc, a.ld_enc, a.qkv_dec, a.ld_dec, b, kidx, nk, (t + 1) * XB, nkv, a.H + col);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (t == 0) xb_load_afrags(qf, Qs, warp, lane);
        float s[8][4];
        xb_mma_nt(s, qf, Ks[buf], lane);
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key = t * XB + n * 8 + tq * 2 + (j & 1);
                const int qi = i0 + warp * 16 + g + (j >> 1) * 8;
                float v = s[n][j] * a.scale_log2;
                if (key >= nkv || !xb_allowed(qi, key, a.Le, nk)) v = -INFINITY;
                s[n][j] = v;
                mx[j >> 1] = fmaxf(mx[j >> 1], v);
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_i[r], mx[r]);
            const float corr = m_new == -INFINITY ? 1.f : exp2f(m_i[r] - m_new);
            float rs = 0.f;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                if (m_new != -INFINITY) rs += exp2f(s[n][r * 2] - m_new) + exp2f(s[n][r * 2 + 1] - m_new);
            }
            l_i[r] = l_i[r] * corr + rs;
            m_i[r] = m_new;
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
        l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
    }
    // D = dO . O: lane quartet shares a row; each lane sums 16 of the 64 dims
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int qi = i0 + warp * 16 + g + r * 8;
        float d = 0.f;
        if (qi < nq) {
            const __nv_bfloat16* op = xb_qrow(a, a.o_enc, a.ldo_enc, a.o_dec, a.ldo_dec, b, qi) + col + tq * 16;
            const __nv_bfloat16* gp = xb_qrow(a, a.do_enc, a.ldg_enc, a.do_dec, a.ldg_dec, b, qi) + col + tq * 16;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint4 ov = *reinterpret_cast<const uint4*>(op + c * 8);
                const uint4 gv = *reinterpret_cast<const uint4*>(gp + c * 8);
                d += bf16lo(ov.x) * bf16lo(gv.x) + bf16hi(ov.x) * bf16hi(gv.x) + bf16lo(ov.y) * bf16lo(gv.y) + bf16hi(ov.y) * bf16hi(gv.y)
                   + bf16lo(ov.z) * bf16lo(gv.z) + bf16hi(ov.z) * bf16hi(gv.z) + bf16lo(ov.w) * bf16lo(gv.w) + bf16hi(ov.w) * bf16hi(gv.w);
            }
        }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        if (qi < nq && tq == 0) {
            float* st = a.stats + (((long long)b * a.heads + h) * nq + qi) * 2;
            if (!a.have_lse) st[0] = m_i[r] + log2f(l_i[r]);
            st[1] = d;
        }
    }
}

// ------------------------------------------------------------------------------- dQ
__global__ void __launch_bounds__(XB_THREADS)
attn_bwd_dq_kernel(AttnBwdArgs a) {
    __shared__ __align__(128) uint8_t QG[2][XB * 128];       // Q tile, dO tile (only read at t == 0)
    __shared__ __align__(128) uint8_t Ks[2][XB * 128];
    __shared__ __align__(128) uint8_t Vs[2][XB * 128];
    const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * XB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int nq = a.Le + a.T;
    const int nk = a.n_keys[b];
    const int* kidx = a.key_idx + (long long)b * a.key_stride;
    const int nkv = (i0 + XB > a.Le) ? nk + a.T : nk;
    const int ntiles = (nkv + XB - 1) / XB;
    const int col = h * XDH;

    xb_load_qtile(QG[0], a, a.qkv_enc, a.ld_enc, a.qkv_dec, a.ld_dec, b, i0, nq, col);
    xb_load_qtile(QG[1], a, a.do_enc, a.ldg_enc, a.do_dec, a.ldg_dec, b, i0, nq, col);
    auto load_kv = [&](int t, int buf) {
        xb_load_ktile(Ks[buf], a, a.qkv_enc, a.ld_enc, a.qkv_dec, a.ld_dec, b, kidx, nk, t * XB, nkv, a.H + col);
        xb_load_ktile(Vs[buf], a, a.qkv_enc, a.ld_enc, a.qkv_dec, a.ld_dec, b, kidx, nk, t * XB, nkv, 2 * a.H + col);
    };
    load_kv(0, 0);
    cp_async_commit();
    float lse[2], dd[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int qi = i0 + warp * 16 + g + r * 8;
        const float* st = a.stats + (((long long)b * a.heads + h) * nq + (qi < nq ? qi : 0)) * 2;
        lse[r] = st[0];
        dd[r] = st[1];
    }
    float dq[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) dq[n][j] = 0.f;
    uint32_t qf[4][4], gf[4][4];
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) load_kv(t + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (t == 0) {
            xb_load_afrags(qf, QG[0], warp, lane);
            xb_load_afrags(gf, QG[1], warp, lane);
        }
        float s[8][4], dp[8][4];
        xb_mma_nt(s, qf, Ks[buf], lane);
        xb_mma_nt(dp, gf, Vs[buf], lane);
        uint32_t hsh = 0;
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key = t * XB + n * 8 + tq * 2 + (j & 1);
                const int r = j >> 1;
                const int qi = i0 + warp * 16 + g + r * 8;
                const bool ok = key < nkv && xb_allowed(qi, key, a.Le, nk);
                const float p = ok ? exp2f(s[n][j] * a.scale_log2 - lse[r]) : 0.f;
                float dpj = dp[n][j];
                if (a.drop.thr) {               // O = sum_j p_j m_j v_j: dL/dp_j = m_j (dO . v_j); D = dO . O is unchanged
                    // accumulator columns j = 0|1 (2|3) are the two keys of one mask pair: one hash serves both
                    if ((j & 1) == 0)
                        hsh = drop_hash(a.drop.s0, a.drop.s1, drop_attn_x(qi, key), drop_attn_y(a.drop, b * a.heads + h));
                    const uint32_t u = (j & 1) ? (hsh >> 16) : (hsh & 0xffffu);
                    dpj = u >= a.drop.thr ? dpj * a.drop.scale : 0.f;
                }
                s[n][j] = p * (dpj - dd[r]) * a.scale;          // dS
            }
        uint32_t dsf[4][4];
        xb_c_to_a(dsf, s);
        xb_mma_nn(dq, dsf, Ks[buf], lane);
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int qi = i0 + warp * 16 + g + r * 8;
        if (qi < nq) {
            __nv_bfloat16* op = (qi < a.Le ? a.dqkv_enc + ((long long)b * a.Le + qi) * a.ldq_enc
                                           : a.dqkv_dec + ((long long)b * a.T + (qi - a.Le)) * a.ldq_dec) + col + tq * 2;
#pragma unroll
            for (int n = 0; n < 8; ++n)
                *reinterpret_cast<uint32_t*>(op + n * 8) = pack_bf16x2(dq[n][r * 2], dq[n][r * 2 + 1]);
        }
    }
}

// ------------------------------------------------------------------------------- dK, dV
constexpr int XB_DKV_SMEM = 2 * XB * 128 /*K,V own tiles*/ + 2 * 2 * XB * 128 /*Q,dO x 2 stages*/ + 2 * 2 * XB * 4 /*stats*/
                            + XB * 128 /*dS^T tile of the fused dQ product*/;

__global__ void __launch_bounds__(XB_THREADS)
attn_bwd_dkv_kernel(AttnBwdArgs a) {
    extern __shared__ __align__(128) uint8_t xsm[];
    uint8_t* KV = xsm;                                   // [2][64*128]: this CTA's K tile, V tile
    uint8_t* Qs = KV + 2 * XB * 128;                     // [2 stages][64*128]
    uint8_t* Gs = Qs + 2 * XB * 128;                     // [2 stages][64*128] dO
    float* Ss = reinterpret_cast<float*>(Gs + 2 * XB * 128);     // [2 stages][2][64]: lse2, D
    uint8_t* Ts = reinterpret_cast<uint8_t*>(Ss + 2 * 2 * XB);   // [64 keys][64 queries] bf16: dS^T of this iteration
    const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * XB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int nq = a.Le + a.T;
    const int nk = a.n_keys[b];
    const int nkv = nk + a.T;
    if (j0 >= nkv) return;
    const int* kidx = a.key_idx + (long long)b * a.key_stride;
    const int col = h * XDH;
    // a key tile made only of decoder keys is seen by decoder queries only
    const int i_first = (j0 >= nk) ? (a.Le / XB) * XB : 0;
    const int ntiles = (nq - i_first + XB - 1) / XB;

    xb_load_ktile(KV, a, a.qkv_enc, a.ld_enc, a.qkv_dec, a.ld_dec, b, kidx, nk, j0, nkv, a.H + col);
    xb_load_ktile(KV + XB * 128, a, a.qkv_enc, a.ld_enc, a.qkv_dec, a.ld_dec, b, kidx, nk, j0, nkv, 2 * a.H + col);
    auto load_q = [&](int t, int buf) {
        const int i0 = i_first + t * XB;
        xb_load_qtile(Qs + buf * XB * 128, a, a.qkv_enc, a.ld_enc, a.qkv_dec, a.ld_dec, b, i0, nq, col);
        xb_load_qtile(Gs + buf * XB * 128, a, a.do_enc, a.ldg_enc, a.do_dec, a.ldg_dec, b, i0, nq, col);
        if (threadIdx.x < XB) {
            const int qi = i0 + threadIdx.x;
            const float* st = a.stats + (((long long)b * a.heads + h) * nq + (qi < nq ? qi : 0)) * 2;
            Ss[(buf * 2 + 0) * XB + threadIdx.x] = st[0];
            Ss[(buf * 2 + 1) * XB + threadIdx.x] = st[1];
        }
    };
    load_q(0, 0);
    cp_async_commit();
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) { dk[n][j] = 0.f; dv[n][j] = 0.f; }
    uint32_t kf[4][4], vf[4][4];
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const int i0 = i_first + t * XB;
        if (t + 1 < ntiles) load_q(t + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (t == 0) {
            xb_load_afrags(kf, KV, warp, lane);
            xb_load_afrags(vf, KV + XB * 128, warp, lane);
        }
        const uint8_t* qt = Qs + buf * XB * 128;
        const uint8_t* gt = Gs + buf * XB * 128;
        const float* lse_s = Ss + (buf * 2 + 0) * XB;
        const float* d_s = Ss + (buf * 2 + 1) * XB;
        float st[8][4], dpt[8][4];
        xb_mma_nt(st, kf, qt, lane);          // S^T  [keys x queries]
        xb_mma_nt(dpt, vf, gt, lane);         // dP^T = V . dO^T
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int qc = n * 8 + tq * 2 + (j & 1);
                const int qi = i0 + qc;
                const int key = j0 + warp * 16 + g + (j >> 1) * 8;
                const bool ok = qi < nq && key < nkv && xb_allowed(qi, key, a.Le, nk);
                const float p = ok ? exp2f(st[n][j] * a.scale_log2 - lse_s[qc]) : 0.f;
                float m = 1.f;
                if (a.drop.thr) {
                    // rows of this accumulator are keys: key and key ^ 1 (one mask pair) sit in lanes g and g ^ 1, i.e.
                    // lane ^ 4 -- the even-g lane hashes j = 0, 1 and the odd-g lane j = 2, 3 of both, then they swap
                    const bool mine = ((g & 1) == 0) == (j < 2);
                    uint32_t hsh = 0;          // (the pair index key >> 1 is the same in both lanes)
                    if (mine) hsh = drop_hash(a.drop.s0, a.drop.s1, drop_attn_x(qi, key), drop_attn_y(a.drop, b * a.heads + h));
                    const uint32_t other = __shfl_xor_sync(0xffffffffu, hsh, 4);
                    if (!mine) hsh = other;
                    const uint32_t u = (key & 1) ? (hsh >> 16) : (hsh & 0xffffu);
                    m = u >= a.drop.thr ? a.drop.scale : 0.f;
                }
                st[n][j] = p * m;                                      // (P o M)^T: what multiplied V in the forward
                dpt[n][j] = p * (dpt[n][j] * m - d_s[qc]) * a.scale;   // dS^T
            }
        uint32_t pf[4][4];
        xb_c_to_a(pf, st);
        xb_mma_nn(dv, pf, gt, lane);          // dV += P^T . dO
        xb_c_to_a(pf, dpt);
        xb_mma_nn(dk, pf, qt, lane);          // dK += dS^T . Q
        if (a.dq32) {
            // dQ[q tile] += dS . K[this key tile]: dS^T goes through shared memory ([key][query], this warp's 16 key rows),
            // each warp then takes 16 QUERY rows over all 64 keys as A fragments (ldmatrix.trans) against the K tile
#pragma unroll
            for (int n = 0; n < 8; ++n)
#pragma unroll
                for (int r = 0; r < 2; ++r)
                    *reinterpret_cast<uint32_t*>(Ts + xb_off(warp * 16 + g + r * 8, n) + tq * 4) =
                        pack_bf16x2(dpt[n][r * 2], dpt[n][r * 2 + 1]);
            __syncthreads();
            float dq[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n)
#pragma unroll
                for (int j = 0; j < 4; ++j) dq[n][j] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t af[4];
                // matrices (query 0-7 | 8-15) x (key 0-7 | 8-15) of this warp's 16 queries, stored [key][query]
                xb_ldmatrix_x4_trans(af, smem_u32(Ts + xb_off(ks * 16 + ((lane >> 4) & 1) * 8 + (lane & 7),
                                                               warp * 2 + ((lane >> 3) & 1))));
#pragma unroll
                for (int dp = 0; dp < 4; ++dp) {
                    uint32_t bf[4];
                    const int row = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                    xb_ldmatrix_x4_trans(bf, smem_u32(KV + xb_off(row, dp * 2 + (lane >> 4))));
                    xb_mma(dq[dp * 2], af, bf[0], bf[1]);
                    xb_mma(dq[dp * 2 + 1], af, bf[2], bf[3]);
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int qi = i0 + warp * 16 + g + r * 8;
                if (qi < nq) {
                    float* op = a.dq32 + ((long long)b * nq + qi) * a.H + col + tq * 2;
#pragma unroll
                    for (int n = 0; n < 8; ++n)
                        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(op + n * 8), "f"(dq[n][r * 2]),
                                     "f"(dq[n][r * 2 + 1]) : "memory");
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int key = j0 + warp * 16 + g + r * 8;
        if (key < nkv) {
            bool is_dec = false;
            const long long row = xb_krow_index(a, kidx, nk, key, is_dec);
            __nv_bfloat16* op = (is_dec ? a.dqkv_dec + ((long long)b * a.T + row) * a.ldq_dec
                                        : a.dqkv_enc + ((long long)b * a.Le + row) * a.ldq_enc) + col + tq * 2;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                *reinterpret_cast<uint32_t*>(op + a.H + n * 8) = pack_bf16x2(dk[n][r * 2], dk[n][r * 2 + 1]);
                *reinterpret_cast<uint32_t*>(op + 2 * a.H + n * 8) = pack_bf16x2(dv[n][r * 2], dv[n][r * 2 + 1]);
            }
        }
    }
}


// ------------------------------------------------------------------------------- dK / dV / dQ on tcgen05 + TMEM
// The same fused backward as attn_bwd_dkv_kernel above, on the 5th-gen tensor cores.  One CTA = one 128-key tile of the
// virtual key list of one (sample, head); it loops over the 128-row query tiles that can see it.  Per tile pair:
//   MMA 1   S^T  = K_j Q_i^T   (M 128 keys, N 128 queries, K 64)  -> TMEM columns [0, 128)
//           dP^T = V_j dO_i^T                                     -> TMEM columns [128, 256)
//   warps 0-3 (thread r == key row r == TMEM lane r): P^T = exp2(S^T scale - lse[q]) (mask, validity), dropout
//           multiplier m(q, key) recomputed, P^T m and dS^T = P^T (dP^T m - D[q]) / 8 written as bf16 into two 128B-
//           swizzled [128 keys x 64 queries] shared-memory tiles each
//   MMA 2   dV += (P^T m) dO_i   (M 128 keys, N 64, K 128 queries; dO_i read as an MN-major B operand from the tile
//                                  MMA 1 read K-major)            -> TMEM columns [256, 320), accumulated over i
//           dK += dS^T Q_i                                         -> [320, 384), accumulated over i
//           dQ_i = dS K_j         (M 128 queries: dS^T read as an MN-major A operand, i.e. transposed by the
//                                  descriptor; K_j as MN-major B)  -> [384, 448), read out per tile and added into the
//                                  fp32 dQ buffer with red.global.add (other key tiles add to the same rows)
// Roles: warps 0-15 softmax / read-out (warp w owns TMEM lanes 32 (w % 4) .. + 31 = 32 key rows and the 32 query
// columns of quarter w / 4: one CTA per SM leaves these warps alone with their latencies -- a first version with four
// such warps ran the element-wise pass at ~0.2 instructions per cycle per scheduler and was slower than mma.sync),
// warps 16-19 loaders (cp.async row gathers of K_j, V_j once and of Q_i, dO_i, {lse, D} per query tile, two stages),
// warp 20 issues the MMAs.  162 KB of shared memory, 512 TMEM columns: one CTA per SM; MMA 1 of tile i + 1 runs under
// the dQ read-out of tile i.
constexpr int BT_SW = 16;                          // softmax warps
constexpr int BT_THREADS = 32 * (BT_SW + 5);
constexpr int BT_TILE = 128 * 128;                 // bytes of a [128 rows x 64 bf16] 128B-swizzled tile
constexpr int BT_SMEM = 2 * BT_TILE /*K, V*/ + 2 * 2 * BT_TILE /*Q, dO x 2 stages*/ + 2 * BT_TILE /*P^T*/ + 2 * BT_TILE /*dS^T*/
                        + 2 * 2 * 128 * 4 /*lse, D x 2 stages*/ + 128 /*barriers*/;

__device__ __forceinline__ uint64_t bt_mnmajor_desc(uint32_t smem_addr) { return make_sw128_mnmajor_desc_lbo(smem_addr, 1024); }

__global__ void __launch_bounds__(BT_THREADS, 1)
attn_bwd_tc_kernel(AttnBwdArgs a) {
    extern __shared__ __align__(1024) uint8_t bt_raw[];
    uint8_t* smem = bt_raw;
    const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * 128;
    const int nq = a.Le + a.T;
    const int nk = a.n_keys[b];
    const int nkv = nk + a.T;
    if (j0 >= nkv) return;                          // uniform per CTA, before any barrier / TMEM set-up
    if (smem_u32(smem) & 1023u) __trap();
    uint8_t* sK = smem;
    uint8_t* sV = sK + BT_TILE;
    uint8_t* sQ = sV + BT_TILE;                     // [2 stages]
    uint8_t* sG = sQ + 2 * BT_TILE;                 // [2 stages] dO
    uint8_t* sPT = sG + 2 * BT_TILE;                // two tiles: queries 0-63 | 64-127
    uint8_t* sDS = sPT + 2 * BT_TILE;
    float* sStat = reinterpret_cast<float*>(sDS + 2 * BT_TILE);      // [stage][lse | D][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 2 * 2 * 128);
    uint64_t* kv_ready = bars;          // count 128
    uint64_t* q_full = bars + 1;        // [2] count 128
    uint64_t* q_empty = bars + 3;       // [2] count 1 (commit behind MMA 2)
    uint64_t* s_full = bars + 5;        // count 1 (commit behind MMA 1)
    uint64_t* p_full = bars + 6;        // count 32 * BT_SW (softmax threads)
    uint64_t* dq_full = bars + 7;       // count 1 (commit behind MMA 2)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int* kidx = a.key_idx + (long long)b * a.key_stride;
    const int col = h * XDH;
    // a key tile made only of decoder keys is seen by decoder queries only
    const int i_first = (j0 >= nk) ? a.Le / 128 : 0;
    const int n_tiles = (nq + 127) / 128 - i_first;

    if (warp == BT_SW + 4) {
        if (lane == 0) {
            mbar_init(kv_ready, 128);
            mbar_init(&q_full[0], 128); mbar_init(&q_full[1], 128);
            mbar_init(&q_empty[0], 1); mbar_init(&q_empty[1], 1);
            mbar_init(s_full, 1); mbar_init(p_full, 32 * BT_SW); mbar_init(dq_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320, tDQ = tmem_base + 384;

    if (warp >= BT_SW && warp < BT_SW + 4) {
        // ------------------------------------------------------------------ loaders
        const int lt = threadIdx.x - 32 * BT_SW;     // 0..127
        const int c = lt & 7, r0 = lt >> 3;          // 16-byte chunk / first row; rows r0 + 16 i
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = r0 + 16 * i;
            const int j = j0 + r;
            const bool ok = j < nkv;
            bool is_dec = false;
            const long long row = ok ? xb_krow_index(a, kidx, nk, j, is_dec) : 0;
            const __nv_bfloat16* src = (is_dec ? a.qkv_dec + ((long long)b * a.T + row) * a.ld_dec
                                               : a.qkv_enc + ((long long)b * a.Le + row) * a.ld_enc) + col + c * 8;
            const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
            cp_async16(sK + off, src + a.H, ok);
            cp_async16(sV + off, src + 2 * a.H, ok);
        }
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async();
        mbar_arrive(kv_ready);
        for (int t = 0; t < n_tiles; ++t) {
            const int s = t & 1;
            const int i0 = (i_first + t) * 128;
            mbar_wait(&q_empty[s], ((t >> 1) & 1) ^ 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + 16 * i;
                const int qi = i0 + r;
                const bool ok = qi < nq;
                const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
                cp_async16(sQ + s * BT_TILE + off,
                           xb_qrow(a, a.qkv_enc, a.ld_enc, a.qkv_dec, a.ld_dec, b, ok ? qi : 0) + col + c * 8, ok);
                cp_async16(sG + s * BT_TILE + off,
                           xb_qrow(a, a.do_enc, a.ldg_enc, a.do_dec, a.ldg_dec, b, ok ? qi : 0) + col + c * 8, ok);
            }
            cp_async_commit();
            {
                const int qi = i0 + lt;
                const float* st = a.stats + (((long long)b * a.heads + h) * nq + (qi < nq ? qi : 0)) * 2;
                // rows past the sequence: lse = +inf makes every probability of the column exp2(-inf) = 0
                sStat[(s * 2 + 0) * 128 + lt] = qi < nq ? st[0] : INFINITY;
                sStat[(s * 2 + 1) * 128 + lt] = qi < nq ? st[1] : 0.f;
            }
            cp_async_wait<0>();
            fence_proxy_async();                     // cp.async tiles -> visible to the tensor core
            mbar_arrive(&q_full[s]);                 // (release: the plain stores of the statistics, for the softmax warps)
        }
    } else if (warp == BT_SW + 4) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);                               // A, B K-major
        constexpr uint32_t idesc_kv = make_idesc_bf16(128, 64) | (1u << 16);                  // B MN-major
        constexpr uint32_t idesc_dq = make_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);     // A and B MN-major
        const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aPT = smem_u32(sPT), aDS = smem_u32(sDS);
        mbar_wait(kv_ready, 0);
        tc_fence_after();
        for (int t = 0; t < n_tiles; ++t) {
            const int s = t & 1;
            const uint32_t aQ = smem_u32(sQ + s * BT_TILE), aG = smem_u32(sG + s * BT_TILE);
            mbar_wait(&q_full[s], (t >> 1) & 1);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tS, make_sw128_kmajor_desc(aK) + 2 * k, make_sw128_kmajor_desc(aQ) + 2 * k, idesc_s, k ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tDP, make_sw128_kmajor_desc(aV) + 2 * k, make_sw128_kmajor_desc(aG) + 2 * k, idesc_s, k ? 1u : 0u);
                umma_commit(s_full);
            }
            __syncwarp();
            mbar_wait(p_full, t & 1);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {        // 16 queries per step: A = P^T / dS^T (K-major), B = dO / Q (MN-major)
                    const uint32_t a_off = (k >> 2) * BT_TILE;
                    umma_bf16(tDV, make_sw128_kmajor_desc(aPT + a_off) + 2 * (k & 3), bt_mnmajor_desc(aG + k * 2048), idesc_kv,
                              (t | k) ? 1u : 0u);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t a_off = (k >> 2) * BT_TILE;
                    umma_bf16(tDK, make_sw128_kmajor_desc(aDS + a_off) + 2 * (k & 3), bt_mnmajor_desc(aQ + k * 2048), idesc_kv,
                              (t | k) ? 1u : 0u);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k)          // 16 keys per step: A = dS (dS^T tiles read MN-major: M = queries)
                    umma_bf16(tDQ, make_sw128_mnmajor_desc_lbo(aDS + k * 2048, BT_TILE), bt_mnmajor_desc(aK + k * 2048), idesc_dq,
                              k ? 1u : 0u);
                umma_commit(&q_empty[s]);
                umma_commit(dq_full);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ softmax / read-out
        const int quarter = warp & 3, cq = warp >> 2;
        const int r = quarter * 32 + lane;           // key row of the tile (S^T, dP^T, dV, dK) / query row (dQ) = TMEM lane
        const int key = j0 + r;
        const bool key_ok = key < nkv;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const int sw = r & 7;
        const uint32_t drop_y = drop_attn_y(a.drop, b * a.heads + h);
        const float row_bias = key_ok ? 0.f : -INFINITY;               // rows past the key list: every probability 0
        // decoder keys are causal: key nk + j is seen by queries Le + j, Le + j + 1, ...; encoder keys by every query
        const bool has_dec = j0 + 127 >= nk;                            // tile-uniform
        const int q_min = key >= nk ? a.Le + (key - nk) : 0;
        for (int t = 0; t < n_tiles; ++t) {
            const int s = t & 1;
            const int i0 = (i_first + t) * 128;
            const float* lse_s = sStat + (s * 2 + 0) * 128 + cq * 32;
            const float* d_s = sStat + (s * 2 + 1) * 128 + cq * 32;
            mbar_wait(&q_full[s], (t >> 1) & 1);     // acquire the loaders' {lse, D} stores of this stage
            mbar_wait(s_full, t & 1);                // S^T / dP^T of this tile (and every earlier MMA) done
            tc_fence_after();
#pragma unroll 1
            for (int hc = 0; hc < 2; ++hc) {         // two 16-column halves (register budget: 96 per thread)
            uint32_t vs[16], vd[16];
            tmem_ld_32x16(tS + lane_addr + cq * 32 + hc * 16, vs);
            tmem_ld_32x16(tDP + lane_addr + cq * 32 + hc * 16, vd);
            tmem_ld_wait_on(vs);
            tmem_ld_wait_on(vd);
            uint32_t pp[8], dd[8];
#pragma unroll
            for (int j2 = 0; j2 < 16; j2 += 2) {
                const int jj = hc * 16 + j2;
                const float2 l2 = *reinterpret_cast<const float2*>(lse_s + jj);
                const float2 d2 = *reinterpret_cast<const float2*>(d_s + jj);
                float p0, p1;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(vs[j2]), a.scale_log2, row_bias) - l2.x));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(vs[j2 + 1]), a.scale_log2, row_bias) - l2.y));
                if (has_dec) {
                    const int qi = i0 + cq * 32 + jj;
                    if (qi < q_min) p0 = 0.f;
                    if (qi + 1 < q_min) p1 = 0.f;
                }
                float m0 = 1.f, m1 = 1.f;
                if (a.drop.thr) {
                    // one hash serves the key pair (key, key ^ 1) = this lane and lane ^ 1: the even lane hashes query column
                    // jj, the odd lane column jj + 1, and they swap
                    const uint32_t mine = drop_hash(a.drop.s0, a.drop.s1, drop_attn_x(i0 + cq * 32 + jj + (r & 1), key), drop_y);
                    const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
                    const uint32_t h0 = (r & 1) ? other : mine, h1 = (r & 1) ? mine : other;
                    const uint32_t u0 = (r & 1) ? (h0 >> 16) : (h0 & 0xffffu), u1 = (r & 1) ? (h1 >> 16) : (h1 & 0xffffu);
                    m0 = u0 >= a.drop.thr ? a.drop.scale : 0.f;
                    m1 = u1 >= a.drop.thr ? a.drop.scale : 0.f;
                }
                pp[j2 >> 1] = pack_bf16x2(p0 * m0, p1 * m1);
                dd[j2 >> 1] = pack_bf16x2(p0 * (__uint_as_float(vd[j2]) * m0 - d2.x) * a.scale,
                                          p1 * (__uint_as_float(vd[j2 + 1]) * m1 - d2.y) * a.scale);
            }
            {
                // 16 query columns = 32 bytes of row r: tile cq >> 1, 16-byte chunks (cq & 1) * 4 + hc * 2 + 0..1
                uint8_t* pdst = sPT + (cq >> 1) * BT_TILE + r * 128;
                uint8_t* ddst = sDS + (cq >> 1) * BT_TILE + r * 128;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int chunk = (cq & 1) * 4 + hc * 2 + q;
                    *reinterpret_cast<uint4*>(pdst + ((chunk ^ sw) << 4)) = make_uint4(pp[4 * q], pp[4 * q + 1], pp[4 * q + 2], pp[4 * q + 3]);
                    *reinterpret_cast<uint4*>(ddst + ((chunk ^ sw) << 4)) = make_uint4(dd[4 * q], dd[4 * q + 1], dd[4 * q + 2], dd[4 * q + 3]);
                }
            }
            }
            tc_fence_before();                       // TMEM reads ordered before the issuer's next MMAs
            fence_proxy_async();                     // P^T / dS^T tiles visible to the tensor core
            mbar_arrive(p_full);
            // ---- dQ of this tile pair: row r = query i0 + r, columns cq * 16 .. + 15
            mbar_wait(dq_full, t & 1);
            tc_fence_after();
            {
                const int qi = i0 + r;
                uint32_t v[16];
                tmem_ld_32x16(tDQ + lane_addr + cq * 16, v);           // warp-collective: rows past nq load too
                tmem_ld_wait_on(v);
                if (qi < nq) {
                    float* op = a.dq32 + ((long long)b * nq + qi) * a.H + col + cq * 16;
#pragma unroll
                    for (int d = 0; d < 16; d += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op + d),
                                     "f"(__uint_as_float(v[d])), "f"(__uint_as_float(v[d + 1])),
                                     "f"(__uint_as_float(v[d + 2])), "f"(__uint_as_float(v[d + 3])) : "memory");
                }
            }
            tc_fence_before();
        }
        // ---- dV, dK of this key tile (complete behind the last dq_full): columns cq * 16 .. + 15 of row r.  tcgen05.ld is
        // warp-collective: every lane loads, only the rows inside the key list store
        {
            bool is_dec = false;
            const long long row = key_ok ? xb_krow_index(a, kidx, nk, key, is_dec) : 0;
            __nv_bfloat16* op = (is_dec ? a.dqkv_dec + ((long long)b * a.T + row) * a.ldq_dec
                                        : a.dqkv_enc + ((long long)b * a.Le + row) * a.ldq_enc) + col + cq * 16;
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                uint32_t v[16];
                tmem_ld_32x16((which ? tDV : tDK) + lane_addr + cq * 16, v);
                tmem_ld_wait_on(v);
                if (key_ok) {
#pragma unroll
                    for (int d = 0; d < 16; d += 8) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(v[d]), __uint_as_float(v[d + 1]));
                        o.y = pack_bf16x2(__uint_as_float(v[d + 2]), __uint_as_float(v[d + 3]));
                        o.z = pack_bf16x2(__uint_as_float(v[d + 4]), __uint_as_float(v[d + 5]));
                        o.w = pack_bf16x2(__uint_as_float(v[d + 6]), __uint_as_float(v[d + 7]));
                        *reinterpret_cast<uint4*>(op + (which ? 2 : 1) * a.H + d) = o;
                    }
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == BT_SW + 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------- decoder rows, forward (nq > 1)
// The teacher-forced passes score several decoder rows of a sample at once (all T of them in the `ref` / `neg` tail of
// the eval forward and in every pass of the training step).  The scalar kernel of csrc/attention.cu reads K / V once
// per chunk of four queries and is latency bound (ncu: 97 us per launch, 0.10 of the HBM rate, on the critical path
// behind the greedy decode); this one is a small flash attention on mma.sync built from the tile helpers above: one
// CTA per (head, sample), the <= 16 query rows are ONE m16 tile, 64-key tiles of the virtual key list (compacted
// encoder keys, then the causal decoder keys) are double buffered with cp.async, each of the four warps takes 16 keys
// of a tile (S = Q K^T, online softmax, O += P V with P as bf16 A fragments), and the four partial (max, sum, O) sets
// meet in shared memory.  K and V are read exactly once.  `drop` / `lse_out`: the training form (see attn_bwd above).
struct AttnDecArgs {
    const __nv_bfloat16* qkv_enc; long long ld_enc; int L_enc;
    const __nv_bfloat16* qkv_dec; long long ld_dec; int T, H;
    const int* key_idx; const int* n_keys; int key_stride;
    int t0, nq;
    __nv_bfloat16* out; long long ldo;
    float scale_log2;
    DropCfg drop;
    float* lse_out;
};

__global__ void __launch_bounds__(XB_THREADS)
attn_dec_mma_kernel(AttnDecArgs a) {
    __shared__ __align__(128) uint8_t Qs[16 * 128];
    __shared__ __align__(128) uint8_t KVs[2][2][XB * 128];        // [stage][K | V]; reused for the final reduction
    const int h = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int heads = gridDim.x;
    const int nk = a.n_keys[b];
    const int nkv = nk + a.t0 + a.nq;                              // the last query sees decoder keys 0 .. t0 + nq - 1
    const int ntiles = (nkv + XB - 1) / XB;
    const int* kidx = a.key_idx + (long long)b * a.key_stride;
    const int col = h * XDH;
    const __nv_bfloat16* enc = a.qkv_enc + (long long)b * a.L_enc * a.ld_enc;
    const __nv_bfloat16* dec = a.qkv_dec + (long long)b * a.T * a.ld_dec;

    auto load_kv = [&](int t, int buf) {
        for (int i = threadIdx.x; i < XB * 8; i += XB_THREADS) {
            const int r = i >> 3, c = i & 7;
            const int j = t * XB + r;
            const bool ok = j < nkv;
            const __nv_bfloat16* src = (!ok ? enc : (j < nk ? enc + (long long)kidx[j] * a.ld_enc
                                                            : dec + (long long)(j - nk) * a.ld_dec)) + col + c * 8;
            cp_async16(KVs[buf][0] + xb_off(r, c), src + a.H, ok);
            cp_async16(KVs[buf][1] + xb_off(r, c), src + 2 * a.H, ok);
        }
    };
    for (int i = threadIdx.x; i < 16 * 8; i += XB_THREADS) {
        const int r = i >> 3, c = i & 7;
        const bool ok = r < a.nq;
        cp_async16(Qs + xb_off(r, c), dec + (long long)(a.t0 + (ok ? r : 0)) * a.ld_dec + col + c * 8, ok);
    }
    load_kv(0, 0);
    cp_async_commit();

    float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
    float acc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[n][j] = 0.f;
    uint32_t qf[4][4];
    const uint32_t drop_y = drop_attn_y(a.drop, b * heads + h);
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) load_kv(t + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (t == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)       // the 16 query rows as A fragments (all four warps hold the same ones)
                xb_ldmatrix_x4(qf[ks], smem_u32(Qs + xb_off(lane & 15, ks * 2 + (lane >> 4))));
        }
        if (t * XB + warp * 16 < nkv) {          // warp-uniform: this warp's 16 keys of the tile exist
            const uint8_t* Kt = KVs[buf][0];
            const uint8_t* Vt = KVs[buf][1];
            float c[2][4];
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int j = 0; j < 4; ++j) c[n][j] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t bf[4];
                xb_ldmatrix_x4(bf, smem_u32(Kt + xb_off(warp * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1))));
                xb_mma(c[0], qf[ks], bf[0], bf[1]);
                xb_mma(c[1], qf[ks], bf[2], bf[3]);
            }
            float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int key = t * XB + warp * 16 + n * 8 + tq * 2 + (j & 1);
                    const int qpos = a.t0 + g + (j >> 1) * 8;            // decoder position of this query row
                    float v = c[n][j] * a.scale_log2;
                    if (key >= nkv || (key >= nk && key - nk > qpos)) v = -INFINITY;
                    c[n][j] = v;
                    mx[j >> 1] = fmaxf(mx[j >> 1], v);
                }
            float corr[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
                mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
                const float m_new = fmaxf(m_i[r], mx[r]);
                corr[r] = m_new == -INFINITY ? 1.f : exp2f(m_i[r] - m_new);
                float rs = 0.f;
#pragma unroll
                for (int n = 0; n < 2; ++n)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float p = m_new == -INFINITY ? 0.f : exp2f(c[n][r * 2 + e] - m_new);
                        c[n][r * 2 + e] = p;
                        rs += p;
                    }
                l_i[r] = l_i[r] * corr[r] + rs;        // this lane's share of the row sum (the quad is summed at the end)
                m_i[r] = m_new;
            }
            if (a.drop.thr) {                           // training: mask the probabilities that multiply V (not the row sum)
#pragma unroll
                for (int n = 0; n < 2; ++n)
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int key = t * XB + warp * 16 + n * 8 + tq * 2;
                        const uint32_t hsh = drop_hash(a.drop.s0, a.drop.s1, drop_attn_x(a.L_enc + a.t0 + g + r * 8, key), drop_y);
                        if ((hsh & 0xffffu) < a.drop.thr) c[n][r * 2] = 0.f;
                        if ((hsh >> 16) < a.drop.thr) c[n][r * 2 + 1] = 0.f;
                    }
            }
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                acc[n][0] *= corr[0]; acc[n][1] *= corr[0];
                acc[n][2] *= corr[1]; acc[n][3] *= corr[1];
            }
            uint32_t pf[4];
            pf[0] = pack_bf16x2(c[0][0], c[0][1]);
            pf[1] = pack_bf16x2(c[0][2], c[0][3]);
            pf[2] = pack_bf16x2(c[1][0], c[1][1]);
            pf[3] = pack_bf16x2(c[1][2], c[1][3]);
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                uint32_t bf[4];
                xb_ldmatrix_x4_trans(bf, smem_u32(Vt + xb_off(warp * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), dp * 2 + (lane >> 4))));
                xb_mma(acc[dp * 2], pf, bf[0], bf[1]);
                xb_mma(acc[dp * 2 + 1], pf, bf[2], bf[3]);
            }
        }
        __syncthreads();
    }
    // ---- combine the four warps' partial results through shared memory (the K / V stages are free now)
    float* red = reinterpret_cast<float*>(&KVs[0][0][0]);       // [4 warps][16 rows][64 + 2]: O, max, sum
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
        l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
        float* row = red + (warp * 16 + g + r * 8) * 66;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            row[n * 8 + tq * 2] = acc[n][r * 2];
            row[n * 8 + tq * 2 + 1] = acc[n][r * 2 + 1];
        }
        if (tq == 0) { row[64] = m_i[r]; row[65] = l_i[r]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 16 * 32; i += XB_THREADS) {
        const int r = i >> 5, d = (i & 31) * 2;
        if (r >= a.nq) continue;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 4; ++w) M = fmaxf(M, red[(w * 16 + r) * 66 + 64]);
        float Lsum = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float* row = red + (w * 16 + r) * 66;
            const float f = row[64] == -INFINITY ? 0.f : exp2f(row[64] - M);
            Lsum += row[65] * f;
            o0 += row[d] * f;
            o1 += row[d + 1] * f;
        }
        const float inv = Lsum > 0.f ? a.drop.scale / Lsum : 0.f;
        *reinterpret_cast<uint32_t*>(a.out + ((long long)b * a.T + a.t0 + r) * a.ldo + col + d) = pack_bf16x2(o0 * inv, o1 * inv);
        if (a.lse_out && d == 0)
            a.lse_out[(((long long)b * heads + h) * (a.L_enc + a.T) + a.L_enc + a.t0 + r) * 2] = M + log2f(Lsum);
    }
}

// declared in csrc/attention.cu (t2s_attn_dec dispatches here when nq > 1)
int launch_attn_dec_mma(const void* qkv_enc, long long ld_enc, int L_enc, const void* qkv_dec, long long ld_dec, int T,
                        int B, int H, int heads, const int* key_idx, const int* n_keys, int key_stride, int t0, int nq,
                        void* out, long long ldo, cudaStream_t st, DropCfg drop, float* lse_out) {
    AttnDecArgs a;
    a.qkv_enc = reinterpret_cast<const __nv_bfloat16*>(qkv_enc); a.ld_enc = ld_enc; a.L_enc = L_enc;
    a.qkv_dec = reinterpret_cast<const __nv_bfloat16*>(qkv_dec); a.ld_dec = ld_dec; a.T = T; a.H = H;
    a.key_idx = key_idx; a.n_keys = n_keys; a.key_stride = key_stride;
    a.t0 = t0; a.nq = nq;
    a.out = reinterpret_cast<__nv_bfloat16*>(out); a.ldo = ldo;
    a.scale_log2 = 0.125f * 1.4426950408889634f;
    a.drop = drop;
    a.lse_out = lse_out;
    attn_dec_mma_kernel<<<dim3(heads, B), XB_THREADS, 0, st>>>(a);
    return launch_status("attn_dec");
}

// fused form: fp32 dQ accumulator -> the q columns of the bf16 dq|dk|dv rows
__global__ void __launch_bounds__(256)
attn_bwd_dq_store_kernel(const float* __restrict__ dq32, int B, int Le, int T, int H, __nv_bfloat16* __restrict__ enc,
                         long long ld_e, __nv_bfloat16* __restrict__ dec, long long ld_d) {
    const int nq = Le + T, h4 = H / 4;
    const long long n = (long long)B * nq * h4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / h4;
        const int c = (int)(i % h4) * 4, b = (int)(row / nq), qi = (int)(row % nq);
        const float4 v = *reinterpret_cast<const float4*>(dq32 + row * H + c);
        __nv_bfloat16* o = qi < Le ? enc + ((long long)b * Le + qi) * ld_e + c : dec + ((long long)b * T + (qi - Le)) * ld_d + c;
        uint2 w;
        w.x = pack_bf16x2(v.x, v.y);
        w.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(o) = w;
    }
}

}  // namespace t2s

using namespace t2s;

static inline long long attn_bwd_stats_bytes(int B, int Le, int T, int heads) {
    return (((long long)B * heads * (Le + T) * 2 * sizeof(float)) + 255) & ~255LL;
}

extern "C" long long t2s_attn_bwd_workspace_bytes(int B, int Le, int T, int heads) {
    // {lse2, D} per (sample, head, query) + the fp32 dQ accumulator of the fused dK / dV / dQ pass (head size 64)
    return attn_bwd_stats_bytes(B, Le, T, heads) + (long long)B * (Le + T) * heads * XDH * sizeof(float);
}

static int attn_bwd_entry(const void* qkv_enc, long long ld_enc, const void* qkv_dec, long long ld_dec,
                          const void* o_enc, long long ldo_enc, const void* o_dec, long long ldo_dec,
                          const void* do_enc, long long ldg_enc, const void* do_dec, long long ldg_dec,
                          void* dqkv_enc, long long ldq_enc, void* dqkv_dec, long long ldq_dec, int B, int Le, int T,
                          int H, int heads, const int* key_idx, const int* n_keys, int key_stride, int max_keys,
                          void* workspace, void* stream, DropCfg drop, float* stats_lse = nullptr) {
    if (H != heads * XDH || B <= 0 || Le <= 0 || T < 0 || max_keys <= 0) {
        set_error("attn_bwd: head size must be 64 (H %d heads %d B %d Le %d T %d)", H, heads, B, Le, T);
        return T2S_ERR_SHAPE;
    }
    if ((ld_enc % 8) || (ldo_enc % 8) || (ldg_enc % 8) || (ldq_enc % 8) ||
        (T > 0 && ((ld_dec % 8) || (ldo_dec % 8) || (ldg_dec % 8) || (ldq_dec % 8) || !qkv_dec || !o_dec || !do_dec || !dqkv_dec))) {
        set_error("attn_bwd: row pitches must be multiples of 8 and the decoder buffers present when T > 0");
        return T2S_ERR_ALIGN;
    }
    typedef __nv_bfloat16 bf;
    AttnBwdArgs a;
    a.qkv_enc = reinterpret_cast<const bf*>(qkv_enc); a.ld_enc = ld_enc;
    a.qkv_dec = reinterpret_cast<const bf*>(qkv_dec); a.ld_dec = ld_dec;
    a.o_enc = reinterpret_cast<const bf*>(o_enc); a.ldo_enc = ldo_enc;
    a.o_dec = reinterpret_cast<const bf*>(o_dec); a.ldo_dec = ldo_dec;
    a.do_enc = reinterpret_cast<const bf*>(do_enc); a.ldg_enc = ldg_enc;
    a.do_dec = reinterpret_cast<const bf*>(do_dec); a.ldg_dec = ldg_dec;
    a.dqkv_enc = reinterpret_cast<bf*>(dqkv_enc); a.ldq_enc = ldq_enc;
    a.dqkv_dec = reinterpret_cast<bf*>(dqkv_dec); a.ldq_dec = ldq_dec;
    a.key_idx = key_idx; a.n_keys = n_keys; a.key_stride = key_stride;
    a.stats = stats_lse ? stats_lse : reinterpret_cast<float*>(workspace);
    a.have_lse = stats_lse ? 1 : 0;
    static const int fused = []() { const char* e = getenv("T2S_ATTN_BWD_FUSED"); return (e && e[0] == '0') ? 0 : 1; }();
    a.dq32 = fused ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + attn_bwd_stats_bytes(B, Le, T, heads)) : nullptr;
    a.Le = Le; a.T = T; a.H = H; a.heads = heads;
    a.scale = 0.125f;
    a.scale_log2 = 0.125f * 1.4426950408889634f;
    a.drop = drop;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // rows outside the key list receive no dK / dV: clear the k | v columns (q columns are fully written by dq)
    cudaError_t e = cudaMemset2DAsync(reinterpret_cast<bf*>(dqkv_enc) + H, ldq_enc * sizeof(bf), 0, 2 * (size_t)H * sizeof(bf),
                                      (size_t)B * Le, st);
    if (e != cudaSuccess) { set_error("attn_bwd memset: %s", cudaGetErrorString(e)); return (int)e; }
    static bool attr = false;
    if (!attr) {
        e = cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XB_DKV_SMEM);
        if (e != cudaSuccess) { set_error("attn_bwd attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    const int nq = Le + T;
    dim3 gq((nq + XB - 1) / XB, heads, B);
    attn_bwd_stats_kernel<<<gq, XB_THREADS, 0, st>>>(a);
    dim3 gk((max_keys + T + XB - 1) / XB, heads, B);
    if (a.dq32) {
        e = cudaMemsetAsync(a.dq32, 0, (size_t)B * nq * H * sizeof(float), st);
        if (e != cudaSuccess) { set_error("attn_bwd memset dq: %s", cudaGetErrorString(e)); return (int)e; }
        // T2S_ATTN_BWD_TC=0: the mma.sync form of the fused pass
        static const int use_tc = []() { const char* ev = getenv("T2S_ATTN_BWD_TC"); return (ev && ev[0] == '0') ? 0 : 1; }();
        if (use_tc) {
            static bool tc_attr = false;
            if (!tc_attr) {
                e = cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM);
                if (e != cudaSuccess) { set_error("attn_bwd_tc attr: %s", cudaGetErrorString(e)); return (int)e; }
                tc_attr = true;
            }
            dim3 gt((max_keys + T + 127) / 128, heads, B);
            attn_bwd_tc_kernel<<<gt, BT_THREADS, BT_SMEM, st>>>(a);
        } else
        attn_bwd_dkv_kernel<<<gk, XB_THREADS, XB_DKV_SMEM, st>>>(a);
        const long long n4 = (long long)B * nq * (H / 4);
        int grid = (int)((n4 + 255) / 256);
        if (grid > num_sms() * 16) grid = num_sms() * 16;
        attn_bwd_dq_store_kernel<<<grid, 256, 0, st>>>(a.dq32, B, Le, T, H, a.dqkv_enc, a.ldq_enc, a.dqkv_dec, a.ldq_dec);
    } else {
        attn_bwd_dq_kernel<<<gq, XB_THREADS, 0, st>>>(a);
        attn_bwd_dkv_kernel<<<gk, XB_THREADS, XB_DKV_SMEM, st>>>(a);
    }
    return launch_status("attn_bwd");
}

extern "C" int t2s_attn_bwd(const void* qkv_enc, long long ld_enc, const void* qkv_dec, long long ld_dec,
                            const void* o_enc, long long ldo_enc, const void* o_dec, long long ldo_dec,
                            const void* do_enc, long long ldg_enc, const void* do_dec, long long ldg_dec,
                            void* dqkv_enc, long long ldq_enc, void* dqkv_dec, long long ldq_dec, int B, int Le, int T,
                            int H, int heads, const int* key_idx, const int* n_keys, int key_stride, int max_keys,
                            void* workspace, void* stream) {
    return attn_bwd_entry(qkv_enc, ld_enc, qkv_dec, ld_dec, o_enc, ldo_enc, o_dec, ldo_dec, do_enc, ldg_enc, do_dec, ldg_dec,
                          dqkv_enc, ldq_enc, dqkv_dec, ldq_dec, B, Le, T, H, heads, key_idx, n_keys, key_stride, max_keys,
                          workspace, stream, DropCfg{0, 0, 0, 0, 1.f});
}

/* backward of t2s_attn_tc_dropout (encoder rows) + t2s_attn_dec_dropout (decoder rows) of one layer: same (p, seed,
 * site), p may be 0.  stats_lse != null: the [B, heads, Le + T, 2] fp32 buffer those kernels wrote the rows' log2-sum-exp
 * into (slot 0; slot 1 is filled here with D = dO . O) -- the S recomputation of the statistics pass is skipped */
extern "C" int t2s_attn_bwd_dropout(const void* qkv_enc, long long ld_enc, const void* qkv_dec, long long ld_dec,
                                    const void* o_enc, long long ldo_enc, const void* o_dec, long long ldo_dec,
                                    const void* do_enc, long long ldg_enc, const void* do_dec, long long ldg_dec,
                                    void* dqkv_enc, long long ldq_enc, void* dqkv_dec, long long ldq_dec, int B, int Le,
                                    int T, int H, int heads, const int* key_idx, const int* n_keys, int key_stride,
                                    int max_keys, void* workspace, float p, unsigned long long seed, unsigned site,
                                    float* stats_lse, void* stream) {
    if (p < 0.f || p >= 1.f || Le + T > 65535) { set_error("attn_bwd_dropout: p in [0, 1), L < 65536"); return T2S_ERR_ARG; }
    return attn_bwd_entry(qkv_enc, ld_enc, qkv_dec, ld_dec, o_enc, ldo_enc, o_dec, ldo_dec, do_enc, ldg_enc, do_dec, ldg_dec,
                          dqkv_enc, ldq_enc, dqkv_dec, ldq_dec, B, Le, T, H, heads, key_idx, n_keys, key_stride, max_keys,
                          workspace, stream, make_drop(p, seed, site), stats_lse);
}
