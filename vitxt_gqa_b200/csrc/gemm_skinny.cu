// K1s: "skinny" bf16 GEMM for the greedy-decode rows (M <= 64 rows per tile: one decoder row per sample).
//   C[M,N] = epi(A[M,K] . W[N,K]^T + bias (+ residual))
//
// Same call sites as t2s_gemm_bf16 (BertSelfAttention q|k|v, BertSelfOutput / BertIntermediate / BertOutput dense,
// ClassifierLayer, OcrPtrNet.query: reference pythia/models/t2s.py:622,653 and modules/layers.py:101-107) for the
// 12 x 23 GEMMs of the decode chain, where M = batch (64) and the weight matrix is the only real traffic.  The
// persistent 128-row tcgen05 kernel puts such a problem on N/64 = 12..48 CTAs that each walk the whole K in
// sequence (10-14 us of pure latency per launch, 276 launches on the critical path of a step).  Here the problem
// is cut into (N/64) x splits CTAs of one [64 x 64] output tile over a K range of 64..256, so ~148 SMs each pull
// <= 64 KB with every cp.async in flight at once, multiply with mma.sync m16n8k16 (tensor-pipe time is irrelevant
// at M = 64) and, when K is split, meet through a workspace: partial tiles are written in fp32, the last CTA to
// arrive on the tile's counter sums them in split order -- deterministic -- and applies the epilogue.
#include "common.cuh"
#include "../../include/t2s_b200.h"

namespace t2s {

constexpr int SK_BM = 64, SK_BN = 64, SK_THREADS = 128;
constexpr int SK_MAXK = 256;                         // K range of one CTA
constexpr int SK_PITCH = SK_MAXK * 2 + 16;           // bytes per smem row: 16 B skew keeps ldmatrix conflict free
constexpr int SK_SMEM = 2 * SK_BM * SK_PITCH;        // A rows then W rows

struct SkinnyEpi {
    void* C;
    const float* bias;
    const void* residual;
    long long ldc, ldr;
    int flags;
};

__device__ __forceinline__ void sk_ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void sk_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(SK_THREADS)
gemm_skinny_kernel(const __nv_bfloat16* __restrict__ A, long long lda, const __nv_bfloat16* __restrict__ W, long long ldw,
                   SkinnyEpi ep, int M, int N, int K, int k_per_split, int splits, float* __restrict__ partial,
                   unsigned int* __restrict__ counters) {
    extern __shared__ __align__(128) uint8_t sk_smem[];
    uint8_t* As = sk_smem;
    uint8_t* Ws = sk_smem + SK_BM * SK_PITCH;
    __shared__ unsigned int s_ticket;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
    const int n0 = blockIdx.x * SK_BN, split = blockIdx.y, m0 = blockIdx.z * SK_BM;
    const int k0 = split * k_per_split;
    const int kn = min(K, k0 + k_per_split) - k0;          // multiple of 8 (host checks K % 8)
    const int chunks = kn >> 3;                            // 16-byte chunks per row

    // ---- stage the A rows and W rows of this K range (zero fill outside M / N / the range, to a multiple of 16)
    const int kpad = (kn + 15) & ~15;
    for (int i = tid; i < SK_BM * (kpad >> 3); i += SK_THREADS) {
        const int r = i / (kpad >> 3), c = i % (kpad >> 3);
        const bool in_k = c < chunks;
        const bool oka = in_k && (m0 + r) < M, okw = in_k && (n0 + r) < N;
        cp_async16(As + r * SK_PITCH + c * 16, A + (long long)(oka ? m0 + r : 0) * lda + k0 + (in_k ? c * 8 : 0), oka);
        cp_async16(Ws + r * SK_PITCH + c * 16, W + (long long)(okw ? n0 + r : 0) * ldw + k0 + (in_k ? c * 8 : 0), okw);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // ---- warp w: rows 16w..16w+15, all 64 columns
    float acc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[n][j] = 0.f;
    for (int ks = 0; ks < (kpad >> 4); ++ks) {
        uint32_t af[4];
        sk_ldmatrix_x4(af, smem_u32(As + (warp * 16 + (lane & 15)) * SK_PITCH + (ks * 2 + (lane >> 4)) * 16));
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t bf[4];
            const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
            sk_ldmatrix_x4(bf, smem_u32(Ws + row * SK_PITCH + (ks * 2 + ((lane >> 3) & 1)) * 16));
            sk_mma(acc[np * 2], af, bf[0], bf[1]);
            sk_mma(acc[np * 2 + 1], af, bf[2], bf[3]);
        }
    }

    // ---- split K: park the partial tile, the last CTA of the tile adds them up in split order
    if (splits > 1) {
        const int tile = blockIdx.z * gridDim.x + blockIdx.x;
        float* mine = partial + ((long long)tile * splits + split) * (SK_BM * SK_BN);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            *reinterpret_cast<float2*>(mine + (warp * 16 + g) * SK_BN + n * 8 + tq * 2) = make_float2(acc[n][0], acc[n][1]);
            *reinterpret_cast<float2*>(mine + (warp * 16 + g + 8) * SK_BN + n * 8 + tq * 2) = make_float2(acc[n][2], acc[n][3]);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) s_ticket = atomicAdd(counters + tile, 1u);
        __syncthreads();
        if (s_ticket != (unsigned)(splits - 1)) return;
        __threadfence();
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[n][j] = 0.f;
        const float* base = partial + (long long)tile * splits * (SK_BM * SK_BN);
        for (int s = 0; s < splits; ++s) {
            const float* p = base + (long long)s * (SK_BM * SK_BN);
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float2 lo = __ldcg(reinterpret_cast<const float2*>(p + (warp * 16 + g) * SK_BN + n * 8 + tq * 2));
                const float2 hi = __ldcg(reinterpret_cast<const float2*>(p + (warp * 16 + g + 8) * SK_BN + n * 8 + tq * 2));
                acc[n][0] += lo.x; acc[n][1] += lo.y; acc[n][2] += hi.x; acc[n][3] += hi.y;
            }
        }
        if (tid == 0) counters[tile] = 0;                   // ready for the next launch on this stream
    }

    // ---- epilogue: bias, erf-GELU, residual, bf16 or fp32 store
    const bool gelu = ep.flags & T2S_GEMM_GELU;
    const bool out_f32 = ep.flags & T2S_GEMM_OUT_F32;
    const bool res_f32 = ep.flags & T2S_GEMM_RES_F32;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = m0 + warp * 16 + g + r * 8;
        if (row >= M) continue;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int col = n0 + n * 8 + tq * 2;
            if (col >= N) continue;
            const bool two = col + 1 < N;
            float v0 = acc[n][r * 2], v1 = acc[n][r * 2 + 1];
            if (ep.bias) { v0 += __ldg(ep.bias + col); if (two) v1 += __ldg(ep.bias + col + 1); }
            if (gelu) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
            if (ep.residual) {
                if (res_f32) {
                    const float* rp = reinterpret_cast<const float*>(ep.residual) + (long long)row * ep.ldr + col;
                    v0 += rp[0]; if (two) v1 += rp[1];
                } else {
                    const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(ep.residual) + (long long)row * ep.ldr + col;
                    v0 += __bfloat162float(rp[0]); if (two) v1 += __bfloat162float(rp[1]);
                }
            }
            if (out_f32) {
                float* cp = reinterpret_cast<float*>(ep.C) + (long long)row * ep.ldc + col;
                cp[0] = v0; if (two) cp[1] = v1;
            } else {
                __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(ep.C) + (long long)row * ep.ldc + col;
                if (two) *reinterpret_cast<uint32_t*>(cp) = pack_bf16x2(v0, v1);
                else cp[0] = __float2bfloat16_rn(v0);
            }
        }
    }
}

static void skinny_plan(int M, int N, int K, int& n_tiles, int& m_tiles, int& splits, int& k_per_split) {
    n_tiles = (N + SK_BN - 1) / SK_BN;
    m_tiles = (M + SK_BM - 1) / SK_BM;
    const int min_splits = (K + SK_MAXK - 1) / SK_MAXK;
    splits = (num_sms() + n_tiles * m_tiles / 2) / (n_tiles * m_tiles);       // ~one CTA per SM
    if (splits < min_splits) splits = min_splits;
    const int max_splits = (K + 63) / 64;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    k_per_split = (((K + splits - 1) / splits) + 15) & ~15;
    if (k_per_split > SK_MAXK) k_per_split = SK_MAXK;
    splits = (K + k_per_split - 1) / k_per_split;
}

}  // namespace t2s

using namespace t2s;

extern "C" long long t2s_gemm_skinny_workspace_bytes(int M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    int nt, mt, sp, kps;
    skinny_plan(M, N, K, nt, mt, sp, kps);
    // counters (one per output tile, zero-initialised by the caller ONCE; the kernel re-zeroes them) + partial tiles
    return 1024 + 16 + (long long)nt * mt * sizeof(unsigned int) + (long long)nt * mt * sp * SK_BM * SK_BN * sizeof(float);
}

extern "C" int t2s_gemm_skinny_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                                    const void* residual, long long ldr, void* C, long long ldc, int M, int N, int K,
                                    int flags, void* workspace, long long workspace_bytes, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0) { set_error("gemm_skinny: bad shape %d %d %d", M, N, K); return T2S_ERR_SHAPE; }
    if ((K % 8) || (lda % 8) || (ldw % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15)) {
        set_error("gemm_skinny: K, lda, ldw must be multiples of 8 and A/W 16-byte aligned (K %d lda %lld ldw %lld)", K, lda, ldw);
        return T2S_ERR_ALIGN;
    }
    if (flags & (T2S_GEMM_OUT_SPLIT | T2S_GEMM_DGELU)) { set_error("gemm_skinny: OUT_SPLIT / DGELU are not supported"); return T2S_ERR_ARG; }
    if ((ldc % 2) || (reinterpret_cast<uintptr_t>(C) & 3)) { set_error("gemm_skinny: C alignment"); return T2S_ERR_ALIGN; }
    int nt, mt, sp, kps;
    skinny_plan(M, N, K, nt, mt, sp, kps);
    const long long need = t2s_gemm_skinny_workspace_bytes(M, N, K);
    if (sp > 1 && (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15))) {
        set_error("gemm_skinny: workspace of %lld bytes (16-byte aligned) required, got %lld", need, workspace_bytes);
        return T2S_ERR_ARG;
    }
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(gemm_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM);
        if (e != cudaSuccess) { set_error("gemm_skinny attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    unsigned int* counters = reinterpret_cast<unsigned int*>(workspace);
    float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 1024 +
                                              (((long long)nt * mt * sizeof(unsigned int) + 15) & ~15LL));
    SkinnyEpi ep{C, bias, residual, ldc, ldr, flags};
    gemm_skinny_kernel<<<dim3(nt, sp, mt), SK_THREADS, SK_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(W), ldw, ep, M, N, K, kps, sp,
        partial, counters);
    return launch_status("gemm_skinny");
}
