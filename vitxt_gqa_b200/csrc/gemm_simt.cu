// K1f: fp32 GEMM on the FMA pipes.  C[M,N] = epi(A[M,K] . W[N,K]^T + bias), all fp32.
//
// Used for the grounding chain (TextBert -> obj/OCR encoders -> QTV -> q_linear), whose
// top-k frame / OCR indices must match the fp32 reference exactly (north_star "bit-exact
// grounded indices", SURVEY hard part 2): bf16 operand rounding moves the q.f logits by
// +-0.2 and flips top-k membership.  Same call sites as gemm_tcgen05.cu, fp32 operands.
//
// 128x128 output tile per CTA, BK=16, 256 threads each owning an 8x8 micro-tile split as
// 2x2 quads of 4x4 (rows {ty*4, 64+ty*4}, cols {tx*4, 64+tx*4}) so that shared-memory
// float4 reads are bank-conflict free; operands are stored k-major in shared memory and
// double buffered, global loads are 16-byte vectors.
#include "common.cuh"
#include "../../include/t2s_b200.h"

namespace t2s {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_THREADS = 256;

struct SimtEpi {
    float* C;
    const float* bias;
    const float* residual;
    long long ldc, ldr;
    int flags;
};

__global__ void __launch_bounds__(SG_THREADS, 2)
gemm_f32_simt_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W, long long ldw,
                     SimtEpi ep, int M, int N, int K, int a_per, int a_group, int a_off) {
    __shared__ __align__(16) float As[2][SG_BK][SG_BM + 4];
    __shared__ __align__(16) float Bs[2][SG_BK][SG_BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;

    // global->smem mapping: 128 rows x 16 k = 512 float4; thread loads rows (tid>>2) and (tid>>2)+64, k4 = tid&3
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lr + h * 64;
            const int gm = m0 + r, gn = n0 + r, gk = k0 + lk;
            // optional gather of A rows out of a [groups, a_group, K] buffer (e.g. the question rows of the joint buffer)
            const long long arow = a_per > 0 ? (long long)(gm / a_per) * a_group + a_off + gm % a_per : gm;
            ra[h] = (gm < M && gk < K) ? *reinterpret_cast<const float4*>(A + arow * lda + gk)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            rb[h] = (gn < N && gk < K) ? *reinterpret_cast<const float4*>(W + (long long)gn * ldw + gk)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lr + h * 64;
            As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y;
            As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
            Bs[buf][lk + 0][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y;
            Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = (K + SG_BK - 1) / SG_BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * SG_BK);
#pragma unroll
        for (int k = 0; k < SG_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) sstore(buf ^ 1);
        __syncthreads();
    }

    const bool gelu = ep.flags & T2S_GEMM_GELU;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (row >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int col = n0 + jh * 64 + tx * 4;
            if (col >= N) continue;
            float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (col + j < N) {
                    if (ep.bias) v[j] += __ldg(ep.bias + col + j);
                    if (gelu) v[j] = gelu_erf(v[j]);
                    if (ep.residual) v[j] += ep.residual[(long long)row * ep.ldr + col + j];
                }
            }
            float* cp = ep.C + (long long)row * ep.ldc + col;
            if (col + 4 <= N && (ep.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.C) & 15) == 0) {
                *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
                for (int j = 0; j < 4; ++j) if (col + j < N) cp[j] = v[j];
            }
        }
    }
}

}  // namespace t2s

using namespace t2s;

extern "C" int t2s_gemm_f32(const float* A, long long lda, const float* W, long long ldw, const float* bias,
                            const float* residual, long long ldr, float* C, long long ldc, int M, int N, int K,
                            int flags, int a_rows_per_group, int a_group_rows, int a_row_off, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0) { set_error("gemm_f32: bad shape %d %d %d", M, N, K); return T2S_ERR_SHAPE; }
    if ((K % 4) || (lda % 4) || (ldw % 4) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15)) {
        set_error("gemm_f32: K, lda, ldw must be multiples of 4 and A/W 16-byte aligned (K %d lda %lld ldw %lld)", K, lda, ldw);
        return T2S_ERR_ALIGN;
    }
    SimtEpi ep{C, bias, residual, ldc, ldr, flags};
    dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
    gemm_f32_simt_kernel<<<grid, SG_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(A, lda, W, ldw, ep, M, N, K,
                                                                                          a_rows_per_group, a_group_rows, a_row_off);
    return launch_status("gemm_f32_simt");
}
