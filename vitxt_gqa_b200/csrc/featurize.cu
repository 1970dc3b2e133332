// Input featurisation that sits immediately before the forward path: the PHOC descriptor of
// every OCR token, built on the device from the token bytes instead of by the reference's CPU extension
// (pythia/utils/phoc/src/cphoc.c:12-113 behind pythia/utils/phoc/build_phoc.py:9-14 and
// pythia/datasets/processors.py:904-928).  HBM-bound: reads a few bytes per token, writes 604 floats per token.
//
// One warp per token.  The warp streams the token's bytes 32 at a time, lower-cases ASCII, drops everything
// outside [a-z0-9] (the wrapper's filter; bytes >= 0x80 of multi-byte UTF-8 sequences can never match, so the
// byte filter equals the reference's character filter) and numbers the kept symbols with a ballot prefix.  Every
// kept symbol sets its unigram bits (levels 2..5) and, together with the previous kept symbol, its bigram bits in a
// 608-bit mask in shared memory; the mask is then expanded into coalesced float stores.  All occupancy arithmetic
// is IEEE binary32 with explicit round-to-nearest intrinsics in the order of the C source, so the result is
// bit-identical to the reference binary (tests/test_kernels_gpu.py against tests/golden/phoc_golden.npz).
#include "common.cuh"
#include "../../include/t2s_b200.h"

namespace t2s {

constexpr int kPhocDim = 604;
constexpr int kPhocWords = 19;   // 19 * 32 = 608 bits
constexpr int kPhocWarps = 8;

// bigram id (0..49) of the symbol pair (a, b), or -1 -- cphoc.c:30; filled once from the host table below
__constant__ signed char c_bigram[36 * 36];

static const char* const kBigrams[50] = {
    "th", "he", "in", "er", "an", "re", "es", "on", "st", "nt", "en", "at", "ed", "nd", "to", "or", "ea",
    "ti", "ar", "te", "ng", "al", "it", "as", "is", "ha", "et", "se", "ou", "of", "le", "sa", "ve", "ro",
    "ra", "ri", "hi", "ne", "me", "de", "co", "ta", "ec", "si", "ll", "so", "na", "li", "la", "el"};

__device__ __forceinline__ int phoc_symbol(unsigned c) {
    if (c - 'A' < 26u) c += 32;            // ASCII lower-case (build_phoc.py:10)
    if (c - 'a' < 26u) return (int)(c - 'a');
    if (c - '0' < 10u) return 26 + (int)(c - '0');
    return -1;                             // dropped by the alphabet filter (build_phoc.py:11)
}

// fraction of [a0, a1] that lies in region `region` of `level` >= 0.5 ?  (cphoc.c:58-63 / 97-102)
__device__ __forceinline__ bool phoc_hit(float a0, float a1, int region, int level) {
    const float r0 = __fdiv_rn((float)region, (float)level);
    const float r1 = __fdiv_rn((float)(region + 1), (float)level);
    const float o0 = fmaxf(a0, r0), o1 = fminf(a1, r1);
    return __fdiv_rn(__fsub_rn(o1, o0), __fsub_rn(a1, a0)) >= 0.5f;
}

__global__ void __launch_bounds__(kPhocWarps * 32)
phoc_build_kernel(const unsigned char* __restrict__ bytes, const int* __restrict__ offsets, int width, int n_tokens,
                  int rows, float* __restrict__ out, long long ldo) {
    __shared__ unsigned s_bits[kPhocWarps][kPhocWords + 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int row = blockIdx.x * kPhocWarps + wid;
    if (row >= rows) return;
    unsigned* bits = s_bits[wid];
    if (lane < kPhocWords + 1) bits[lane] = 0u;
    __syncwarp();
    if (row < n_tokens) {
        // offsets == null: fixed-width records (`width` bytes per token, zero padded: byte 0 is outside the alphabet)
        const int beg = offsets ? offsets[row] : row * width, end = offsets ? offsets[row + 1] : beg + width;
        // pass 1: length of the filtered word (cphoc.c:32 strlen)
        int n = 0;
        for (int p = beg; p < end; p += 32) {
            const int sym = (p + lane < end) ? phoc_symbol(bytes[p + lane]) : -1;
            n += __popc(__ballot_sync(0xffffffffu, sym >= 0));
        }
        // pass 2: unigram + bigram bits
        const float fn = (float)n;
        int base = 0, carry = -1;      // kept symbols before this chunk; last kept symbol before this chunk
        for (int p = beg; p < end; p += 32) {
            const int sym = (p + lane < end) ? phoc_symbol(bytes[p + lane]) : -1;
            const unsigned kept = __ballot_sync(0xffffffffu, sym >= 0);
            const unsigned below = kept & ((1u << lane) - 1u);
            const int src = below ? 31 - __clz(below) : 0;
            int prev = __shfl_sync(0xffffffffu, sym, src);
            if (!below) prev = carry;
            if (sym >= 0) {
                const int index = base + __popc(below);
                const float occ0 = __fdiv_rn((float)index, fn), occ1 = __fdiv_rn((float)(index + 1), fn);
                int level_off = 0;                                   // 36 * (2 + .. + (level-1)), cphoc.c:66-68
#pragma unroll
                for (int level = 2; level < 6; ++level) {
#pragma unroll
                    for (int region = 0; region < level; ++region)
                        if (phoc_hit(occ0, occ1, region, level)) {
                            const int f = level_off + region * 36 + sym;
                            atomicOr(&bits[f >> 5], 1u << (f & 31));
                        }
                    level_off += level * 36;
                }
                if (prev >= 0) {                                     // bigram (index-1, index), cphoc.c:77-107
                    const int k = c_bigram[prev * 36 + sym];
                    if (k >= 0) {
                        const float g0 = __fdiv_rn((float)(index - 1), fn), g1 = __fdiv_rn((float)(index + 1), fn);
#pragma unroll
                        for (int region = 0; region < 2; ++region)
                            if (phoc_hit(g0, g1, region, 2)) {
                                const int f = 504 + region * 50 + k;
                                atomicOr(&bits[f >> 5], 1u << (f & 31));
                            }
                    }
                }
            }
            if (kept) carry = __shfl_sync(0xffffffffu, sym, 31 - __clz(kept));
            base += __popc(kept);
        }
    }
    __syncwarp();
    float* o = out + (long long)row * ldo;
    if ((ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        for (int j = lane; j < kPhocDim / 4; j += 32) {              // 151 float4 per row
            const unsigned w = bits[(4 * j) >> 5] >> ((4 * j) & 31);
            float4 v;
            v.x = (float)(w & 1u); v.y = (float)((w >> 1) & 1u); v.z = (float)((w >> 2) & 1u); v.w = (float)((w >> 3) & 1u);
            reinterpret_cast<float4*>(o)[j] = v;
        }
    } else {
        for (int j = lane; j < kPhocDim; j += 32) o[j] = (float)((bits[j >> 5] >> (j & 31)) & 1u);
    }
}

static int upload_bigram_table() {
    signed char tab[36 * 36];
    for (int i = 0; i < 36 * 36; ++i) tab[i] = -1;
    auto sym = [](char c) { return c >= 'a' ? c - 'a' : 26 + (c - '0'); };
    for (int k = 49; k >= 0; --k) tab[sym(kBigrams[k][0]) * 36 + sym(kBigrams[k][1])] = (signed char)k;  // first match wins
    return (int)cudaMemcpyToSymbol(c_bigram, tab, sizeof(tab));
}

}  // namespace t2s

static int phoc_entry(const unsigned char* bytes, const int* offsets, int width, int n_tokens, int rows, float* out,
                      long long ldo, void* stream) {
    using namespace t2s;
    if (n_tokens < 0 || rows < n_tokens || ldo < kPhocDim || (!offsets && (width <= 0 || (long long)rows * width > 0x7fffffffLL))) {
        set_error("t2s_phoc_build: need 0 <= n_tokens <= rows and ldo >= 604 (got %d, %d, %lld)", n_tokens, rows, ldo);
        return T2S_ERR_SHAPE;
    }
    if (rows == 0) return T2S_OK;
    if (!out || (n_tokens > 0 && !bytes)) {
        set_error("t2s_phoc_build: null pointer");
        return T2S_ERR_ARG;
    }
    static thread_local int table_dev = -1;     // the constant table is per device (one host thread per GPU)
    int dev = 0;
    cudaGetDevice(&dev);
    if (table_dev != dev) {
        const int rc = upload_bigram_table();
        if (rc != 0) {
            set_error("t2s_phoc_build: constant upload failed: %s", cudaGetErrorString((cudaError_t)rc));
            return rc;
        }
        table_dev = dev;
    }
    const int grid = (rows + kPhocWarps - 1) / kPhocWarps;
    phoc_build_kernel<<<grid, kPhocWarps * 32, 0, (cudaStream_t)stream>>>(bytes, offsets, width, n_tokens, rows, out, ldo);
    return launch_status("t2s_phoc_build");
}

extern "C" int t2s_phoc_build(const unsigned char* bytes, const int* offsets, int n_tokens, int rows, float* out,
                              long long ldo, void* stream) {
    if (n_tokens > 0 && !offsets) { t2s::set_error("t2s_phoc_build: null offsets"); return T2S_ERR_ARG; }
    return phoc_entry(bytes, offsets, 0, n_tokens, rows, out, ldo, stream);
}

extern "C" int t2s_phoc_build_fixed(const unsigned char* bytes, int width, int n_tokens, float* out, long long ldo,
                                    void* stream) {
    return phoc_entry(bytes, nullptr, width, n_tokens, n_tokens, out, ldo, stream);
}
