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

// ------------------------------------------------------------------------------- frame sampling + OCR pad / pack
// One CTA per (sampled frame slot i, sample b).  Replaces the python list work of the dataset
// (vtextgqa/dataset.py:103-158 frame ids + per-frame truncate / pad, :166-195 the middle-frame ids, :199-243 the padded
// id / mask vectors and the float64 box normalisation; sample_frames :371-381).  Detections of a video lie back to back
// in OCR-info frame order; frame_ptr is their CSR index over ALL videos of the batch.
struct PackOut {
    float* bbox;                 // [B, F * Of, 4]
    long long* track;            // [B, F * Of]
    long long* temporal;         // [B, F * Of]
    long long* ocr_mask;         // [B, F * Of]
    long long* frame_id;         // [B, F]
    long long* frame_mask;       // [B, F]
    long long* frame_num;        // [B]
    long long* mid_id;           // [B]
    long long* mid_idx;          // [B]
    unsigned char* tokens;       // [B, F * Of, width]
};

__global__ void __launch_bounds__(128)
pack_ocr_frames_kernel(const float* __restrict__ det_points, const long long* __restrict__ det_track,
                       const unsigned char* __restrict__ det_tokens, int width, const int* __restrict__ frame_ptr,
                       const int* __restrict__ info_base, const int* __restrict__ n_info,
                       const int* __restrict__ n_frames, const double* __restrict__ vid_w,
                       const double* __restrict__ vid_h, int F, int Of, PackOut out) {
    const int i = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int n = n_frames[b];
    const int count = n < F ? n : F;                                   // len(idxs)
    const int step = n / F;                                            // sample_frames: frames[i * step] when n > F
    const long long slot0 = ((long long)b * F + i) * Of;
    if (tid == 0) {
        const int fid = i < count ? (n <= F ? i + 1 : 1 + i * step) : 0;
        out.frame_id[(long long)b * F + i] = fid;
        out.frame_mask[(long long)b * F + i] = i < count ? 1 : 0;
        if (i == 0) {
            const int last = n <= F ? count : 1 + (count - 1) * step;  // the LAST sampled frame (dataset.py:180)
            out.frame_num[b] = count;
            out.mid_id[b] = last;
            out.mid_idx[b] = last >= F ? count / 2 + 1 : last;
        }
    }
    if (i >= count) {        // frame slots past the video: zero ids, zero boxes, empty token records
        for (int o = tid; o < Of; o += blockDim.x) {
            out.track[slot0 + o] = 0;
            out.temporal[slot0 + o] = 0;
            out.ocr_mask[slot0 + o] = 0;
            *reinterpret_cast<float4*>(out.bbox + (slot0 + o) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int e = tid; e < Of * width; e += blockDim.x) out.tokens[slot0 * width + e] = 0;
        return;
    }
    const int frame_idx = n <= F ? i + 1 : 1 + i * step;
    const int j = n_info[b] >= frame_idx ? frame_idx : frame_idx - 1;  // dataset.py:121-124
    const int row = info_base[b] + j - 1;
    const int lo = frame_ptr[row];
    const int k = min(frame_ptr[row + 1] - lo, Of);
    const double sx = 1.0 / vid_w[b], sy = 1.0 / vid_h[b];            // python floats: the product is binary64
    for (int o = tid; o < Of; o += blockDim.x) {
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        long long trk = 0;
        if (o < k) {
            const float* p = det_points + (long long)(lo + o) * 8;
            const float x1 = fminf(p[0], p[6]), y1 = fminf(p[1], p[3]), x2 = fmaxf(p[2], p[4]), y2 = fmaxf(p[5], p[7]);
            box = make_float4(__double2float_rn(__dmul_rn((double)x1, sx)), __double2float_rn(__dmul_rn((double)y1, sy)),
                              __double2float_rn(__dmul_rn((double)x2, sx)), __double2float_rn(__dmul_rn((double)y2, sy)));
            trk = det_track[lo + o];
        }
        *reinterpret_cast<float4*>(out.bbox + (slot0 + o) * 4) = box;
        out.track[slot0 + o] = trk;
        out.temporal[slot0 + o] = frame_idx;                           // padded slots keep the frame index
        out.ocr_mask[slot0 + o] = o < k ? 1 : 0;
    }
    for (int e = tid; e < Of * width; e += blockDim.x) {
        const int o = e / width, c = e - o * width;
        unsigned char v;
        if (o < k) v = det_tokens[(long long)(lo + o) * width + c];
        else v = c < 5 ? (unsigned char)"<pad>"[c] : 0;                // the dataset's literal pad token (dataset.py:140)
        out.tokens[slot0 * width + e] = v;
    }
}

}  // namespace t2s

extern "C" int t2s_pack_ocr_frames(const float* det_points, const long long* det_track, const unsigned char* det_tokens,
                                   int width, const int* frame_ptr, const int* info_base, const int* n_info,
                                   const int* n_frames, const double* vid_w, const double* vid_h, int B, int F, int Of,
                                   float* ocr_bbox, long long* track_id, long long* temporal_id, long long* ocr_mask,
                                   long long* frame_id, long long* frame_mask, long long* frame_num,
                                   long long* mid_frame_id, long long* mid_frame_idx, unsigned char* ocr_token_bytes,
                                   void* stream) {
    using namespace t2s;
    if (B <= 0 || F <= 0 || Of <= 0 || width < 5 || B > 65535) {
        set_error("t2s_pack_ocr_frames: bad shape (B %d, F %d, Of %d, width %d >= 5)", B, F, Of, width);
        return T2S_ERR_SHAPE;
    }
    if (!frame_ptr || !info_base || !n_info || !n_frames || !vid_w || !vid_h || !ocr_bbox || !track_id || !temporal_id ||
        !ocr_mask || !frame_id || !frame_mask || !frame_num || !mid_frame_id || !mid_frame_idx || !ocr_token_bytes) {
        set_error("t2s_pack_ocr_frames: null pointer");
        return T2S_ERR_ARG;
    }
    if (reinterpret_cast<uintptr_t>(ocr_bbox) & 15) { set_error("t2s_pack_ocr_frames: ocr_bbox needs 16-byte alignment"); return T2S_ERR_ALIGN; }
    PackOut out{ocr_bbox, track_id, temporal_id, ocr_mask, frame_id, frame_mask, frame_num, mid_frame_id, mid_frame_idx,
                ocr_token_bytes};
    pack_ocr_frames_kernel<<<dim3(F, B), 128, 0, (cudaStream_t)stream>>>(det_points, det_track, det_tokens, width,
                                                                        frame_ptr, info_base, n_info, n_frames, vid_w,
                                                                        vid_h, F, Of, out);
    return launch_status("t2s_pack_ocr_frames");
}

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
