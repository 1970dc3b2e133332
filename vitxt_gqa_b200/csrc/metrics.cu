// K9: the evaluation step that consumes the forward's outputs, on the device.
//
// Replaces, per batch, what the reference does on the host after two device->host copies of the whole output:
//   * `pos_scores.argmax(-1)` + the python loop that cuts the answer at EOS and tells vocabulary ids from OCR copies
//     (pythia/modules/metrics.py:186-207, repeated at 395-416 and 498-519)            -> answer_decode
//   * BoxGroundAccuracyEvaluator.eval_pred_list / check_iou / calculate_iou over `.tolist()`ed frames and boxes
//     (pythia/utils/m4c_evaluators.py:331-405, called from metrics.py:233-339 and 341-546) and
//     TempGroundAccuracyEvaluator.eval_pred_list (m4c_evaluators.py:301-328)           -> ground_metrics
// The string side of the answer metrics (vocabulary lookup, EvalAI normalisation, soft accuracy, ANLS) stays on the
// host: it needs B x T small integers from here instead of the B x T x (V + O) score tensor.
//
// HBM-bound integer / fp64 work.  answer_decode reads the score tensor once (one CTA per decoding row, 16-byte
// loads, lowest index wins ties like torch.argmax on CPU).  ground_metrics is a few hundred fp64 operations per
// sample: one thread per sample walks the packed annotation of that question in the evaluator's order, because
// the evaluator's result depends on that order (see the quirks below).  All IoU arithmetic is IEEE binary64 with
// explicit round-to-nearest intrinsics in the order of the python source, so the hit / miss decisions are
// bit-identical to the reference evaluator.
//
// Quirks of the reference evaluator that are reproduced on purpose (tests pin them against the real one):
//   E1 every (annotated span, grounded frame) hit whose frame has a labelled box appends a 1 to the score list when
//      the best IoU passes the threshold, so one sample can contribute several 1s (m4c_evaluators.py:363-364,397);
//   E2 `flag` is overwritten by every check, so the trailing 0 is appended iff the LAST check failed or there was
//      none (m4c_evaluators.py:366,397-400);
//   E3 the boxes paired with the i-th grounded frame are pred_box[i*ocr_topk:(i+1)*ocr_topk] -- an index into the
//      list of ALL per-frame top-k boxes, not the boxes of that frame (m4c_evaluators.py:394);
//   E4 accuracy = sum(list) / len(list) over the concatenated list, and GQA@x reads entry i of that list as "sample
//      i" (metrics.py:428-433).
#include "common.cuh"
#include "../../include/t2s_b200.h"

namespace t2s {

// ------------------------------------------------------------------------------------------------ answer decode
constexpr int kArgmaxThreads = 256;

__global__ void __launch_bounds__(kArgmaxThreads)
answer_argmax_kernel(const float* __restrict__ scores, long long ld, int N, int* __restrict__ ids) {
    __shared__ float s_v[kArgmaxThreads / 32];
    __shared__ int s_i[kArgmaxThreads / 32];
    const float* row = scores + (long long)blockIdx.x * ld;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
    const int n4 = vec ? (N >> 2) : 0;
    for (int i = threadIdx.x; i < n4; i += kArgmaxThreads) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row) + i);
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (e[j] > best) { best = e[j]; bi = 4 * i + j; }      // strict: the lowest index of a tie stays
    }
    for (int i = 4 * n4 + threadIdx.x; i < N; i += kArgmaxThreads) {
        const float e = row[i];
        if (e > best) { best = e; bi = i; }
    }
    // a row of -inf only: every thread still holds bi = INT_MAX; index 0 is what torch.argmax returns
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_v[wid] = best; s_i[wid] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kArgmaxThreads / 32; ++w)
            if (s_v[w] > best || (s_v[w] == best && s_i[w] < bi)) { best = s_v[w]; bi = s_i[w]; }
        ids[blockIdx.x] = bi == 0x7fffffff ? 0 : bi;
    }
}

// metrics.py:194-206: walk the T ids; an id >= V is an OCR copy and is always kept, a vocabulary id equal to EOS
// ends the answer (and is not part of it)
__global__ void answer_cut_kernel(const int* __restrict__ ids, int B, int T, int V, int eos, int* __restrict__ len) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int n = T;
    for (int t = 0; t < T; ++t) {
        const int id = ids[b * T + t];
        if (id < V && id == eos) { n = t; break; }
    }
    len[b] = n;
}

// ------------------------------------------------------------------------------------------------ grounding metrics
struct IouBox { double x1, y1, x2, y2; };

// m4c_evaluators.py:335-356, operation for operation (python floats are binary64)
__device__ __forceinline__ double iou_py(const IouBox& a, const IouBox& b) {
    const double x1 = fmax(a.x1, b.x1), y1 = fmax(a.y1, b.y1);
    const double x2 = fmin(a.x2, b.x2), y2 = fmin(a.y2, b.y2);
    const double w = fmax(0.0, __dadd_rn(__dsub_rn(x2, x1), 1.0));
    const double h = fmax(0.0, __dadd_rn(__dsub_rn(y2, y1), 1.0));
    const double inter = __dmul_rn(w, h);
    const double area_a = __dmul_rn(__dadd_rn(__dsub_rn(a.x2, a.x1), 1.0), __dadd_rn(__dsub_rn(a.y2, a.y1), 1.0));
    const double area_b = __dmul_rn(__dadd_rn(__dsub_rn(b.x2, b.x1), 1.0), __dadd_rn(__dsub_rn(b.y2, b.y1), 1.0));
    const double uni = __dsub_rn(__dadd_rn(area_a, area_b), inter);
    return __ddiv_rn(inter, uni);
}

__global__ void ground_metrics_kernel(const long long* __restrict__ ground_frame, int kf,
                                      const float* __restrict__ ground_box, int n_box, int ocr_topk,
                                      const int* __restrict__ rec_index, const int* __restrict__ span_ptr,
                                      const long long* __restrict__ span_st, const long long* __restrict__ span_ed,
                                      const int* __restrict__ box_ptr, const long long* __restrict__ box_frame,
                                      const double* __restrict__ box_xyxy, const double* __restrict__ rec_wh, int B,
                                      double thr_a, double thr_b, int* __restrict__ ones, int* __restrict__ tail_zero,
                                      int* __restrict__ t_hit, int* __restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int rec = rec_index[b];
    int st = 0;
    int n1[2] = {0, 0};
    bool flag[2] = {false, false};
    int temporal = 0;
    if (rec < 0) {
        st = 4;                      // no annotation for this question: the reference fails on `None['...']`
    } else {
        const double width = rec_wh[2 * rec], height = rec_wh[2 * rec + 1];
        const long long* gf = ground_frame + (long long)b * kf;
        const float* gb = ground_box + (long long)b * n_box * 4;
        for (int s = span_ptr[rec]; s < span_ptr[rec + 1]; ++s) {
            const long long f0 = span_st[s], f1 = span_ed[s];
            for (int i = 0; i < kf; ++i) {
                const long long fr = gf[i];
                if (!(f0 <= fr && fr <= f1)) continue;
                if (!temporal) temporal = 1;      // TempGroundAccuracyEvaluator: any grounded frame inside any span
                // bboxs_gt[str(frame - 1)] (m4c_evaluators.py:390-391): first entry of the span with that key
                int hit = -1;
                for (int j = box_ptr[s]; j < box_ptr[s + 1]; ++j)
                    if (box_frame[j] == fr - 1) { hit = j; break; }
                if (hit < 0) continue;
                IouBox gt = {box_xyxy[4 * hit], box_xyxy[4 * hit + 1], box_xyxy[4 * hit + 2], box_xyxy[4 * hit + 3]};
                if (!(gt.x1 <= gt.x2 && gt.y1 <= gt.y2)) st |= 1;             // `assert bbox_gt[0]<=bbox_gt[2] ...`
                double best = 0.0;
                const long long lo = (long long)i * ocr_topk;                  // E3; python slicing clamps
                const long long hi = min((long long)n_box, lo + ocr_topk);
                for (long long j = lo; j < hi; ++j) {
                    IouBox pb = {__dmul_rn((double)gb[4 * j], width), __dmul_rn((double)gb[4 * j + 1], height),
                                 __dmul_rn((double)gb[4 * j + 2], width), __dmul_rn((double)gb[4 * j + 3], height)};
                    if (!(pb.x1 <= pb.x2 && pb.y1 <= pb.y2)) st |= 2;         // `assert pred_bbox[0]<=pred_bbox[2] ...`
                    const double v = iou_py(gt, pb);
                    if (v > best) best = v;
                }
                flag[0] = best > thr_a;                                        // E2
                flag[1] = best > thr_b;
                n1[0] += flag[0];                                              // E1
                n1[1] += flag[1];
            }
        }
    }
    ones[b] = n1[0];
    ones[B + b] = n1[1];
    tail_zero[b] = flag[0] ? 0 : 1;
    tail_zero[B + b] = flag[1] ? 0 : 1;
    t_hit[b] = temporal;
    status[b] = st;
}

// E4: accuracy over the concatenated list and its first B entries, for both thresholds; temporal accuracy.
// One warp; lane t < 2 handles threshold t, lane 2 the temporal accuracy.
__global__ void ground_reduce_kernel(const int* __restrict__ ones, const int* __restrict__ tail_zero,
                                     const int* __restrict__ t_hit, int B, float* __restrict__ acc,
                                     int* __restrict__ head) {
    const int t = threadIdx.x;
    if (t < 2) {
        long long s1 = 0, n = 0;
        for (int b = 0; b < B; ++b) {
            const int o = ones[t * B + b], z = tail_zero[t * B + b];
            for (int j = 0; j < o + z && n + j < B; ++j) head[t * B + n + j] = j < o ? 1 : 0;
            s1 += o;
            n += o + z;
        }
        acc[t] = (float)__ddiv_rn((double)s1, (double)n);       // python: sum(ints) / len -> binary64 -> float32 tensor
    } else if (t == 2) {
        long long s = 0;
        for (int b = 0; b < B; ++b) s += t_hit[b];
        acc[2] = (float)__ddiv_rn((double)s, (double)B);
    }
}

}  // namespace t2s

using namespace t2s;

extern "C" int t2s_answer_decode(const float* scores, long long ld_scores, int B, int T, int N, int V, int eos_idx,
                                 int* ans_ids, int* ans_len, void* stream) {
    if (B <= 0 || T <= 0 || N <= 0 || V < 0 || V > N || ld_scores < N) {
        set_error("t2s_answer_decode: bad shape B=%d T=%d N=%d V=%d ld=%lld", B, T, N, V, ld_scores);
        return T2S_ERR_SHAPE;
    }
    if (!scores || !ans_ids || !ans_len) {
        set_error("t2s_answer_decode: null pointer");
        return T2S_ERR_ARG;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    answer_argmax_kernel<<<B * T, kArgmaxThreads, 0, st>>>(scores, ld_scores, N, ans_ids);
    int rc = launch_status("answer_argmax");
    if (rc) return rc;
    answer_cut_kernel<<<(B + 127) / 128, 128, 0, st>>>(ans_ids, B, T, V, eos_idx, ans_len);
    return launch_status("answer_cut");
}

extern "C" int t2s_ground_metrics(const long long* ground_frame, int kf, const float* ground_box, int n_box,
                                  int ocr_topk, const int* rec_index, const int* span_ptr, const long long* span_st,
                                  const long long* span_ed, const int* box_ptr, const long long* box_frame,
                                  const double* box_xyxy, const double* rec_wh, int B, double thr_a, double thr_b,
                                  int* ones, int* tail_zero, int* t_hit, int* status, float* acc, int* head,
                                  void* stream) {
    if (B <= 0 || kf <= 0 || n_box < 0 || ocr_topk < 0) {
        set_error("t2s_ground_metrics: bad shape B=%d kf=%d n_box=%d ocr_topk=%d", B, kf, n_box, ocr_topk);
        return T2S_ERR_SHAPE;
    }
    if (!ground_frame || !ground_box || !rec_index || !span_ptr || !box_ptr || !rec_wh || !ones || !tail_zero ||
        !t_hit || !status || !acc || !head) {
        set_error("t2s_ground_metrics: null pointer");
        return T2S_ERR_ARG;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    ground_metrics_kernel<<<(B + 63) / 64, 64, 0, st>>>(ground_frame, kf, ground_box, n_box, ocr_topk, rec_index,
                                                        span_ptr, span_st, span_ed, box_ptr, box_frame, box_xyxy,
                                                        rec_wh, B, thr_a, thr_b, ones, tail_zero, t_hit, status);
    int rc = launch_status("ground_metrics");
    if (rc) return rc;
    ground_reduce_kernel<<<1, 32, 0, st>>>(ones, tail_zero, t_hit, B, acc, head);
    return launch_status("ground_reduce");
}
