// K5: question pooling, question->frame / question->OCR similarity, masked softmax,
// Gumbel pos/neg split, temporal top-k and per-frame spatial top-k -- fp32, sync-free.
//
// Reference call sites replaced (all of which round-trip to the host there):
//   Grounding_Module._calculate_self_attn            pythia/models/t2s.py:453-459
//   AttentionScore.forward (x4 instances)            pythia/modules/spatio_temporal_grounding.py:15-23
//   Temporal_Grounding_Indicator.forward             stg.py:34-68   (.item() x3, nonzero)
//   frame id -> OCR slot mask                        t2s.py:486-494 (nonzero)
//   Spatial_Grounding_Indicator.forward              stg.py:79-142  (masked_select)
//   PostHoc_Attention.forward (M4C)                  pythia/models/m4c.py:356-422
//
// Selection rule: rank counting with "lowest index first" among equal scores, which is what the
// reference's stable torch.sort gives for the spatial stage; for torch.topk (temporal stage) the
// reference's tie order is implementation-defined (SURVEY hard part 3) and ties only occur among
// -10000 entries.  Gumbel noise (= -log(Exp(1))) is an input so CPU and GPU runs can share it.
#include "common.cuh"
#include "../../include/t2s_b200.h"

namespace t2s {

constexpr float kMasked = -10000.0f;

// ------------------------------------------------------------------------------- masks
// joint[b] = [ arange(Lt) < text_len[b] | frame_mask[b] | ocr_mask[b] ]  as fp32 (t2s.py:726-732, Q11)
__global__ void mask_prep_kernel(const long long* __restrict__ text_len, const long long* __restrict__ frame_mask,
                                 const long long* __restrict__ ocr_mask, int B, int Lt, int F, int O,
                                 float* __restrict__ joint) {
    const int L = Lt + F + O;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)B * L;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / L), p = (int)(i % L);
        float v;
        if (p < Lt) v = p < text_len[b] ? 1.f : 0.f;
        else if (p < Lt + F) v = (float)frame_mask[(long long)b * F + p - Lt];
        else v = (float)ocr_mask[(long long)b * O + p - Lt - F];
        joint[i] = v;
    }
}

// key_idx[b, 0..n) = positions p with mask[b, p] != 0 (ascending); one warp per sample
__global__ void build_keys_kernel(const float* __restrict__ mask, int B, int L, int* __restrict__ key_idx,
                                  int* __restrict__ n_keys, int key_stride) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    int count = 0;
    for (int p0 = 0; p0 < L; p0 += 32) {
        const int p = p0 + lane;
        const bool on = p < L && mask[(long long)b * L + p] != 0.f;
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        if (on) key_idx[(long long)b * key_stride + count + __popc(bal & ((1u << lane) - 1))] = p;
        count += __popc(bal);
    }
    if (lane == 0) n_keys[b] = count;
}

// ------------------------------------------------------------------------------- question pooling
// qp: [B, Lt, H] = q_linear(txt).  attn = softmax_i(w.qp_i + bw) over ALL Lt tokens, then
// attn *= mask; attn /= (sum + 1e-12); gq = sum_i attn_i qp_i.
__global__ void __launch_bounds__(256)
question_pool_kernel(const float* __restrict__ qp, int Lt, int H, const float* __restrict__ w, const float* __restrict__ bw,
                     const float* __restrict__ txt_mask, int mask_stride, float* __restrict__ gq) {
    extern __shared__ float sa[];    // [Lt]
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    const float* q = qp + (long long)b * Lt * H;
    for (int i = warp; i < Lt; i += nw) {
        float s = 0.f;
        for (int d = lane; d < H; d += 32) s = fmaf(q[(long long)i * H + d], w[d], s);
        s = warp_sum(s);
        if (lane == 0) sa[i] = s + bw[0];
    }
    __syncthreads();
    if (warp == 0) {
        float mx = -INFINITY;
        for (int i = lane; i < Lt; i += 32) mx = fmaxf(mx, sa[i]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int i = lane; i < Lt; i += 32) sum += expf(sa[i] - mx);
        sum = warp_sum(sum);
        float msum = 0.f;
        for (int i = lane; i < Lt; i += 32) {
            const float a = (expf(sa[i] - mx) / sum) * txt_mask[(long long)b * mask_stride + i];
            sa[i] = a;
            msum += a;
        }
        msum = warp_sum(msum) + 1e-12f;
        for (int i = lane; i < Lt; i += 32) sa[i] = sa[i] / msum;
    }
    __syncthreads();
    for (int d = tid; d < H; d += blockDim.x) {
        float acc = 0.f;
        for (int i = 0; i < Lt; ++i) acc = fmaf(sa[i], q[(long long)i * H + d], acc);
        gq[(long long)b * H + d] = acc;
    }
}

// sim[b, n] = gq[b] . X[b, row0 + n]   (no projection, no scale: stg.py:17); one warp per row
__global__ void __launch_bounds__(256)
sim_scores_kernel(const float* __restrict__ gq, const float* __restrict__ X, long long batch_stride, long long ldx,
                  int row0, int N, int H, int B, float* __restrict__ sim) {
    const long long w = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= (long long)B * N) return;
    const int b = (int)(w / N), n = (int)(w % N);
    const float* x = X + b * batch_stride + (long long)(row0 + n) * ldx;
    const float* q = gq + (long long)b * H;
    float s = 0.f;
    for (int d = lane * 4; d < H; d += 128) {
        const float4 a = *reinterpret_cast<const float4*>(x + d);
        const float4 c = *reinterpret_cast<const float4*>(q + d);
        s = fmaf(a.x, c.x, s); s = fmaf(a.y, c.y, s); s = fmaf(a.z, c.z, s); s = fmaf(a.w, c.w, s);
    }
    s = warp_sum(s);
    if (lane == 0) sim[w] = s;
}

// AttentionScore tail + Gumbel-hard split for one element (stg.py:18-23, 41-50; F.gumbel_softmax hard=True)
struct Split { float att, pos_score, neg_score; };
__device__ __forceinline__ Split gumbel_split(float att, float m, float g0, float g1) {
    const float y0 = att + g0, y1 = att + g1;                 // tau = 1
    const float mx = fmaxf(y0, y1);
    const float e0 = expf(y0 - mx), e1 = expf(y1 - mx);
    const float s0 = e0 / (e0 + e1), s1 = e1 / (e0 + e1);
    const bool pick1 = s1 > s0;                                // max() returns the first index on ties
    const float hard0 = ((pick1 ? 0.f : 1.f) - s0) + s0;       // y_hard - y_soft + y_soft
    const float hard1 = ((pick1 ? 1.f : 0.f) - s1) + s1;
    const float pm = hard0 * m, nm = hard1 * m;
    Split r;
    r.att = att;
    r.pos_score = pm == 0.f ? kMasked : att * pm;
    r.neg_score = nm == 0.f ? kMasked : att * nm;
    return r;
}

// softmax over n logits in shared memory followed by mask-renormalise; in place -> att (with -10000 on masked)
template <typename TM>
__device__ void masked_attention(float* s, const TM* m, int n, float* red) {
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, s[i]);
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sum += expf(s[i] - mx);
    sum = block_sum(sum, red);
    float ms = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float a = (expf(s[i] - mx) / sum) * (float)m[i];
        s[i] = a;
        ms += a;
    }
    ms = block_sum(ms, red) + 1e-12f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = (float)m[i] == 0.f ? kMasked : s[i] / ms;
    __syncthreads();
}

// ------------------------------------------------------------------------------- temporal stage
__global__ void __launch_bounds__(256)
temporal_select_kernel(const float* __restrict__ sim, int sim_stride, const float* __restrict__ joint_mask, int L,
                       int Lt, int F, int O, int Of, const float* __restrict__ gumbel /*[B,2,F]*/,
                       const long long* __restrict__ frame_id, const long long* __restrict__ temporal_id, int topk,
                       const float* __restrict__ pos_override /*[B,F] or null; <0 = compute*/,
                       const float* __restrict__ neg_override /*[B,F] or null; <0 = compute*/,
                       long long* __restrict__ ground_frame,
                       float* __restrict__ pos_joint, float* __restrict__ neg_joint, float* __restrict__ slot_mask,
                       float* __restrict__ dbg_score) {
    extern __shared__ float sm[];
    float* att = sm;            // [F]
    float* msk = att + F;       // [F]
    float* pos = msk + F;       // [F]
    float* neg = pos + F;       // [F]
    float* sel = neg + F;       // [F] pos top-k flag
    long long* gf = reinterpret_cast<long long*>(sel + F + (F & 1));   // [topk], 8-byte aligned
    __shared__ float red[33];
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int f = tid; f < F; f += blockDim.x) {
        att[f] = sim[(long long)b * sim_stride + f];
        msk[f] = joint_mask[(long long)b * L + Lt + f];
    }
    __syncthreads();
    masked_attention(att, msk, F, red);
    for (int f = tid; f < F; f += blockDim.x) {
        const Split s = gumbel_split(att[f], msk[f], gumbel[((long long)b * 2) * F + f], gumbel[((long long)b * 2 + 1) * F + f]);
        pos[f] = s.pos_score;
        neg[f] = s.neg_score;
        if (dbg_score) dbg_score[(long long)b * F + f] = s.pos_score;
    }
    __syncthreads();
    for (int f = tid; f < F; f += blockDim.x) {
        const float pv = pos[f], nv = neg[f];
        int rp = 0, rn = 0;
        for (int j = 0; j < F; ++j) {
            rp += (pos[j] > pv) || (pos[j] == pv && j < f);     // largest-k
            rn += (neg[j] < nv) || (neg[j] == nv && j < f);     // smallest-k
        }
        float ps = rp < topk ? 1.f : 0.f;
        float ns = rn < topk ? 1.f : 0.f;
        // test hook: the reference's torch.topk tie order among -10000 entries is implementation-defined
        if (pos_override && pos_override[(long long)b * F + f] >= 0.f) ps = pos_override[(long long)b * F + f];
        if (neg_override && neg_override[(long long)b * F + f] >= 0.f) ns = neg_override[(long long)b * F + f];
        sel[f] = ps;
        pos_joint[(long long)b * L + Lt + f] = ps * msk[f];       // ground_frame_mask * frame_mask (t2s.py:480)
        neg_joint[(long long)b * L + Lt + f] = ns * msk[f];       // t2s.py:481
    }
    __syncthreads();
    for (int f = tid; f < F; f += blockDim.x) {
        if (sel[f] != 0.f) {
            int p = 0;
            for (int j = 0; j < f; ++j) p += sel[j] != 0.f;
            const long long id = frame_id[(long long)b * F + f];     // ascending position order (Q8)
            ground_frame[(long long)b * topk + p] = id;
            gf[p] = id == 0 ? 1 : id;                                // Q9
        }
    }
    __syncthreads();
    // OCR slots of the grounded frames, pads included (t2s.py:486-494, Q4)
    for (int o = tid; o < O; o += blockDim.x) {
        const long long t = temporal_id[(long long)b * O + o];
        bool hit = false;
        for (int k = 0; k < topk; ++k) hit |= (t == gf[k]);
        slot_mask[(long long)b * O + o] = hit ? 1.f : 0.f;
    }
    // question-token part of the pos/neg joint masks is the text mask itself
    for (int i = tid; i < Lt; i += blockDim.x) {
        const float v = joint_mask[(long long)b * L + i];
        pos_joint[(long long)b * L + i] = v;
        neg_joint[(long long)b * L + i] = v;
    }
}

// ------------------------------------------------------------------------------- spatial stage
// mode 0 (T2S, stg.py:79-142): attention mask = slot_mask; Gumbel split; pos top-k per frame over ALL frames
//   (not multiplied by the mask, Q2/Q3), neg top-k * slot_mask; ground_box = boxes of pos top-k in slot order.
// mode 1 (M4C PostHoc, m4c.py:395-421): attention mask = ocr_mask, no Gumbel; top-k per frame * slot_mask selects
//   the boxes (x ocr_mask at those slots); the answer transformer's OCR mask = slot_mask * ocr_mask.
__global__ void __launch_bounds__(256)
spatial_select_kernel(const float* __restrict__ sim, int sim_stride, int sim_off, const float* __restrict__ slot_mask,
                      const float* __restrict__ joint_mask, int L, int ocr_off, int F, int O, int Of,
                      const float* __restrict__ gumbel /*[B,2,O]*/, const float* __restrict__ boxes, int topk, int mode,
                      float* __restrict__ ground_box, float* __restrict__ pos_joint, float* __restrict__ neg_joint,
                      float* __restrict__ dbg_score) {
    extern __shared__ float sm[];
    float* att = sm;           // [O]
    float* pos = att + O;      // [O]
    float* neg = pos + O;      // [O]
    // the attention mask is 0 / 1: one byte per slot, so that the long-video shapes of the stress sweep (256 frames x 60
    // OCR slots = 15 360 slots, 13 B each) still fit the 227 KB of shared memory
    unsigned char* msk = reinterpret_cast<unsigned char*>(neg + O);      // [O]
    __shared__ float red[33];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int kk = topk < Of ? topk : Of;
    for (int o = tid; o < O; o += blockDim.x) {
        att[o] = sim[(long long)b * sim_stride + sim_off + o];
        const float mv = (mode != 1 && mode != 4) ? slot_mask[(long long)b * O + o] : joint_mask[(long long)b * L + ocr_off + o];
        msk[o] = mv != 0.f ? 1 : 0;
    }
    __syncthreads();
    masked_attention(att, msk, O, red);
    for (int o = tid; o < O; o += blockDim.x) {
        if (mode == 0 || mode == 3) {
            const Split s = gumbel_split(att[o], (float)msk[o], gumbel[((long long)b * 2) * O + o], gumbel[((long long)b * 2 + 1) * O + o]);
            pos[o] = s.pos_score;
            neg[o] = s.neg_score;
        } else {
            pos[o] = att[o];
            neg[o] = att[o];
        }
        if (dbg_score) dbg_score[(long long)b * O + o] = pos[o];
    }
    __syncthreads();
    if (mode == 4) {
        // T5-ViteVQA post-hoc attention (models/t5vitevqa.py:396-408): the `topk` OCR tokens with the largest score over
        // ALL frames (stable order: lowest index first among equals); ground_box [B, topk, 4] = their boxes in slot
        // order, zeroed where the slot is padding.  The joint masks are not touched (the answer transformer sees the
        // dataset masks).
        for (int o = tid; o < O; o += blockDim.x) {
            const float pv = pos[o];
            int r = 0;
            for (int j = 0; j < O; ++j) r += (pos[j] > pv) || (pos[j] == pv && j < o);
            neg[o] = r < topk ? 1.f : 0.f;
        }
        __syncthreads();
        for (int o = tid; o < O; o += blockDim.x) {
            if (neg[o] == 0.f) continue;
            int before = 0;
            for (int j = 0; j < o; ++j) before += neg[j] != 0.f;
            const float om = (float)msk[o];
            float4 bx = *reinterpret_cast<const float4*>(boxes + ((long long)b * O + o) * 4);
            bx.x *= om; bx.y *= om; bx.z *= om; bx.w *= om;
            *reinterpret_cast<float4*>(ground_box + ((long long)b * topk + before) * 4) = bx;
        }
        return;
    }
    for (int o = tid; o < O; o += blockDim.x) {
        const int f = o / Of, i = o % Of, base = f * Of;
        const float pv = pos[o], nv = neg[o];
        int rp = 0, rn = 0, before = 0;
        for (int j = 0; j < Of; ++j) {
            rp += (pos[base + j] > pv) || (pos[base + j] == pv && j < i);   // stable descending sort rank
            rn += (neg[base + j] < nv) || (neg[base + j] == nv && j < i);   // stable ascending sort rank
        }
        const bool psel = rp < kk;
        if (mode == 2) {
            // ablation "w/o SG" (models/t2s_wo_sg.py:503-506): every slot of the grounded frames is positive, every
            // other slot negative (pads included), ground_box = the boxes of the positive slots in slot order
            const float slot = (float)msk[o];
            pos_joint[(long long)b * L + ocr_off + o] = slot;
            neg_joint[(long long)b * L + ocr_off + o] = 1.f - slot;
            if (slot != 0.f) {
                for (int j = 0; j < o; ++j) before += msk[j] != 0;
                if (before < topk * Of)
                    *reinterpret_cast<float4*>(ground_box + ((long long)b * topk * Of + before) * 4) =
                        *reinterpret_cast<const float4*>(boxes + ((long long)b * O + o) * 4);
            }
        } else if (mode == 0 || mode == 3) {
            // mode 3 = ablation "w/o TG" (models/t2s_wo_tg.py:503-507): both masks are also multiplied by ocr_mask
            const float om = mode == 3 ? joint_mask[(long long)b * L + ocr_off + o] : 1.f;
            pos_joint[(long long)b * L + ocr_off + o] = (psel ? 1.f : 0.f) * om;
            neg_joint[(long long)b * L + ocr_off + o] = (rn < kk ? 1.f : 0.f) * (float)msk[o] * om;
            if (psel) {
                // masked_select keeps slot order: index inside the frame = #selected slots before this one
                for (int j = 0; j < i; ++j) {
                    const float pj = pos[base + j];
                    int r = 0;
                    for (int q = 0; q < Of; ++q) r += (pos[base + q] > pj) || (pos[base + q] == pj && q < j);
                    before += r < kk;
                }
                const float4 bx = *reinterpret_cast<const float4*>(boxes + ((long long)b * O + o) * 4);
                *reinterpret_cast<float4*>(ground_box + ((long long)b * F * kk + f * kk + before) * 4) = bx;
            }
        } else {
            const float slot = slot_mask[(long long)b * O + o];
            const float om = joint_mask[(long long)b * L + ocr_off + o];
            pos_joint[(long long)b * L + ocr_off + o] = slot * om;      // middle_ocr_mask (m4c.py:390)
            if (psel && slot != 0.f) {
                for (int j = 0; j < i; ++j) {
                    const float pj = pos[base + j];
                    int r = 0;
                    for (int q = 0; q < Of; ++q) r += (pos[base + q] > pj) || (pos[base + q] == pj && q < j);
                    before += r < kk;
                }
                float4 bx = *reinterpret_cast<const float4*>(boxes + ((long long)b * O + o) * 4);
                bx.x *= om; bx.y *= om; bx.z *= om; bx.w *= om;
                *reinterpret_cast<float4*>(ground_box + ((long long)b * kk + before) * 4) = bx;
            }
        }
    }
}

// OCR slots whose temporal id equals one of n_ids frame ids per sample, id 0 (padding) read as 1
// (models/t2s.py:486-494 with `tensor1` = any id list; models/t2s_wo_tg.py:483-495 passes sample_list.frame_id)
__global__ void frame_slots_kernel(const long long* __restrict__ ids, int n_ids, const long long* __restrict__ temporal_id,
                                   int B, int O, float* __restrict__ slot_mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * O) return;
    const int b = (int)(i / O);
    const long long t = temporal_id[i];
    bool hit = false;
    for (int k = 0; k < n_ids; ++k) {
        const long long id = ids[(long long)b * n_ids + k];
        hit |= t == (id == 0 ? 1 : id);
    }
    slot_mask[i] = hit ? 1.f : 0.f;
}

// ablation "w/o TG" (models/t2s_wo_tg.py:511-535): the frame masks are derived from the OCR masks -- the first n_pick
// frames that own a positive (negative) OCR slot; a sample with fewer gets index -1, which the reference's advanced
// indexing reads as the LAST frame.  ground_frame = the picked positive frame POSITIONS (not ids), -1 padded.
// One CTA per sample; also copies the question part of the joint mask like temporal_select does.
__global__ void __launch_bounds__(128)
frames_from_ocr_kernel(const float* __restrict__ joint_mask, float* __restrict__ pos_joint, float* __restrict__ neg_joint,
                       int L, int Lt, int F, int Of, int n_pick, long long* __restrict__ ground_frame) {
    extern __shared__ unsigned char s_any[];     // [2][F]
    const int b = blockIdx.x, tid = threadIdx.x;
    float* jm[2] = {pos_joint + (long long)b * L, neg_joint + (long long)b * L};
    for (int f = tid; f < 2 * F; f += blockDim.x) {
        const int which = f / F, fr = f % F;
        const float* row = jm[which] + Lt + F + fr * Of;
        bool any = false;
        for (int j = 0; j < Of; ++j) any |= row[j] != 0.f;
        s_any[f] = any;
    }
    __syncthreads();
    for (int f = tid; f < 2 * F; f += blockDim.x) {
        const int which = f / F, fr = f % F;
        int before = 0, total = 0;
        for (int j = 0; j < F; ++j) {
            total += s_any[which * F + j];
            if (j < fr) before += s_any[which * F + j];
        }
        bool on = s_any[f] && before < n_pick;
        if (fr == F - 1 && total < n_pick) on = true;                 // index -1
        jm[which][Lt + fr] = on ? 1.f : 0.f;
        if (which == 0) {
            if (s_any[f] && before < n_pick) ground_frame[(long long)b * n_pick + before] = fr;
            if (fr == 0)
                for (int k = total; k < n_pick; ++k) ground_frame[(long long)b * n_pick + k] = -1;
        }
    }
    for (int i = tid; i < Lt; i += blockDim.x) {
        const float v = joint_mask[(long long)b * L + i];
        jm[0][i] = v;
        jm[1][i] = v;
    }
}

// M4C: slot mask of the middle frame (m4c.py:376-385) and obj part of the joint mask = ones (m4c.py:420)
__global__ void middle_frame_slots_kernel(const long long* __restrict__ mid_id, const long long* __restrict__ temporal_id,
                                          int B, int O, float* __restrict__ slot_mask) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)B * O;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / O);
        slot_mask[i] = temporal_id[i] == mid_id[b] ? 1.f : 0.f;
    }
}

}  // namespace t2s

using namespace t2s;

extern "C" int t2s_mask_prep(const long long* text_len, const long long* frame_mask, const long long* ocr_mask, int B,
                             int Lt, int F, int O, float* joint, void* stream) {
    const long long n = (long long)B * (Lt + F + O);
    if (n <= 0) { set_error("mask_prep: empty"); return T2S_ERR_SHAPE; }
    int grid = (int)((n + 255) / 256);
    mask_prep_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(text_len, frame_mask, ocr_mask, B, Lt, F, O, joint);
    return launch_status("mask_prep");
}

extern "C" int t2s_build_keys(const float* mask, int B, int L, int* key_idx, int* n_keys, int key_stride, void* stream) {
    if (B <= 0 || L <= 0 || key_stride < L) { set_error("build_keys: bad shape"); return T2S_ERR_SHAPE; }
    build_keys_kernel<<<(B + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(mask, B, L, key_idx, n_keys, key_stride);
    return launch_status("build_keys");
}

extern "C" int t2s_question_pool(const float* qp, int B, int Lt, int H, const float* w, const float* bw,
                                 const float* txt_mask, int mask_stride, float* gq, void* stream) {
    if (B <= 0 || Lt <= 0 || H <= 0) { set_error("question_pool: bad shape"); return T2S_ERR_SHAPE; }
    question_pool_kernel<<<B, 256, Lt * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(qp, Lt, H, w, bw, txt_mask, mask_stride, gq);
    return launch_status("question_pool");
}

extern "C" int t2s_sim_scores(const float* gq, const float* X, long long batch_stride, long long ldx, int row0, int N,
                              int H, int B, float* sim, void* stream) {
    if (B <= 0 || N <= 0 || (H % 4) || (ldx % 4)) { set_error("sim_scores: bad shape"); return T2S_ERR_SHAPE; }
    const long long rows = (long long)B * N;
    sim_scores_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(gq, X, batch_stride, ldx, row0, N, H, B, sim);
    return launch_status("sim_scores");
}

extern "C" int t2s_temporal_select(const float* sim, int sim_stride, const float* joint_mask, int B, int Lt, int F,
                                   int Of, const float* gumbel, const long long* frame_id, const long long* temporal_id,
                                   int topk, const float* pos_override, const float* neg_override,
                                   long long* ground_frame, float* pos_joint,
                                   float* neg_joint, float* slot_mask, float* dbg_score, void* stream) {
    if (B <= 0 || F <= 0 || topk <= 0 || topk > F) { set_error("temporal_select: need 0 < topk <= frames"); return T2S_ERR_SHAPE; }
    const int O = F * Of, L = Lt + F + O;
    const size_t smem = (5 * F + (F & 1)) * sizeof(float) + topk * sizeof(long long) + 8;
    temporal_select_kernel<<<B, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        sim, sim_stride, joint_mask, L, Lt, F, O, Of, gumbel, frame_id, temporal_id, topk, pos_override, neg_override, ground_frame,
        pos_joint, neg_joint, slot_mask, dbg_score);
    return launch_status("temporal_select");
}

extern "C" int t2s_spatial_select(const float* sim, int sim_stride, int sim_off, const float* slot_mask,
                                  const float* joint_mask, int B, int L_joint, int ocr_off, int F, int Of,
                                  const float* gumbel, const float* boxes, int topk, int mode, float* ground_box,
                                  float* pos_joint, float* neg_joint, float* dbg_score, void* stream) {
    if (B <= 0 || F <= 0 || Of <= 0 || topk <= 0) { set_error("spatial_select: bad shape"); return T2S_ERR_SHAPE; }
    const int O = F * Of, L = L_joint;
    if (ocr_off < 0 || ocr_off + O > L) { set_error("spatial_select: OCR part outside the joint mask"); return T2S_ERR_SHAPE; }
    const size_t smem = 3 * (size_t)O * sizeof(float) + (((size_t)O + 15) & ~(size_t)15);
    static size_t attr = 48 * 1024;
    if (smem > attr) {
        if (smem > 227 * 1024) { set_error("spatial_select: %d OCR slots exceed shared memory", O); return T2S_ERR_SHAPE; }
        cudaError_t e = cudaFuncSetAttribute(spatial_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("spatial_select attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = smem;
    }
    spatial_select_kernel<<<B, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        sim, sim_stride, sim_off, slot_mask, joint_mask, L, ocr_off, F, O, Of, gumbel, boxes, topk, mode, ground_box,
        pos_joint, neg_joint, dbg_score);
    return launch_status("spatial_select");
}

extern "C" int t2s_middle_frame_slots(const long long* mid_id, const long long* temporal_id, int B, int O,
                                      float* slot_mask, void* stream) {
    const long long n = (long long)B * O;
    if (n <= 0) { set_error("middle_frame_slots: empty"); return T2S_ERR_SHAPE; }
    middle_frame_slots_kernel<<<(int)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(mid_id, temporal_id, B, O, slot_mask);
    return launch_status("middle_frame_slots");
}

extern "C" int t2s_frame_slots(const long long* ids, int n_ids, const long long* temporal_id, int B, int O,
                               float* slot_mask, void* stream) {
    const long long n = (long long)B * O;
    if (n <= 0 || n_ids <= 0) { set_error("frame_slots: empty"); return T2S_ERR_SHAPE; }
    frame_slots_kernel<<<(int)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ids, n_ids, temporal_id, B, O, slot_mask);
    return launch_status("frame_slots");
}

extern "C" int t2s_frames_from_ocr(const float* joint_mask, float* pos_joint, float* neg_joint, int B, int Lt, int F,
                                   int Of, int n_pick, long long* ground_frame, void* stream) {
    if (B <= 0 || F <= 0 || Of <= 0 || n_pick <= 0) { set_error("frames_from_ocr: bad shape"); return T2S_ERR_SHAPE; }
    const int L = Lt + F + F * Of;
    frames_from_ocr_kernel<<<B, 128, 2 * F, reinterpret_cast<cudaStream_t>(stream)>>>(joint_mask, pos_joint, neg_joint, L, Lt,
                                                                                   F, Of, n_pick, ground_frame);
    return launch_status("frames_from_ocr");
}
