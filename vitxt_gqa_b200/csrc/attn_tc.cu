// K2t: masked multi-head attention on the 5th-gen tensor cores (tcgen05 + TMEM), head size 64.
//
// Replaces BertSelfAttention's matmul / +mask / softmax / matmul chain for the encoder rows of the
// answer transformer (reference pythia/models/t2s.py:622 through pytorch_transformers) and, in its
// bf16x3 form, for TextBert / QTV of the grounding chain (t2s.py:423,538).  Masked keys are skipped
// through the compacted per-sample key list (see attention.cu): exp(-10000 - max) == 0 in fp32.
//
// One CTA = 128 query rows of one (sample, head) at a time (TC_NQ consecutive query tiles per CTA); key tiles of 128
// gathered rows.
//   softmax warps (4, or 8 with two threads per row in the bf16x3 form): a thread owns (half of) query row r == TMEM
//              lane r.  One software-pipelined sweep over S in TMEM (exp2 against the carried maximum, row sum); P goes
//              back into TENSOR MEMORY as packed bf16 pairs (tcgen05.st, 64 columns per plane) and is the A operand of
//              P.V from there -- it never touches shared memory.  O accumulates in TMEM across key tiles; the online-
//              softmax correction is LAZY: the running maximum is only raised (and O rescaled in TMEM by tcgen05.ld /
//              st) when a tile exceeds it by more than 2^8, which keeps the result exact (the final division uses the
//              same maximum) and the rescale rare.
//   4 loader warps: cp.async 16-byte gathers of the K / V rows named by the key list into the same 128B-swizzled
//              layout a TMA tile load would produce (TMA cannot gather rows), two stages, K and V with their own
//              full / empty barriers, fence.proxy.async + mbarrier hand-off to the tensor core.
//   1 issuer warp: one thread issues tcgen05.mma: S = Q.K^T (M128 N128 K64, both operands from shared memory) and
//              O += P.V (M128 N64 K128: A = P from tensor memory, B = V consumed as an MN-major operand straight from
//              its row-per-key layout); tcgen05.commit drives the mbarriers.
// bf16 form: S(t+1) is issued right behind P.V(t); two CTAs per SM (80 KB smem, 256 TMEM columns each: S 128, O 64,
// P 64) overlap one CTA's softmax with the other's tensor work.
// X3 = true: q|k|v are bf16 hi|lo pairs, S = Ql.Kh + Qh.Kl + Qh.Kh, O = Pl.Vh + Ph.Vl + Ph.Vh (fp32-class; 162 KB smem,
// one CTA per SM, which therefore pipelines itself: two S buffers, S(t+1) issued under the softmax of tile t, the P of a
// tile held in registers until P.V(t-1) has released the P columns; 448 of 512 TMEM columns).
#include "common.cuh"
#include "../../include/t2s_b200.h"

namespace t2s {

constexpr int TC_BQ = 128, TC_BK = 128, TC_DH = 64;
// warps: NSW softmax, then 4 loaders, then 1 MMA issuer (NSW = 4, or 8 in the bf16x3 form)
constexpr int TC_TILE = 128 * 128;           // bytes of a [128 rows x 64 bf16] 128B-swizzled tile
// Query tiles per CTA.  A CTA that lives for one 128-row query tile spends ~4 us on things that are not attention
// (launch, barrier + TMEM set-up, the dependent index -> row gathers of Q and the first K/V tile, the drain of the last
// P.V, the O read-out) -- 35 % of the bf16 launches' time at the short `pos` / `neg` key lists.  With TC_NQ consecutive
// query tiles of one (sample, head) per CTA the set-up is paid once and Q / K / V of tile i + 1 are gathered while tile
// i is in its last softmax, P.V and read-out.
constexpr int TC_NQ = 3;

// Threads per query row in the bf16x3 form (2 or 4).  One CTA per SM: with two threads per row the eight softmax warps
// (two per scheduler) took ~3 000 cycles per 128 x 128 tile against ~1 500 of tensor work -- latency bound (tcgen05.ld,
// MUFU, the hi / lo conversions), not issue bound; four threads per row = 16 softmax warps, 32 key columns each -- measured
// on the t2s_abinet step: 1.65 ms against 1.62 ms with two, because the launches are already at ~0.7 of the power-capped
// tensor rate once the padded tile work is counted.  The default stays 2.
#ifndef T2S_TC_X3_HS
#define T2S_TC_X3_HS 2
#endif

template <bool X3>
struct TcCfg {
    static constexpr int NP = X3 ? 2 : 1;                     // hi (+ lo) planes
    static constexpr int Q_BYTES = NP * TC_TILE;               // Q tile(s)
    static constexpr int KV_STAGE = NP * 2 * TC_TILE;          // K plane(s) then V plane(s)
    // no alignment slack: two CTAs (2 x (114 816 + 1 024 reserved)) must fit the 227 KB of an SM, so the kernel
    // relies on the dynamic shared-memory window starting 1024-byte aligned (it has no static shared memory)
    // and traps otherwise
    // X3 runs one CTA per SM (shared memory), i.e. one softmax warp per scheduler if a thread owned a whole row; its
    // softmax also does twice the conversions (hi and lo planes).  So each row is split between two threads (columns
    // 0-63 / 64-127, warps w and w + 4 reach the same TMEM lane quarter): 8 softmax warps, two per scheduler, which
    // exchange the tile maximum (per tile) and the row sum (per query tile) through shared memory.
    static constexpr int HS = X3 ? T2S_TC_X3_HS : 1;          // threads per query row
    static constexpr int NSW = 4 * HS;                        // softmax warps
    static constexpr int THREADS = 32 * (NSW + 5);
    static constexpr int XCH_BYTES = HS > 1 ? 2 * HS * 128 * 4 : 0;     // [tile parity][part][row] floats
    // PIPE (the one-CTA-per-SM form): nothing else on the SM hides the softmax, so the CTA pipelines itself -- S lives in
    // two TMEM buffers and S(t+1) is issued BEFORE the tensor core waits for P(t), i.e. it runs under the softmax of tile
    // t; the softmax keeps the P of a tile in registers and stores it once P.V(t-1) has released the single P tile.  Per
    // key tile the tensor core then sees S + P.V back to back (2 x 768 cycles in the bf16x3 form) instead of
    // S + softmax + P.V in series.  512 TMEM columns: S 2 x 128, O 64.
    static constexpr bool PIPE = X3;
    static constexpr int TMEM_COLS = PIPE ? 512 : 256;
    static constexpr int SMEM = Q_BYTES + 2 * KV_STAGE + 128 + XCH_BYTES;
};

// MN-major (rows = K index, 64 contiguous N elements = one 128-byte row) 128B-swizzled operand: 8-row groups
// 1024 B apart (SBO); a single 64-element atom along N, so LBO is not used.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024 >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// 2^x by one MUFU.EX2 (flush-to-zero); exp2f() wraps the same instruction in denormal-range scaling (three more
// issue slots per element) that the softmax never needs: its arguments are <= 8 and results below 2^-126 may vanish
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// DROP (training step): attention_probs dropout of BertSelfAttention -- the normalised probabilities are masked before
// P.V, i.e. O = (1/l) sum_j p_j m_j v_j with the row sum l over the UNMASKED p: P is masked where it is written for
// the tensor core, the row sum is not, and the keep scale goes into the final 1/l.  Mask addressing: common.cuh.
template <bool X3, bool DROP>
__global__ void __launch_bounds__(TcCfg<X3>::THREADS, X3 ? 1 : 2)
attn_tc_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld, int lo_off, int L, int H,
               const int* __restrict__ key_idx, const int* __restrict__ n_keys, int key_stride,
               __nv_bfloat16* __restrict__ out, long long ldo, float scale_log2, DropCfg drop,
               float* __restrict__ lse_out, int lse_rows) {
    using Cfg = TcCfg<X3>;
    constexpr int NP = Cfg::NP;
    extern __shared__ __align__(1024) uint8_t tc_raw[];
    uint8_t* smem = tc_raw;
    if (smem_u32(smem) & 1023u) __trap();      // 128B-swizzled tiles need 1024-byte alignment
    uint8_t* sQ = smem;
    uint8_t* sKV = sQ + Cfg::Q_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + 2 * Cfg::KV_STAGE);
    uint64_t* k_full = bars;           // [2] count 128 (loader threads): K plane(s) of a stage have landed
    uint64_t* k_empty = bars + 2;      // [2] count 1   (tcgen05.commit behind S: the K plane(s) may be overwritten)
    uint64_t* v_full = bars + 4;       // [2] count 128
    uint64_t* v_empty = bars + 6;      // [2] count 1   (tcgen05.commit behind P.V)
    uint64_t* s_full = bars + 8;       // [2] count 1   (PIPE: one per S buffer; otherwise only [0])
    uint64_t* p_full = bars + 10;      // count 128 (softmax threads)
    uint64_t* o_done = bars + 11;      // count 1: committed behind the last P.V of a query tile
    uint64_t* q_empty = bars + 12;     // count 1: committed behind the last S of a query tile (Q may be overwritten)
    uint64_t* p_free = bars + 13;      // count 1: committed behind every P.V (PIPE: the P tile may be overwritten)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
    float* xch = reinterpret_cast<float*>(sKV + 2 * Cfg::KV_STAGE + 128);
    constexpr int HS = Cfg::HS, NSW = Cfg::NSW;
    constexpr bool PIPE = Cfg::PIPE;

    const int b = blockIdx.z, h = blockIdx.y;
    const int it0 = blockIdx.x * TC_NQ;                                    // first query tile of this CTA
    const int n_items = min(TC_NQ, (L + TC_BQ - 1) / TC_BQ - it0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = n_keys[b];
    const int nt = (nk + TC_BK - 1) / TC_BK;
    const int* kidx = key_idx + (long long)b * key_stride;
    const __nv_bfloat16* base = qkv + (long long)b * L * ld + h * TC_DH;

    if (warp == NSW + 4) {
        if (lane == 0) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(&k_full[i], 128); mbar_init(&v_full[i], 128);
                mbar_init(&k_empty[i], 1); mbar_init(&v_empty[i], 1);
                mbar_init(&s_full[i], 1);
            }
            mbar_init(p_full, 128 * HS); mbar_init(o_done, 1); mbar_init(q_empty, 1); mbar_init(p_free, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // columns: S (PIPE: two buffers) | O (64) | P as packed bf16 pairs, 64 per plane (the A operand of P.V, read by the
    // tensor core straight from tensor memory: 256 columns in the two-CTA form, 448 of 512 in the PIPE form)
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + (PIPE ? 256 : 128), tmem_P = tmem_O + 64;

    if (warp >= NSW && warp < NSW + 4) {
        // ------------------------------------------------------------------ loaders
        // K and V of a stage have their own full / empty barriers: the K plane(s) are free again as soon as S has
        // retired, long before the P.V that frees the V plane(s).
        const int lt = threadIdx.x - 32 * NSW;       // 0..127
        const int c = lt & 7, r0 = lt >> 3;          // 16-byte chunk / first row; rows r0 + 16 i
        auto load_q = [&](int q0) {              // Q tile(s): rows q0 .. q0+127 (zero-filled past L)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + 16 * i;
                const bool ok = q0 + r < L;
                const __nv_bfloat16* src = base + (long long)(ok ? q0 + r : 0) * ld + c * 8;
                const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
                cp_async16(sQ + off, src, ok);
                if (X3) cp_async16(sQ + TC_TILE + off, src + lo_off, ok);
            }
        };
        int krow[8];                             // gathered rows of the key tile in flight (-1: past the key list)
        auto load_k = [&](int t, int s) {
            uint8_t* dst = sKV + s * Cfg::KV_STAGE;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + 16 * i;
                const int k = t * TC_BK + r;
                krow[i] = k < nk ? kidx[k] : -1;
                const bool ok = krow[i] >= 0;
                const __nv_bfloat16* src = base + (long long)(ok ? krow[i] : 0) * ld + c * 8 + H;
                const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
                cp_async16(dst + off, src, ok);                                       // K (hi)
                if (X3) cp_async16(dst + TC_TILE + off, src + lo_off, ok);            // K lo
            }
        };
        auto load_v = [&](int s) {               // V rows of the tile load_k() was last called for
            uint8_t* dst = sKV + s * Cfg::KV_STAGE + NP * TC_TILE;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + 16 * i;
                const bool ok = krow[i] >= 0;
                const __nv_bfloat16* src = base + (long long)(ok ? krow[i] : 0) * ld + c * 8 + 2 * H;
                const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
                cp_async16(dst + off, src, ok);                                       // V (hi)
                if (X3) cp_async16(dst + TC_TILE + off, src + lo_off, ok);            // V lo
            }
        };
        auto publish = [&](uint64_t* bar) {
            fence_proxy_async();                     // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(bar);
        };
        int g = 0;                               // key tiles handled so far by this CTA: stage g & 1, phase (g >> 1) & 1
        for (int it = 0; it < n_items; ++it) {
            // Q of this query tile may land once every S of the previous one has retired; its first K/V tile goes to
            // the stage the tile before last has left.  The waits pass at once for the first query tile.
            mbar_wait(q_empty, (it & 1) ^ 1);
            mbar_wait(&k_empty[g & 1], ((g >> 1) & 1) ^ 1);
            load_q((it0 + it) * TC_BQ);
            load_k(0, g & 1);
            cp_async_commit();
            mbar_wait(&v_empty[g & 1], ((g >> 1) & 1) ^ 1);
            load_v(g & 1);
            cp_async_commit();
            if (PIPE) {
                // K runs one tile ahead of V.  In flight at the top of iteration t: V(t) and K(t+1), both requested a
                // whole iteration ago, so neither wait exposes a gather latency (index load + row gather, ~1.7 us,
                // longer than the ~1.3 us the tensor core needs per tile); K(t+1) is published as soon as it has landed
                // -- S(t+1) goes out under the softmax of tile t -- and only then does the loader block on the V planes
                // P.V(t-1) still reads.  K(t+2) follows into the planes S(t) has left (S(t) retires before P.V(t-1)).
                if (nt > 1) {
                    load_k(1, (g + 1) & 1);      // stage free: its last S belongs to the tile before last (waited above
                    cp_async_commit();           // for g, and for g + 1 one query tile / iteration earlier)
                    cp_async_wait<2>();
                } else {
                    cp_async_wait<1>();
                }
                publish(&k_full[g & 1]);
                for (int t = 0; t < nt; ++t, ++g) {
                    const bool more = t + 1 < nt;
                    const int s1 = (g + 1) & 1;
                    if (more) cp_async_wait<1>(); else cp_async_wait<0>();       // V(t)
                    publish(&v_full[g & 1]);
                    if (more) {
                        cp_async_wait<0>();                                      // K(t+1)
                        publish(&k_full[s1]);
                        mbar_wait(&v_empty[s1], (((g + 1) >> 1) & 1) ^ 1);
                        load_v(s1);
                        cp_async_commit();
                        if (t + 2 < nt) {
                            mbar_wait(&k_empty[g & 1], (((g + 2) >> 1) & 1) ^ 1);
                            load_k(t + 2, g & 1);
                            cp_async_commit();
                        }
                    }
                }
            } else {
                for (int t = 0; t < nt; ++t, ++g) {
                    // publish tile t BEFORE gathering tile t + 1: the other stage only frees when P.V(t-1) retires, and
                    // a clock trace showed S(t) waiting ~1500 cycles behind that wait and the issue of the next gathers
                    cp_async_wait<0>();                  // this thread's share of tile t (and Q) has landed
                    fence_proxy_async();
                    mbar_arrive(&k_full[g & 1]);
                    mbar_arrive(&v_full[g & 1]);
                    if (t + 1 < nt) {
                        const int s1 = (g + 1) & 1;
                        mbar_wait(&v_empty[s1], (((g + 1) >> 1) & 1) ^ 1);       // P.V(t-1): S(t-1) retired before it
                        load_k(t + 1, s1);
                        load_v(s1);
                        cp_async_commit();
                    }
                }
            }
        }
    } else if (warp == NSW + 4) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc_s = make_idesc_bf16(TC_BQ, TC_BK);
        constexpr uint32_t idesc_o = make_idesc_bf16(TC_BQ, TC_DH) | (1u << 16);     // B operand (V) is MN-major
        const uint32_t aQ = smem_u32(sQ);
        auto issue_s = [&](int t) {              // t = running key-tile count g: stage t & 1, (PIPE) S buffer t & 1
            const uint32_t aK = smem_u32(sKV + (t & 1) * Cfg::KV_STAGE);
            const uint32_t d_s = tmem_S + (PIPE ? (uint32_t)(t & 1) * 128u : 0u);
            if (lane == 0) {
                bool first = true;
                // small terms first (X3): Ql.Kh, Qh.Kl, then Qh.Kh
#pragma unroll
                for (int term = X3 ? 0 : 2; term < 3; ++term) {
                    const uint64_t dq = make_sw128_kmajor_desc(aQ + ((X3 && term == 0) ? TC_TILE : 0));
                    const uint64_t dk = make_sw128_kmajor_desc(aK + ((X3 && term == 1) ? TC_TILE : 0));
#pragma unroll
                    for (int k = 0; k < TC_DH / 16; ++k) {
                        umma_bf16(d_s, dq + 2 * k, dk + 2 * k, idesc_s, first ? 0u : 1u);
                        first = false;
                    }
                }
                umma_commit(&k_empty[t & 1]);
                umma_commit(&s_full[PIPE ? (t & 1) : 0]);
            }
            __syncwarp();
        };
        auto issue_next_s = [&](int g, int t) {      // S of the next key tile of the same query tile
            mbar_wait(&k_full[(g + 1) & 1], ((g + 1) >> 1) & 1);
            tc_fence_after();
            issue_s(g + 1);
            if (t + 2 == nt && lane == 0) umma_commit(q_empty);     // last S of this query tile
        };
        int g = 0;                               // key tiles issued so far by this CTA (all query tiles)
        for (int it = 0; it < n_items; ++it) {
            // S of the first key tile goes out right behind the previous query tile's last P.V: S in TMEM is free (its
            // softmax has arrived on p_full), and O is only overwritten by this tile's first P.V, which waits for a
            // p_full that the softmax warps raise after they have read the previous O out
            mbar_wait(&k_full[g & 1], (g >> 1) & 1);
            tc_fence_after();
            issue_s(g);
            if (nt == 1 && lane == 0) umma_commit(q_empty);
            for (int t = 0; t < nt; ++t, ++g) {
                const int s = g & 1;
                // PIPE: S(t+1) goes out first, into the other S buffer (its last reader, the softmax of tile t-1, has
                // arrived on the p_full this warp waited for one iteration ago) -- it runs under the softmax of tile t
                if (PIPE && t + 1 < nt) issue_next_s(g, t);
                mbar_wait(p_full, g & 1);
                mbar_wait(&v_full[s], (g >> 1) & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t aV = smem_u32(sKV + s * Cfg::KV_STAGE + NP * TC_TILE);
                    bool first = t == 0;         // O accumulates across the key tiles of one query tile
#pragma unroll
                    for (int term = X3 ? 0 : 2; term < 3; ++term) {      // Pl.Vh, Ph.Vl, Ph.Vh
                        const uint32_t p_plane = tmem_P + ((X3 && term == 0) ? 64u : 0u);      // Pl, then Ph
                        const uint32_t v_plane = aV + ((X3 && term == 1) ? TC_TILE : 0);
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k) {
                            // P: 16 keys = 8 packed columns of its plane in tensor memory; V: 16 keys = 2048 B per k-step
                            const uint64_t dv = make_sw128_mnmajor_desc(v_plane + k * 2048);
                            umma_bf16_ts(tmem_O, p_plane + 8 * k, dv, idesc_o, first ? 0u : 1u);
                            first = false;
                        }
                    }
                    umma_commit(&v_empty[s]);
                    if (PIPE) umma_commit(p_free);
                    if (t == nt - 1) umma_commit(o_done);
                }
                __syncwarp();
                if (!PIPE && t + 1 < nt) issue_next_s(g, t);
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax (thread == query row == TMEM lane)
        const int r = threadIdx.x & 127;                   // query row of this thread
        const int half = threadIdx.x >> 7;                 // HS > 1: which 128 / HS key columns of the row (0 otherwise)
        constexpr int NC = 8 / HS;                         // 16-column steps per sweep of this thread
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        constexpr float kRescale = 8.0f;                   // log2 domain: P stays below 2^8
        float m_run = -INFINITY, l_run = 0.f;
        int g = 0;                                         // key tiles consumed so far by this CTA
        // One sweep over S per tile: P = exp2(S * scale - m_run) is formed with the maximum carried from the earlier
        // tiles while the tile's own maximum is tracked alongside; only when some row of the warp exceeds m_run by
        // more than 2^8 (or on the first tile, where no maximum exists yet) is the maximum raised, O rescaled in
        // TMEM and the tile's P recomputed.  (A separate maximum sweep doubled the TMEM reads and cost ~40 % of the
        // softmax warps' issue slots; ncu showed the tensor pipe waiting on them 85 % of the time.)
        // The eight 16-column TMEM loads of a sweep are software pipelined (chunk c + 1 is in flight while chunk c goes
        // through exp2 / pack / store).  On the first tile no maximum exists yet: it is seeded from the first 16 keys
        // of the row (already in registers) instead of a separate maximum sweep over S -- any value within 2^8 of the
        // true maximum is exact, and the rare larger excess takes the same raise-and-recompute path as later tiles.
        uint32_t drop_x0 = 0;                              // DROP: (query << 16) | (first key of the tile >> 1)
        const uint32_t drop_y = drop_attn_y(drop, b * (H / TC_DH) + h);
        // PIPE: the thread's P of the current tile (hi / lo planes, 16 keys per row of the arrays) waits in registers
        // until P.V of the previous tile has released the shared-memory P tile; s_buf = S buffer of the current tile
        uint32_t PH[PIPE ? NC : 1][8], PL[PIPE ? NC : 1][8];
        uint32_t s_buf = tmem_S;
        auto exp_sweep = [&](int valid, float& m_use, bool seed, float& tile_max_raw, float& sum) {
            tile_max_raw = -INFINITY;
            sum = 0.f;
            const bool full = valid == TC_BK;
            uint32_t va[16], vb[16];
            auto step = [&](uint32_t (&v)[16], int c, int li) {       // 16 keys: columns c*16 .. c*16+15; li = c - c0
                uint32_t ph[8], pl[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    float s0 = __uint_as_float(v[j]), s1 = __uint_as_float(v[j + 1]);
                    if (!full) {                 // warp-uniform: only the last key tile of a sample is ragged
                        if (c * 16 + j >= valid) s0 = -INFINITY;        // c = global 16-column step
                        if (c * 16 + j + 1 >= valid) s1 = -INFINITY;
                    }
                    tile_max_raw = fmaxf(tile_max_raw, fmaxf(s0, s1));
                    float p0 = ex2_fast(fmaf(s0, scale_log2, -m_use));
                    float p1 = ex2_fast(fmaf(s1, scale_log2, -m_use));
                    sum += p0 + p1;
                    if (DROP) {      // thr == 0 (p = 0) drops nothing; no run-time test inside the pipelined sweep
                        const uint32_t hsh = drop_hash(drop.s0, drop.s1, drop_x0 + (uint32_t)(c * 8 + (j >> 1)), drop_y);
                        if ((hsh & 0xffffu) < drop.thr) p0 = 0.f;
                        if ((hsh >> 16) < drop.thr) p1 = 0.f;
                    }
                    ph[j >> 1] = pack_bf16x2(p0, p1);
                    if (X3) pl[j >> 1] = pack_bf16x2(p0 - bf16lo(ph[j >> 1]), p1 - bf16hi(ph[j >> 1]));
                }
                if (PIPE) {                      // li is a literal at every call site: PH / PL stay in registers
#pragma unroll
                    for (int q = 0; q < 8; ++q) { PH[li][q] = ph[q]; PL[li][q] = pl[q]; }
                    return;
                }
                // -> packed columns c * 8 .. + 7 of this row's lane in the P plane(s)
                tmem_st_32x8(tmem_P + lane_addr + c * 8, ph);
                if (X3) tmem_st_32x8(tmem_P + 64 + lane_addr + c * 8, pl);
            };
            const int c0 = half * NC;                // first global 16-column step of this thread
            if (seed) {
                // the seed must be the same for all threads of a row: all read the row's first 16 keys
                if (HS > 1 && half > 0) {
                    tmem_ld_32x16(s_buf + lane_addr, vb);
                    tmem_ld_wait_on(vb);
                }
            }
            tmem_ld_32x16(s_buf + lane_addr + c0 * 16, va);
            auto pair = [&](int c) {                 // 16-column steps c and c + 1 of this thread
                tmem_ld_wait_on(va);
                if (c == 0 && seed) {
                    float m0 = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (full || j < valid)
                            m0 = fmaxf(m0, __uint_as_float((HS > 1 && half > 0) ? vb[j] : va[j]));
                    m_use = m0 * scale_log2;         // key lists are compacted: column 0 always exists
                }
                tmem_ld_32x16(s_buf + lane_addr + (c0 + c + 1) * 16, vb);
                step(va, c0 + c, c);
                tmem_ld_wait_on(vb);
                if (c + 2 < NC) tmem_ld_32x16(s_buf + lane_addr + (c0 + c + 2) * 16, va);
                step(vb, c0 + c + 1, c + 1);
            };
            if (PIPE) {
                static_assert(!PIPE || NC == 4 || NC == 2, "the PIPE form unrolls two or four 16-column steps per thread");
                pair(0);
                if (NC == 4) pair(2);
            } else {
#pragma unroll 1
                for (int c = 0; c < NC; c += 2) pair(c);
            }
        };
        // PIPE: the registers of exp_sweep -> the P tile (columns half * 64 .. + 63 of row r, hi and lo planes)
        auto store_p = [&]() {
#pragma unroll
            for (int li = 0; li < (PIPE ? NC : 0); ++li) {
                tmem_st_32x8(tmem_P + lane_addr + (half * NC + li) * 8, PH[li]);
                tmem_st_32x8(tmem_P + 64 + lane_addr + (half * NC + li) * 8, PL[li]);
            }
        };
        // HS == 2: the two threads of a row agree on the tile maximum through shared memory (slot = tile parity, so the
        // next tile's write cannot overtake a slow reader); named barrier 1 covers the 256 softmax threads
        auto joint_max = [&](float mine, int slot) {
            if (HS == 1) return mine;
            xch[(slot * HS + half) * 128 + r] = mine;
            asm volatile("bar.sync 1, %0;" ::"n"(128 * HS) : "memory");
#pragma unroll
            for (int o = 1; o < HS; ++o) mine = fmaxf(mine, xch[(slot * HS + ((half + o) & (HS - 1))) * 128 + r]);
            return mine;
        };
        for (int it = 0; it < n_items; ++it) {
        const int q0 = (it0 + it) * TC_BQ;
        m_run = -INFINITY;
        l_run = 0.f;
        for (int t = 0; t < nt; ++t, ++g) {
            const int valid = min(TC_BK, nk - t * TC_BK);          // keys of this tile that exist
            if (DROP) drop_x0 = drop_attn_x(q0 + r, t * TC_BK);
            // !PIPE: S(t) done; MMAs retire in order, so P.V(t-1) is done as well.  PIPE: S(t) was issued ahead of
            // P.V(t-1); whoever touches O or the P tile waits for p_free (phase g - 1, passes at once for g == 0)
            if (PIPE) s_buf = tmem_S + (uint32_t)(g & 1) * 128u;
            mbar_wait(&s_full[PIPE ? (g & 1) : 0], PIPE ? (g >> 1) & 1 : g & 1);
            tc_fence_after();
            bool pv_done = !PIPE;
            float mt_raw, sum;
            exp_sweep(valid, m_run, t == 0, mt_raw, sum);
            mt_raw = joint_max(mt_raw, g & 1);
            const float mt = mt_raw * scale_log2;
            const bool raise = mt > m_run + kRescale;
            if (__any_sync(0xffffffffu, raise)) {
                float corr = 1.0f;
                if (raise) {
                    corr = exp2f(m_run - mt);
                    m_run = mt;
                    l_run *= corr;
                }
                if (t > 0) {
                    if (!pv_done) {
                        mbar_wait(p_free, (g & 1) ^ 1);
                        tc_fence_after();
                        pv_done = true;
                    }
                    // rescale this warp's 32 rows of O in TMEM (rows that keep their maximum use corr == 1);
                    // HS == 2: each thread of a row takes 32 of the 64 columns
                    if (HS == 4) {               // 16 of the 64 columns per thread
                        uint32_t v[16];
                        tmem_ld_32x16(tmem_O + lane_addr + half * 16, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * corr);
                        tmem_st_32x16(tmem_O + lane_addr + half * 16, v);
                    } else {
#pragma unroll
                        for (int c = 0; c < 2 / HS; ++c) {
                            uint32_t v[32];
                            const uint32_t oc = (uint32_t)(HS == 2 ? half : c) * 32;
                            tmem_ld_32x32(tmem_O + lane_addr + oc, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * corr);
                            tmem_st_32x32(tmem_O + lane_addr + oc, v);
                        }
                    }
                    tmem_st_wait();
                }
                exp_sweep(valid, m_run, false, mt_raw, sum);      // P of this tile against the raised maximum
            }
            l_run += sum;
            if (PIPE) {
                if (!pv_done) mbar_wait(p_free, (g & 1) ^ 1);
                store_p();
            }
            tmem_st_wait();                      // P (and a rescaled O) have reached tensor memory
            tc_fence_before();                   // S reads / P, O writes ordered before the issuer's next MMAs
            mbar_arrive(p_full);
        }
        mbar_wait(o_done, it & 1);
        tc_fence_after();
        const int row = q0 + r;
        float l_row = l_run;
        if (HS > 1) {                            // row sum = the threads' partial sums (slots are free: the last
            xch[half * 128 + r] = l_run;         // joint_max of this query tile lies behind a barrier all have passed)
            asm volatile("bar.sync 1, %0;" ::"n"(128 * HS) : "memory");
#pragma unroll
            for (int o = 1; o < HS; ++o) l_row += xch[((half + o) & (HS - 1)) * 128 + r];
            asm volatile("bar.sync 1, %0;" ::"n"(128 * HS) : "memory");      // reads done before the next query tile writes
        }
        const float inv = l_row > 0.f ? (DROP ? drop.scale : 1.0f) / l_row : 0.f;
        // training step: log2-sum-exp of the row (exp2 domain of the scaled scores) in the {lse2, D} layout of the
        // backward's statistics buffer, so that t2s_attn_bwd does not have to recompute S for it
        if (DROP && lse_out && row < L && (HS == 1 || half == 0))
            lse_out[(((long long)b * (H / TC_DH) + h) * lse_rows + row) * 2] = m_run + log2f(l_row);
        __nv_bfloat16* op = out + ((long long)b * L + row) * ldo + h * TC_DH;
        constexpr int OW = HS == 4 ? 16 : 32;    // O columns per read-out chunk
#pragma unroll
        for (int cc = 0; cc < 64 / (HS * OW); ++cc) {
            const int c = HS > 1 ? half : cc;    // chunk of O this thread writes out
            uint32_t v[OW];
            if (HS == 4) tmem_ld_32x16(tmem_O + lane_addr + c * OW, reinterpret_cast<uint32_t (&)[16]>(v));
            else tmem_ld_32x32(tmem_O + lane_addr + c * OW, reinterpret_cast<uint32_t (&)[32]>(v));      // warp-collective: rows past L load too
            tmem_ld_wait();
            if (row < L) {
#pragma unroll
                for (int d = 0; d < OW; d += 8) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[d + e]) * inv;
                    uint4 hi;
                    hi.x = pack_bf16x2(f[0], f[1]); hi.y = pack_bf16x2(f[2], f[3]);
                    hi.z = pack_bf16x2(f[4], f[5]); hi.w = pack_bf16x2(f[6], f[7]);
                    *reinterpret_cast<uint4*>(op + c * OW + d) = hi;
                    if (X3) {
                        uint4 lo;
                        lo.x = pack_bf16x2(f[0] - bf16lo(hi.x), f[1] - bf16hi(hi.x));
                        lo.y = pack_bf16x2(f[2] - bf16lo(hi.y), f[3] - bf16hi(hi.y));
                        lo.z = pack_bf16x2(f[4] - bf16lo(hi.z), f[5] - bf16hi(hi.z));
                        lo.w = pack_bf16x2(f[6] - bf16lo(hi.w), f[7] - bf16hi(hi.w));
                        *reinterpret_cast<uint4*>(op + H + c * OW + d) = lo;
                    }
                }
            }
        }
        tc_fence_before();                       // O read-out ordered before the next query tile's first P.V
        }   // query tiles of this CTA
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NSW + 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <bool X3, bool DROP>
static int launch_attn_tc(const void* qkv, long long ld, int lo_off, int B, int L, int H, int heads,
                          const int* key_idx, const int* n_keys, int key_stride, void* out, long long ldo,
                          cudaStream_t st, DropCfg drop = DropCfg{0, 0, 0, 0, 1.f}, float* lse_out = nullptr,
                          int lse_rows = 0) {
    using Cfg = TcCfg<X3>;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<X3, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) { set_error("attn_tc attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr = true;
    }
    const int n_qt = (L + TC_BQ - 1) / TC_BQ;
    dim3 grid((n_qt + TC_NQ - 1) / TC_NQ, heads, B);
    attn_tc_kernel<X3, DROP><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv), ld, lo_off, L, H, key_idx, n_keys, key_stride,
        reinterpret_cast<__nv_bfloat16*>(out), ldo, 0.125f * 1.4426950408889634f, drop, lse_out, lse_rows);
    return launch_status("attn_tc");
}

}  // namespace t2s

using namespace t2s;

static int attn_tc_entry(const void* qkv, long long ld, int lo_off, int B, int L, int H, int heads,
                         const int* key_idx, const int* n_keys, int key_stride, void* out, long long ldo,
                         void* stream, bool train, DropCfg drop, float* lse_out = nullptr, int lse_rows = 0) {
    if (H != heads * TC_DH || (ld % 8) || (ldo % 8) || (lo_off % 8) || B <= 0 || L <= 0) {
        set_error("attn_tc: head size must be 64 and pitches multiples of 8 (H %d heads %d ld %lld ldo %lld)", H, heads, ld, ldo);
        return T2S_ERR_SHAPE;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (lo_off > 0) {
        if (lo_off < 3 * H || ld < lo_off + 3 * H || ldo < 2LL * H) { set_error("attn_tc: bad hi|lo layout"); return T2S_ERR_SHAPE; }
        if (train) return launch_attn_tc<true, true>(qkv, ld, lo_off, B, L, H, heads, key_idx, n_keys, key_stride, out, ldo, st, drop, lse_out, lse_rows);
        return launch_attn_tc<true, false>(qkv, ld, lo_off, B, L, H, heads, key_idx, n_keys, key_stride, out, ldo, st);
    }
    if (train) return launch_attn_tc<false, true>(qkv, ld, 0, B, L, H, heads, key_idx, n_keys, key_stride, out, ldo, st, drop, lse_out, lse_rows);
    return launch_attn_tc<false, false>(qkv, ld, 0, B, L, H, heads, key_idx, n_keys, key_stride, out, ldo, st);
}

extern "C" int t2s_attn_tc(const void* qkv, long long ld, int lo_off, int B, int L, int H, int heads,
                           const int* key_idx, const int* n_keys, int key_stride, void* out, long long ldo,
                           void* stream) {
    return attn_tc_entry(qkv, ld, lo_off, B, L, H, heads, key_idx, n_keys, key_stride, out, ldo, stream, false,
                         DropCfg{0, 0, 0, 0, 1.f});
}

/* t2s_attn_tc of the training step: attention_probs dropout (p may be 0) whose mask t2s_attn_bwd_dropout recomputes
 * from (seed, site) -- the query rows are positions 0..L-1 of its virtual sequence -- and, when lse_out != null, the
 * rows' log2-sum-exp written to lse_out[((b * heads + h) * lse_rows + row) * 2] (the backward's statistics layout) */
extern "C" int t2s_attn_tc_dropout(const void* qkv, long long ld, int lo_off, int B, int L, int H, int heads,
                                   const int* key_idx, const int* n_keys, int key_stride, void* out, long long ldo,
                                   float p, unsigned long long seed, unsigned site, float* lse_out, int lse_rows,
                                   void* stream) {
    if (p < 0.f || p >= 1.f || L > 65535 || B * heads >= (1 << 20) || (lse_out && lse_rows < L)) {
        set_error("attn_tc_dropout: p in [0, 1), L < 65536, lse_rows >= L");
        return T2S_ERR_ARG;
    }
    return attn_tc_entry(qkv, ld, lo_off, B, L, H, heads, key_idx, n_keys, key_stride, out, ldo, stream, true,
                         make_drop(p, seed, site), lse_out, lse_rows);
}
