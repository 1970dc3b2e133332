// K1: bf16 GEMM on the 5th-gen tensor cores.  C[M,N] = epi(A[M,K] . W[N,K]^T + bias)
//
// Replaces every `addmm` the reference issues for the dense layers of the
// fusion transformer (BertSelfAttention.query/key/value fused to N=2304,
// BertSelfOutput.dense, BertIntermediate.dense + erf-GELU, BertOutput.dense,
// the classifier `nn.Linear(768, V)` of reference pythia/modules/layers.py:101,
// OcrPtrNet.query/key of pythia/models/t2s.py:645-646) -- SURVEY 2.3.
//
// Design (one CTA per SM, persistent over output tiles, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads of a 128x64 A tile and a
//               BNx64 W tile (both K-major, 128B swizzle) into a STAGES-deep smem ring,
//               completion on `full[s]` mbarriers.
//   warp 1      MMA issuer: one elected lane issues 4 x tcgen05.mma (M=128, N=BN, K=16)
//               per k-block into a TMEM accumulator; tcgen05.commit releases the smem
//               slot (`empty[s]`) and, after the last k-block, publishes the accumulator
//               (`tfull[a]`).  Two accumulators (2 x BN TMEM columns) so the epilogue of
//               tile i overlaps the main loop of tile i+1.
//   warps 2..9  epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> bias /
//               erf-GELU / residual -> bf16, fp32 or bf16 hi|lo -> per-warp 128B-swizzled
//               shared-memory box (32 rows x 128 B) -> TMA store (full-line writes, clipped
//               to the matrix bounds by the tensor map; a lane-per-row register layout would
//               otherwise scatter 16-byte stores over 32 rows per instruction).
// Tiles are walked n-fastest so the CTAs running concurrently share one A row-block
// through L2 and A streams from HBM once.
// The throughput launches (BN = 256, M >= 256) run as 2-CTA clusters: by default ONE tcgen05.mma.cta_group::2 of
// M = 256 per pair, issued by the leader, each CTA holding its A tile and half of the W tile (PAIR == 2 below); their
// + bf16 residual epilogue gets the residual tile by TMA into the staging box and overwrites it in place (RT below).
#include "common.cuh"
#include <stdlib.h>
#include <mutex>
#include "../../include/t2s_b200.h"

namespace t2s {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;          // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_THREADS = 320;    // 10 warps
constexpr int GEMM_EPI_WARPS = 8;
// 1: two staging boxes per epilogue warp in the cta_group::2 form (and five stages instead of six).  Measured on one box,
// default build vs -DT2S_GEMM_DB_STAGING=1: every fusion shape 0-4 % slower, the epilogue-heavy ones (attn_out + residual,
// ffn_up + GELU) unchanged -- their epilogues do not wait for the store to read its box.  Off.
#ifndef T2S_GEMM_DB_STAGING
#define T2S_GEMM_DB_STAGING 0
#endif
#ifndef T2S_GEMM_RES_TMA
#define T2S_GEMM_RES_TMA 1
#endif

struct GemmEpi {
    void* C;
    const float* bias;
    const void* residual;
    long long ldc, ldr;
    int flags;
    int c_lo_off;   // T2S_GEMM_OUT_SPLIT: column offset of the `lo` half of the bf16 hi|lo output
    int stages;     // BN == 64 only: pipeline depth of this launch (GemmCfg<64>::STAGES or DEEP_STAGES)
};

// BN = 64 is the latency tile of the greedy decode (M = batch rows, a few dozen CTAs per launch, run next to the capped
// throughput GEMMs of the other stream).  A clock trace of one decode launch (M = 64, K = 768) showed the kernel spending
// its time on round trips, not on data: with 3 stages a k-block arrived every ~700 cycles (three loads per ~2000-cycle
// TMA latency) and half of every A stage was zero fill.  So this tile is 64 rows x 64 columns: the A box is 64 rows
// (8 KB; the M = 128 MMA reads the W tile behind it as rows 64..127, whose accumulator rows are never read), a stage is
// 16 KB and six of them fit next to a second CTA on the same SM (105 KB shared memory, 192 threads, 128 TMEM columns
// each).  Epilogue: the two warps that own TMEM lane quarters 0 and 1.
template <int BN>
struct GemmCfg {
    static constexpr int BM = BN == 64 ? 64 : GEMM_BM;          // rows of an output tile (and of the A box)
    static constexpr int A_BYTES = BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int TMEM_COLS = 2 * BN;     // 512 / 256 / 128: powers of two >= 32
    static constexpr int EPI_WARPS = BN == 64 ? 4 : GEMM_EPI_WARPS;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static constexpr int MIN_CTAS = BN == 64 ? 2 : 1;
    static constexpr int ROW_WARPS = BM / 32;                 // epilogue warps per column group that own tile rows
    static constexpr int STAGING_BYTES = (BN == 64 ? 2 : EPI_WARPS) * 4096;    // one 32-row x 128-byte box per active epilogue warp
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    // BN == 64 launches that cannot fill the SMs left to them twice over (N = 768: 12 CTAs) give up the second resident
    // CTA for a pipeline that holds 12 k-blocks: a K = 768 product is then ONE round of loads and K = 3072 four, where
    // six stages took two and eight -- the decode chain is bound by exactly these round trips (profiles/r2_tail_kernels.md)
    // PAIR == 2 (cta_group::2): a stage holds this CTA's A tile and HALF of the W tile -- 32 KB, six stages
    static constexpr int STAGE2_BYTES = A_BYTES + B_BYTES / 2;
    // ... or five with TWO staging boxes per epilogue warp (T2S_GEMM_DB_STAGING): the epilogue then writes box i + 1 while
    // the TMA store of box i still reads its shared memory, instead of waiting for it
    // The + bf16 residual epilogue of the pair form (MODE == GEMM_MODE_RES, "RT") always runs with two boxes: the residual
    // tile is brought into the box by TMA one 64-column group ahead and the result is written over it in place (below).
    __host__ __device__ static constexpr bool rt(int mode, int pair) { return pair == 2 && mode == 32 && T2S_GEMM_RES_TMA; }
    __host__ __device__ static constexpr bool db(int mode, int pair) { return pair == 2 && (T2S_GEMM_DB_STAGING || rt(mode, pair)); }
    __host__ __device__ static constexpr int stages2(int mode) { return db(mode, 2) ? 5 : 6; }
    __host__ __device__ static constexpr int staging2(int mode) { return (db(mode, 2) ? 2 : 1) * STAGING_BYTES; }
    __host__ __device__ static constexpr int smem2(int mode) { return stages2(mode) * STAGE2_BYTES + staging2(mode) + 1024 + 256; }
    static constexpr int DEEP_STAGES = 12;
    static constexpr int DEEP_SMEM_BYTES = DEEP_STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 512;
};

// MODE >= 0: the epilogue flags (low 5 bits of ep.flags) and "has a residual operand" (bit 5) are compile-time
// constants -- the hot epilogues of the fusion transformer get their own lean instantiation; MODE < 0: read at run time.
constexpr int GEMM_MODE_RES = 32;
// PAIR == 1: the CTAs of a 2-CTA cluster work on vertically adjacent tiles (rows 2p and 2p + 1 of the same n block) in step.
// Each loads its own A tile and HALF of the shared W tile, multicast into both CTAs' shared memory, so a k-block costs
// each SM 32 KB of L2 reads instead of 48 KB; a stage is refilled only when BOTH tensor cores have released it (the
// MMA commits are multicast to the `empty` barriers of both CTAs).
// PAIR == 2: the same pair of tiles as ONE tcgen05.mma.cta_group::2 of M = 256 issued by the leader CTA.  Each CTA keeps
// only its own half of the W tile in shared memory (the tensor cores exchange the halves), so a k-block costs each SM
// 32 KB of shared-memory fill and 8 KB instead of 12 KB of operand reads per K = 16 step: the single-CTA form asks its
// shared memory for 96 + 96 B per clock (TMA fill + operand reads), the pair form for 64 + 64 of the 128 there are.
template <int BN, int MODE, int PAIR>
__global__ void __launch_bounds__(GemmCfg<BN>::THREADS, GemmCfg<BN>::MIN_CTAS)     // 10 warps = 3 on two of the four 16 K-register partitions: 168 registers at most
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, GemmEpi ep,
                         int M, int N, int K, int k_lo_off) {
    using Cfg = GemmCfg<BN>;
    constexpr bool TWO = PAIR == 2;
    constexpr int STAGE_BYTES = TWO ? Cfg::STAGE2_BYTES : Cfg::STAGE_BYTES;
    constexpr bool RT = Cfg::rt(MODE, PAIR);          // bf16 residual by TMA into the staging box, result written in place
    constexpr bool DB = Cfg::db(MODE, PAIR);          // two staging boxes per epilogue warp
    const int STAGES = BN == 64 ? ep.stages : (TWO ? Cfg::stages2(MODE) : Cfg::STAGES);      // run-time depth for the latency tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + STAGES * STAGE_BYTES;           // 1024-byte aligned (STAGE_BYTES is a multiple of 1024)
    uint64_t* full = reinterpret_cast<uint64_t*>(staging + (TWO ? Cfg::staging2(MODE) : Cfg::STAGING_BYTES));
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* rbar = tempty + 2;           // RT: [epilogue warp][box]: the residual tile of a 64-column group has landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + (RT ? 2 * Cfg::EPI_WARPS : 0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_n = (N + BN - 1) / BN;
    const int num_m = (M + Cfg::BM - 1) / Cfg::BM;
    // PAIR: "tiles" counts pair tiles (two row blocks x one n block); work item w of this CTA is pair tile
    // first_tile + w * tile_step, of which it takes row block 2 * (pair tile / num_n) + rank
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    const int tiles = (PAIR ? (num_m + 1) / 2 : num_m) * num_n;
    const int first_tile = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    // k_lo_off > 0: "bf16x3" mode.  A and W hold fp32 values split as bf16 hi|lo (lo at column k_lo_off) and
    // the k loop runs three segments -- hi.hi, hi.lo, lo.hi -- into the same fp32 accumulator (the lo.lo
    // term is below 2^-16 relative and is dropped), giving fp32-class products on the bf16 tensor pipe.
    const int kseg = (K + GEMM_BK - 1) / GEMM_BK;
    const int kblocks = k_lo_off > 0 ? 3 * kseg : kseg;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        if (RT) tma_prefetch_desc(&tmR);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], PAIR == 1 ? 2 : 1);
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(&tfull[a], 1);
                mbar_init(&tempty[a], TWO ? 2 * Cfg::EPI_WARPS : Cfg::EPI_WARPS);       // TWO: the leader's counts both CTAs' epilogues
            }
            if (RT)
                for (int i = 0; i < 2 * Cfg::EPI_WARPS; ++i) mbar_init(&rbar[i], 1);
            fence_barrier_init();
        }
        __syncwarp();
        if (TWO) {
            tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
            tmem_relinquish_2sm();
        } else {
            tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();        // the peer's barriers are initialised before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // launched through launch_pdl() (decode tile): everything above ran next to the tail of the preceding kernel
    pdl_wait();
    pdl_release();

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = first_tile; tile < tiles; tile += tile_step) {
            const int m_blk = PAIR ? 2 * (tile / num_n) + rank : tile / num_n, n_blk = tile % num_n;
            for (int kb = 0; kb < kblocks; ++kb) {
                if (lane == 0) {
                    const int seg = kb / kseg, kk = (kb - seg * kseg) * GEMM_BK;
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    if (TWO) {
                        // both CTAs' bytes are counted on the LEADER's barrier; each CTA fills only its own shared memory
                        const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
                        if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * STAGE_BYTES);
                        tma_load_2d_2sm(sa, &tmA, lead_full, kk + (seg == 2 ? k_lo_off : 0), m_blk * Cfg::BM);
                        tma_load_2d_2sm(sa + Cfg::A_BYTES, &tmB, lead_full, kk + (seg == 1 ? k_lo_off : 0),
                                        n_blk * BN + rank * (BN / 2));
                    } else {
                        mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                        tma_load_2d(sa, &tmA, &full[stage], kk + (seg == 2 ? k_lo_off : 0), m_blk * Cfg::BM);
                        if (PAIR)       // this CTA's half of the W tile (tmB box = BN / 2 rows), to both CTAs
                            tma_load_2d_multicast(sa + Cfg::A_BYTES + rank * (Cfg::B_BYTES / 2), &tmB, &full[stage],
                                                  kk + (seg == 1 ? k_lo_off : 0), n_blk * BN + rank * (BN / 2), 3);
                        else
                            tma_load_2d(sa + Cfg::A_BYTES, &tmB, &full[stage], kk + (seg == 1 ? k_lo_off : 0), n_blk * BN);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (TWO: the leader CTA issues for the pair)
        constexpr uint32_t idesc = make_idesc_bf16(TWO ? 2 * GEMM_BM : GEMM_BM, BN);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int tile = first_tile; tile < tiles && !(TWO && rank != 0); tile += tile_step) {
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t da = make_sw128_kmajor_desc(sa);
                    const uint64_t db = make_sw128_kmajor_desc(sa + Cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        // +32 B per K=16 step inside the swizzle atom == +2 in the (addr >> 4) field
                        if (TWO) umma_bf16_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        else umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    if (TWO) umma_commit_2sm(&empty[stage], 3);
                    else if (PAIR) umma_commit_multicast(&empty[stage], 3);
                    else umma_commit(&empty[stage]);
                    if (kb == kblocks - 1) {
                        if (TWO) umma_commit_2sm(&tfull[acc], 3); else umma_commit(&tfull[acc]);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else {
        // ------------------------------------------------ epilogue (8 warps)
        const int ew = warp - 2;
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        // BN >= 128: two warps per lane quarter split the columns; BN == 64: four warps take all 64 columns
        constexpr int COLS_PER_WARP = BN / (Cfg::EPI_WARPS / 4);
        const bool active = quarter < Cfg::ROW_WARPS;       // BN == 64: TMEM lanes 64..127 hold no tile rows
        const int half = ew >> 2;
        const int fl = MODE >= 0 ? (MODE & 31) : ep.flags;
        const bool has_res = MODE >= 0 ? (MODE & GEMM_MODE_RES) != 0 : ep.residual != nullptr;
        const bool gelu = fl & T2S_GEMM_GELU;
        const bool out_f32 = fl & T2S_GEMM_OUT_F32;
        const bool res_f32 = fl & T2S_GEMM_RES_F32;
        const bool out_split = fl & T2S_GEMM_OUT_SPLIT;
        const bool dgelu = fl & T2S_GEMM_DGELU;
        // this warp's staging box: 32 rows x 128 B, 128B swizzle.  DB: two boxes used in turn -- the wait before a box is
        // rewritten then covers the store issued two stores ago
        uint8_t* const box0 = staging + (BN == 64 ? quarter & 1 : (DB ? 2 * ew : ew)) * 4096;
        uint8_t* box = box0;
        uint8_t* my_row = box + lane * 128;
        auto next_box = [&]() {                  // after every committed store (warp-uniform)
            if (DB) {
                box = box == box0 ? box0 + 4096 : box0;
                my_row = box + lane * 128;
            }
        };
        auto box_wait = [&]() {                  // the store that last used `box` has read it
            if (lane == 0) { if (DB) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
            __syncwarp();
        };
        const int sw = lane & 7;
        int acc = 0;
        uint32_t acc_phase = 0;
        // RT: 64-column groups handled so far by this warp; group u uses box u & 1, barrier phase (u >> 1) & 1.  The
        // residual tile of group u + 1 is requested while group u is processed (the box it goes to was last read by the
        // store of group u - 1), the first group of the first tile before the loop.
        int ruse = 0;
        auto res_issue = [&](int use, int col, int row_first) {
            if (lane == 0) {
                uint64_t* b = &rbar[2 * ew + (use & 1)];
                mbar_arrive_expect_tx(b, 4096);
                tma_load_2d(box0 + (use & 1) * 4096, &tmR, b, col, row_first);
            }
        };
        auto group_origin = [&](int t, int& col, int& row_first) {       // first group of tile t for this warp
            const int mb = PAIR ? 2 * (t / num_n) + rank : t / num_n, nb = t % num_n;
            col = nb * BN + (ew >> 2) * COLS_PER_WARP;
            row_first = mb * Cfg::BM + quarter * 32;
        };
        if (RT && first_tile < tiles) {
            int col, row_first;
            group_origin(first_tile, col, row_first);
            res_issue(0, col, row_first);
        }
        for (int tile = first_tile; tile < tiles; tile += tile_step) {
            const int m_blk = PAIR ? 2 * (tile / num_n) + rank : tile / num_n, n_blk = tile % num_n;
            // the bias slice of this warp (first touched here by the whole grid at once) comes in under the main loop
            if (ep.bias && lane * 32 < COLS_PER_WARP && n_blk * BN + (ew >> 2) * COLS_PER_WARP + lane * 32 < N)
                prefetch_l1(ep.bias + n_blk * BN + (ew >> 2) * COLS_PER_WARP + lane * 32);
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const int row0 = m_blk * Cfg::BM + quarter * 32;
            const int row = row0 + lane;
            const bool row_ok = row < M;
            if (active) {
                uint32_t lo_keep[16];            // OUT_SPLIT: lo half of the first chunk of a box
                // Software pipeline over the 32-column chunks: the tcgen05.ld of chunk i+1 (and the load of its
                // bf16 residual / GELU' operand) is in flight while chunk i goes through bias / GELU / pack / store,
                // and the accumulator is handed back to the MMA warp as soon as its last chunk sits in registers.
                constexpr int NCH = COLS_PER_WARP / 32;
                const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + half * COLS_PER_WARP;
                const int col_w = n_blk * BN + half * COLS_PER_WARP;
                const bool pre_res = !RT && has_res && !res_f32 && row_ok;
                const __nv_bfloat16* res16 =
                    reinterpret_cast<const __nv_bfloat16*>(ep.residual) + (long long)row * ep.ldr + col_w;
                uint32_t rbuf[2][32];
                uint4 qbuf[2][4];
                tmem_ld_32x32(t_acc, rbuf[0]);
                if (pre_res && col_w + 32 <= N) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) qbuf[0][j] = *reinterpret_cast<const uint4*>(res16 + 8 * j);
                }
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci) {
                    const int c0 = ci * 32;
                    uint32_t (&r)[32] = rbuf[ci & 1];
                    const uint4 (&q)[4] = qbuf[ci & 1];
                    tmem_ld_wait_on(r);
                    if (ci + 1 == NCH) {             // the accumulator is in registers: hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (TWO) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0)); else mbar_arrive(&tempty[acc]);
                        }
                    }
                    const int col0 = col_w + c0;
                    if (col0 >= N) continue;     // warp-uniform: the whole 32-column chunk is outside the matrix
                    const bool full32 = col0 + 32 <= N;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (ep.bias) {
                        if (full32) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
                                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (col0 + j < N) v[j] += __ldg(ep.bias + col0 + j);
                        }
                    }
                    if (gelu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                    }
                    if (RT) {
                        // The group's residual tile sits in this warp's box (same swizzled positions the result is written
                        // to below).  Warp-uniform, whatever rows of the tile exist: at the first chunk of a group wait
                        // for it, at the second request the next group's tile into the other box.
                        if (((c0 >> 5) & 1) == 0) {
                            mbar_wait(&rbar[2 * ew + (ruse & 1)], (ruse >> 1) & 1);
                        } else {
                            int ncol = col0 + 32, nrow = row0;
                            bool have = c0 + 32 < COLS_PER_WARP;
                            if (!have && tile + tile_step < tiles) {
                                group_origin(tile + tile_step, ncol, nrow);
                                have = true;
                            }
                            if (have) {
                                if (lane == 0) tma_store_wait_read<0>();     // the other box: its store has read it
                                __syncwarp();
                                res_issue(ruse + 1, ncol, nrow);
                            }
                        }
                    }
                    if (has_res && row_ok) {
                        // residual add, or (T2S_GEMM_DGELU) multiply by GELU'(aux) of the saved pre-activation
                        auto comb = [dgelu](float acc_v, float aux) { return dgelu ? acc_v * gelu_grad(aux) : acc_v + aux; };
                        if (res_f32) {
                            const float* rp = reinterpret_cast<const float*>(ep.residual) + (long long)row * ep.ldr + col0;
                            if (full32) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    const float4 a = *reinterpret_cast<const float4*>(rp + j);
                                    v[j] = comb(v[j], a.x); v[j + 1] = comb(v[j + 1], a.y);
                                    v[j + 2] = comb(v[j + 2], a.z); v[j + 3] = comb(v[j + 3], a.w);
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) if (col0 + j < N) v[j] = comb(v[j], rp[j]);
                            }
                        } else {
                            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(ep.residual) + (long long)row * ep.ldr + col0;
                            if (RT) {
                                const int hp = (c0 >> 5) & 1;
#pragma unroll
                                for (int j = 0; j < 32; j += 8) {
                                    const uint4 a = *reinterpret_cast<const uint4*>(my_row + (((hp * 4 + (j >> 3)) ^ sw) << 4));
                                    v[j] = comb(v[j], bf16lo(a.x)); v[j + 1] = comb(v[j + 1], bf16hi(a.x));
                                    v[j + 2] = comb(v[j + 2], bf16lo(a.y)); v[j + 3] = comb(v[j + 3], bf16hi(a.y));
                                    v[j + 4] = comb(v[j + 4], bf16lo(a.z)); v[j + 5] = comb(v[j + 5], bf16hi(a.z));
                                    v[j + 6] = comb(v[j + 6], bf16lo(a.w)); v[j + 7] = comb(v[j + 7], bf16hi(a.w));
                                }
                            } else if (full32) {
#pragma unroll
                                for (int j = 0; j < 32; j += 8) {
                                    const uint4 a = q[j >> 3];       // prefetched one chunk ahead
                                    v[j] = comb(v[j], bf16lo(a.x)); v[j + 1] = comb(v[j + 1], bf16hi(a.x));
                                    v[j + 2] = comb(v[j + 2], bf16lo(a.y)); v[j + 3] = comb(v[j + 3], bf16hi(a.y));
                                    v[j + 4] = comb(v[j + 4], bf16lo(a.z)); v[j + 5] = comb(v[j + 5], bf16hi(a.z));
                                    v[j + 6] = comb(v[j + 6], bf16lo(a.w)); v[j + 7] = comb(v[j + 7], bf16hi(a.w));
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) if (col0 + j < N) v[j] = comb(v[j], __bfloat162float(rp[j]));
                            }
                        }
                    }
                    if (ci + 1 < NCH) {
                        // chunk i + 1: TMEM load and bf16 residual load fly while this chunk is packed and stored
                        tmem_ld_32x32(t_acc + c0 + 32, rbuf[(ci + 1) & 1]);
                        if (pre_res && col0 + 64 <= N) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                qbuf[(ci + 1) & 1][j] = *reinterpret_cast<const uint4*>(res16 + c0 + 32 + 8 * j);
                        }
                    }
                    if (out_f32) {
                        // one box = 32 rows x 32 fp32 columns; the previous store must have read the box
                        box_wait();
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            *reinterpret_cast<float4*>(my_row + ((c ^ sw) << 4)) =
                                make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmC, box, col0, row0);
                            tma_store_commit();
                        }
                        next_box();
                    } else {
                        // one box = 32 rows x 64 bf16 columns = two 32-column chunks
                        const int hpos = (c0 >> 5) & 1;          // which half of the box this chunk fills
                        uint32_t hi[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) hi[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
                        if (hpos == 0 && !RT) box_wait();       // RT: the box already holds the residual tile
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            *reinterpret_cast<uint4*>(my_row + (((hpos * 4 + c) ^ sw) << 4)) =
                                make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                        const bool last_chunk = hpos == 1 || c0 + 32 >= COLS_PER_WARP || col0 + 32 >= N;
                        if (last_chunk) {
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_2d(&tmC, box, col0 - hpos * 32, row0);
                                tma_store_commit();
                            }
                            next_box();
                            if (RT) ++ruse;
                        }
                        if (out_split) {
                            // lo = bf16(v - hi), stored ep.c_lo_off columns to the right; 32-column boxes reuse the
                            // staging box after the hi store has read it (N % 64 == 0 is required by the host)
                            uint32_t lo[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                lo[j] = pack_bf16x2(v[2 * j] - bf16lo(hi[j]), v[2 * j + 1] - bf16hi(hi[j]));
                            if (hpos == 0) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) lo_keep[j] = lo[j];
                            } else {
                                box_wait();
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    *reinterpret_cast<uint4*>(my_row + ((c ^ sw) << 4)) =
                                        make_uint4(lo_keep[4 * c], lo_keep[4 * c + 1], lo_keep[4 * c + 2], lo_keep[4 * c + 3]);
                                    *reinterpret_cast<uint4*>(my_row + (((4 + c) ^ sw) << 4)) =
                                        make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
                                }
                                fence_proxy_async();
                                __syncwarp();
                                if (lane == 0) {
                                    tma_store_2d(&tmC, box, ep.c_lo_off + col0 - 32, row0);
                                    tma_store_commit();
                                }
                                next_box();
                            }
                        }
                    }
                }
            }
            if (!active) {                           // BN == 64: the warps of lane quarters 2 and 3
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (lane == 0) tma_store_wait_all<0>();      // global writes complete before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    // PAIR: every MMA commit of this CTA has been delivered once its last accumulator was published, and every multicast
    // load aimed at it has been consumed; the peer may still be behind, so neither leaves before the other is done
    if (PAIR) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        if (TWO) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Descriptor cache: a tensor map is a pure function of (type, base pointer, rows, cols, pitch, box), and the
// forward enqueues the same ~150 GEMMs over the same workspace / weight buffers every step, so the encoded maps are kept
// in a small direct-mapped table instead of calling cuTensorMapEncodeTiled (a driver call, ~1 us) three times per launch.
// A map holds no reference to the memory: a recycled pointer with the same geometry encodes to the same 128 bytes, so
// stale entries are harmless.  One mutex: the library serves one host thread per GPU, contention is nil.
struct TmapKey {
    const void* ptr;
    long long rows, cols, ld;
    int box_cols, box_rows, f32, dev;
    bool operator==(const TmapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols &&
               box_rows == o.box_rows && f32 == o.f32 && dev == o.dev;
    }
};
struct TmapSlot { TmapKey key; CUtensorMap map; bool used; };
constexpr int kTmapSlots = 4096;
static TmapSlot* g_tmap_cache = nullptr;
static std::mutex g_tmap_mutex;
static long long g_tmap_hits = 0, g_tmap_misses = 0;

static inline unsigned tmap_hash(const TmapKey& k) {
    unsigned long long h = reinterpret_cast<uintptr_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    h ^= (unsigned long long)k.rows * 0xC2B2AE3D27D4EB4Full + (unsigned long long)k.cols * 0x165667B19E3779F9ull;
    h ^= (unsigned long long)k.ld * 0x27D4EB2F165667C5ull + (unsigned)(k.box_rows * 131 + k.box_cols * 7 + k.f32 + k.dev * 1009);
    h ^= h >> 29;
    return (unsigned)(h * 0xBF58476D1CE4E5B9ull >> 40) % kTmapSlots;
}

static int encode_tmap_2d(CUtensorMap* tm, bool f32, const void* ptr, long long rows, long long cols, long long ld,
                          int box_cols, int box_rows);

// 2D row-major [rows, cols] tensor with row pitch `ld` elements; box = box_rows x box_cols (box_cols * esize == 128 B),
// 128B swizzle.
static int make_tmap_2d(CUtensorMap* tm, bool f32, const void* ptr, long long rows, long long cols, long long ld,
                        int box_cols, int box_rows) {
    int dev = 0;
    cudaGetDevice(&dev);
    const TmapKey key{ptr, rows, cols, ld, box_cols, box_rows, f32 ? 1 : 0, dev};
    const unsigned slot = tmap_hash(key);
    {
        std::lock_guard<std::mutex> lock(g_tmap_mutex);
        if (!g_tmap_cache) g_tmap_cache = new TmapSlot[kTmapSlots]();
        TmapSlot& s = g_tmap_cache[slot];
        if (s.used && s.key == key) {
            *tm = s.map;
            ++g_tmap_hits;
            return T2S_OK;
        }
    }
    const int rc = encode_tmap_2d(tm, f32, ptr, rows, cols, ld, box_cols, box_rows);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    TmapSlot& s = g_tmap_cache[slot];
    s.key = key;
    s.map = *tm;
    s.used = true;
    ++g_tmap_misses;
    return T2S_OK;
}

static int encode_tmap_2d(CUtensorMap* tm, bool f32, const void* ptr, long long rows, long long cols, long long ld,
                          int box_cols, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return T2S_ERR_DRIVER;
    }
    const int es = f32 ? 4 : 2;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                    const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed: CUresult %d (ptr %p rows %lld cols %lld ld %lld)", (int)r, ptr, rows,
                  cols, ld);
        return T2S_ERR_DRIVER;
    }
    return T2S_OK;
}
static int make_tmap_bf16(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
    return make_tmap_2d(tm, false, ptr, rows, cols, ld, GEMM_BK, box_rows);
}

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int BN, int MODE, int PAIR>
static int launch_gemm_mode(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tr,
                            const GemmEpi& ep, int M, int N, int K, int k_lo_off, int sm_cap, cudaStream_t st) {
    using Cfg = GemmCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, MODE, PAIR>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             BN == 64 ? Cfg::DEEP_SMEM_BYTES : (PAIR == 2 ? Cfg::smem2(MODE) : Cfg::SMEM_BYTES));
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gemm BN=%d): %s", BN, cudaGetErrorString(e));
            return (int)e;
        }
        attr_set = true;
    }
    const int tiles = ((M + Cfg::BM - 1) / Cfg::BM) * ((N + BN - 1) / BN);
    // T2S_GEMM_SM_CAP: the persistent grid leaves SMs free for latency-bound work on another stream
    const int sms = ((sm_cap > 0 && sm_cap < num_sms()) ? sm_cap : num_sms()) * Cfg::MIN_CTAS;
    if (PAIR) {
        const int pair_tiles = (((M + GEMM_BM - 1) / GEMM_BM + 1) / 2) * ((N + BN - 1) / BN);
        const int pairs = pair_tiles < sms / 2 ? pair_tiles : sms / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(Cfg::THREADS);
        cfg.dynamicSmemBytes = PAIR == 2 ? Cfg::smem2(MODE) : Cfg::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<BN, MODE, PAIR>, ta, tb, tc, tr, ep, M, N, K, k_lo_off);
        if (e != cudaSuccess) {
            set_error("gemm_bf16_tcgen05 (pair): %s", cudaGetErrorString(e));
            return (int)e;
        }
        return launch_status("gemm_bf16_tcgen05 (pair)");
    }
    const int grid = tiles < sms ? tiles : sms;
    if (BN == 64) {
        static int deep_tiles = -1;          // T2S_GEMM_DEEP_TILES: largest tile count that takes the deep pipeline (0 = never)
        if (deep_tiles < 0) {
            const char* e = getenv("T2S_GEMM_DEEP_TILES");
            deep_tiles = e ? atoi(e) : 24;
        }
        GemmEpi ep64 = ep;
        const bool deep = tiles <= deep_tiles && (K + GEMM_BK - 1) / GEMM_BK > Cfg::STAGES;
        ep64.stages = deep ? Cfg::DEEP_STAGES : Cfg::STAGES;
        launch_pdl(true, gemm_bf16_tcgen05_kernel<BN, MODE, PAIR>, dim3(grid), dim3(Cfg::THREADS),
                   deep ? Cfg::DEEP_SMEM_BYTES : Cfg::SMEM_BYTES, st, ta, tb, tc, tr, ep64, M, N, K, k_lo_off);
        return launch_status("gemm_bf16_tcgen05");
    }
    gemm_bf16_tcgen05_kernel<BN, MODE, PAIR><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(ta, tb, tc, tr, ep, M, N, K, k_lo_off);
    return launch_status("gemm_bf16_tcgen05");
}

// The epilogues of the eval forward (qkv / ptr-net plain, +residual, GELU, fp32 out, fp32 out + fp32 residual,
// GELU + hi|lo out, hi|lo out) and the GELU' dgrad of the training step are compiled with constant flags for the
// throughput tiles; everything else (BN = 64 decode tiles, rare combinations) takes the run-time-flag instantiation.
template <int BN, int PAIR>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tr,
                       const GemmEpi& ep, int M, int N, int K, int k_lo_off, int sm_cap, cudaStream_t st) {
    if (BN >= 128) {
        int mode = (ep.flags & 31) | (ep.residual ? GEMM_MODE_RES : 0);
        // the TMA-fed residual epilogue works on whole 64-column groups of whole 256-column blocks
        if (GemmCfg<BN>::rt(mode, PAIR) && (N % 256)) mode = -1;
#define T2S_GEMM_MODE_CASE(m) \
        case (m): return launch_gemm_mode<(BN >= 128 ? BN : 128), (m), PAIR>(ta, tb, tc, tr, ep, M, N, K, k_lo_off, sm_cap, st);
        switch (mode) {
            T2S_GEMM_MODE_CASE(0)
            T2S_GEMM_MODE_CASE(GEMM_MODE_RES)
            T2S_GEMM_MODE_CASE(T2S_GEMM_GELU)
            T2S_GEMM_MODE_CASE(T2S_GEMM_OUT_F32)
            T2S_GEMM_MODE_CASE(T2S_GEMM_OUT_F32 | T2S_GEMM_RES_F32 | GEMM_MODE_RES)
            T2S_GEMM_MODE_CASE(T2S_GEMM_GELU | T2S_GEMM_OUT_SPLIT)
            T2S_GEMM_MODE_CASE(T2S_GEMM_OUT_SPLIT)
            T2S_GEMM_MODE_CASE(T2S_GEMM_DGELU | GEMM_MODE_RES)
            T2S_GEMM_MODE_CASE(T2S_GEMM_DGELU | T2S_GEMM_RES_F32 | GEMM_MODE_RES)
            default: break;
        }
#undef T2S_GEMM_MODE_CASE
    }
    return launch_gemm_mode<BN, -1, PAIR>(ta, tb, tc, tr, ep, M, N, K, k_lo_off, sm_cap, st);
}


// ------------------------------------------------------------------------------------ K8b: weight-gradient GEMM
// dW[P, Q] += sum_r G[r, P]^T . X[r, Q]      (P = out features, Q = in features, r = the M rows of the layer)
//
// Replaces the `addmm` autograd issues for nn.Linear.weight.grad (reference: loss.backward() in
// pythia/trainers/base_trainer.py:264 over every Linear of pythia/models/t2s.py).  Both operands are read as they
// lie in HBM (row-major [rows, features], the layout the forward wrote): a TMA box of [64 rows x 64 features] is an
// MN-major 128B-swizzled operand tile, so no transposed copy of the activations is ever made.  The contraction runs
// over the rows: split into `splits` row ranges so that tiles x splits fills the SMs; each work item accumulates
// its range in TMEM and adds its 128 x BN fp32 tile into dW with TMA reduce-add (fp32 red at L2) -- the same
// instruction also accumulates the contributions of the three grounding variants and of encoder / decoder rows.
// Warp roles and pipeline as in gemm_bf16_tcgen05_kernel.
template <int BN>
struct WgradCfg {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;         // two [64 x 64] boxes
    static constexpr int B_BYTES = BN * GEMM_BK * 2;              // BN / 64 boxes
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int STAGING_BYTES = GEMM_EPI_WARPS * 4096;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 256;
    // TWO (cta_group::2, as in the forward GEMM): a pair takes two vertically adjacent P tiles of one Q block; each CTA
    // keeps its [128 P x 64 rows] tile of G and HALF of the X tile -- 32 KB stages, six of them
    static constexpr int STAGE2_BYTES = A_BYTES + B_BYTES / 2;
    static constexpr int STAGES2 = 6;
    static constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + STAGING_BYTES + 1024 + 256;
};

template <int BN, bool TWO>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                          const __grid_constant__ CUtensorMap tmD, int rows, int P, int Q, int splits,
                          int kb_per_split) {
    using Cfg = WgradCfg<BN>;
    constexpr int STAGES = TWO ? Cfg::STAGES2 : Cfg::STAGES;
    constexpr int STAGE_BYTES = TWO ? Cfg::STAGE2_BYTES : Cfg::STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + STAGES * STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_q = (Q + BN - 1) / BN;
    const int num_p = (P + GEMM_BM - 1) / GEMM_BM;
    // TWO: "tiles" counts pair tiles (P tiles 2p and 2p + 1 of one Q block; the host only takes this form for an even
    // number of P tiles); this CTA's item list is that of its pair, of which it takes P tile 2p + rank
    const int rank = TWO ? (int)cluster_ctarank() : 0;
    const int tiles = (TWO ? num_p / 2 : num_p) * num_q;
    const int items = tiles * splits;             // item = split * tiles + tile: neighbours share a row range in L2
    const int item0 = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int item_step = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int kb_total = (rows + GEMM_BK - 1) / GEMM_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmG);
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmD);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(&tfull[a], 1);
                mbar_init(&tempty[a], TWO ? 2 * GEMM_EPI_WARPS : GEMM_EPI_WARPS);
            }
            fence_barrier_init();
        }
        __syncwarp();
        if (TWO) {
            tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
            tmem_relinquish_2sm();
        } else {
            tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (TWO) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int item = item0; item < items; item += item_step) {
            const int split = item / tiles, tile = item - split * tiles;
            const int p0 = (TWO ? 2 * (tile / num_q) + rank : tile / num_q) * GEMM_BM, q0 = (tile % num_q) * BN;
            const int kb0 = split * kb_per_split;
            const int kb1 = min(kb_total, kb0 + kb_per_split);
            for (int kb = kb0; kb < kb1; ++kb) {
                if (lane == 0) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    const int r = kb * GEMM_BK;
                    if (TWO) {
                        // both CTAs' bytes are counted on the leader's barrier; this CTA's half of the X tile
                        const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
                        if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * STAGE_BYTES);
                        tma_load_2d_2sm(sa, &tmG, lead_full, p0, r);
                        tma_load_2d_2sm(sa + 8192, &tmG, lead_full, p0 + 64, r);
#pragma unroll
                        for (int j = 0; j < BN / 128; ++j)
                            tma_load_2d_2sm(sa + Cfg::A_BYTES + j * 8192, &tmX, lead_full, q0 + rank * (BN / 2) + 64 * j, r);
                    } else {
                        mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                        tma_load_2d(sa, &tmG, &full[stage], p0, r);
                        tma_load_2d(sa + 8192, &tmG, &full[stage], p0 + 64, r);
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j)
                            tma_load_2d(sa + Cfg::A_BYTES + j * 8192, &tmX, &full[stage], q0 + 64 * j, r);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_bf16(TWO ? 2 * GEMM_BM : GEMM_BM, BN) | (1u << 15) | (1u << 16);    // A and B MN-major
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int item = item0; item < items && !(TWO && rank != 0); item += item_step) {
            const int split = item / tiles;
            const int kb0 = split * kb_per_split;
            const int kb1 = min(kb_total, kb0 + kb_per_split);
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        // 16 rows of the contraction = two 8-row groups = 2048 B
                        const uint64_t da = make_sw128_mnmajor_desc_lbo(sa + k * 2048, 8192);
                        const uint64_t db = make_sw128_mnmajor_desc_lbo(sa + Cfg::A_BYTES + k * 2048, 8192);
                        if (TWO) umma_bf16_2sm(d_tmem, da, db, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                        else umma_bf16(d_tmem, da, db, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                    }
                    if (TWO) umma_commit_2sm(&empty[stage], 3); else umma_commit(&empty[stage]);
                    if (kb == kb1 - 1) {
                        if (TWO) umma_commit_2sm(&tfull[acc], 3); else umma_commit(&tfull[acc]);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else {
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int half = ew >> 2;
        constexpr int COLS_PER_WARP = BN / 2;
        uint8_t* box = staging + ew * 4096;
        uint8_t* my_row = box + lane * 128;
        const int sw = lane & 7;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = item0; item < items; item += item_step) {
            const int split = item / tiles, tile = item - split * tiles;
            const int p0 = (TWO ? 2 * (tile / num_q) + rank : tile / num_q) * GEMM_BM, q0 = (tile % num_q) * BN;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const int row0 = p0 + quarter * 32;
#pragma unroll 1
            for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 32) {
                uint32_t r[32];
                const int cw = half * COLS_PER_WARP + c0;
                tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + cw, r);
                tmem_ld_wait();
                const int col0 = q0 + cw;
                if (col0 >= Q || row0 >= P) continue;      // warp-uniform; the tensor map clips partial boxes
                if (lane == 0) tma_store_wait_read<0>();
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4*>(my_row + ((c ^ sw) << 4)) =
                        make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_reduce_add_2d(&tmD, box, col0, row0);
                    tma_store_commit();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (TWO) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0)); else mbar_arrive(&tempty[acc]);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (lane == 0) tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (TWO) cluster_sync_all();          // neither CTA leaves while the pair's MMAs may still touch its memory
    if (warp == 1) {
        tc_fence_after();
        if (TWO) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int BN, bool TWO>
static int launch_wgrad(const CUtensorMap& tg, const CUtensorMap& tx, const CUtensorMap& td, int rows, int P, int Q,
                        int splits, int kb_per_split, cudaStream_t st) {
    using Cfg = WgradCfg<BN>;
    constexpr int SMEM = TWO ? Cfg::SMEM2_BYTES : Cfg::SMEM_BYTES;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_tcgen05_kernel<BN, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             SMEM);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(wgrad BN=%d): %s", BN, cudaGetErrorString(e)); return (int)e; }
        attr_set = true;
    }
    const int num_p = (P + GEMM_BM - 1) / GEMM_BM;
    const int items = (TWO ? num_p / 2 : num_p) * ((Q + BN - 1) / BN) * splits;
    if (TWO) {
        const int pairs = items < num_sms() / 2 ? items : num_sms() / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(GEMM_THREADS);
        cfg.dynamicSmemBytes = SMEM;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_wgrad_tcgen05_kernel<BN, TWO>, tg, tx, td, rows, P, Q, splits, kb_per_split);
        if (e != cudaSuccess) { set_error("gemm_wgrad_tcgen05 (pair): %s", cudaGetErrorString(e)); return (int)e; }
        return launch_status("gemm_wgrad_tcgen05 (pair)");
    }
    const int grid = items < num_sms() ? items : num_sms();
    gemm_wgrad_tcgen05_kernel<BN, TWO><<<grid, GEMM_THREADS, SMEM, st>>>(tg, tx, td, rows, P, Q, splits, kb_per_split);
    return launch_status("gemm_wgrad_tcgen05");
}

}  // namespace t2s

using namespace t2s;

static int gemm_entry(const char* who, bool x3, const void* A, long long lda, const void* W, long long ldw,
                      const float* bias, const void* residual, long long ldr, void* C, long long ldc, int M, int N,
                      int K, int flags, int block_n, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0) { set_error("%s: bad shape %d %d %d", who, M, N, K); return T2S_ERR_SHAPE; }
    if ((lda % 8) || (ldw % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15)) {
        set_error("%s: A/W need 16-byte aligned base and row pitch (lda %lld ldw %lld)", who, lda, ldw);
        return T2S_ERR_ALIGN;
    }
    if (x3 && ((K % GEMM_BK) || lda < 2LL * K || ldw < 2LL * K)) {
        set_error("%s: split operands need K %% 64 == 0 and row pitch >= 2K (K %d lda %lld ldw %lld)", who, K, lda, ldw);
        return T2S_ERR_SHAPE;
    }
    const bool out_f32 = flags & T2S_GEMM_OUT_F32;
    const bool out_split = flags & T2S_GEMM_OUT_SPLIT;
    if (out_f32 && out_split) { set_error("%s: OUT_F32 and OUT_SPLIT are exclusive", who); return T2S_ERR_ARG; }
    if ((ldc % (out_f32 ? 4 : 8)) || (reinterpret_cast<uintptr_t>(C) & 15)) {
        set_error("%s: C needs 16-byte aligned base and row pitch (ldc %lld)", who, ldc);
        return T2S_ERR_ALIGN;
    }
    if (out_split && ((N % 64) || ldc < 2LL * N)) { set_error("%s: OUT_SPLIT needs N %% 64 == 0 and ldc >= 2N", who); return T2S_ERR_SHAPE; }
    if (residual && ((ldr % ((flags & T2S_GEMM_RES_F32) ? 4 : 8)) || (reinterpret_cast<uintptr_t>(residual) & 15))) {
        set_error("%s: residual alignment (ldr %lld)", who, ldr);
        return T2S_ERR_ALIGN;
    }
    if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) { set_error("%s: bias alignment", who); return T2S_ERR_ALIGN; }
    int bn = block_n;
    if (bn == 0) {
        // largest tile that still gives every SM work; small problems take the narrow tile
        const long long mt = (M + GEMM_BM - 1) / GEMM_BM;
        if (mt * ((N + 255) / 256) >= num_sms()) bn = 256;
        else if (mt * ((N + 127) / 128) >= num_sms()) bn = 128;
        else bn = 64;
    }
    const int kcols = x3 ? 2 * K : K;      // columns the tensor maps may touch
    CUtensorMap ta, tb;
    int rc = make_tmap_bf16(&ta, A, M, kcols, lda, bn == 64 ? 64 : GEMM_BM);
    if (rc) return rc;
    // CTA pairs for the throughput tile (see the kernel); T2S_GEMM_PAIR=0 in the environment turns them off
    // (default 2: one cta_group::2 MMA per pair; T2S_GEMM_PAIR=1: two MMAs over a multicast W tile; 0: single CTAs.
    // Same box, eval step: 22.4-22.5 ms with 1, 21.6-21.8 ms with 2, results bit-identical)
    static int pair_env = -1;
    if (pair_env < 0) {
        const char* e = getenv("T2S_GEMM_PAIR");
        pair_env = e ? (e[0] == '0' ? 0 : (e[0] == '1' ? 1 : 2)) : 2;
    }
    const int pair = (bn == 256 && M >= 2 * GEMM_BM) ? pair_env : 0;
    rc = make_tmap_bf16(&tb, W, N, kcols, ldw, pair ? bn / 2 : bn);
    if (rc) return rc;
    // C is written by TMA stores of 32-row x 128-byte boxes (clipped to [M, N] / [M, 2N] by the map)
    CUtensorMap tc;
    rc = make_tmap_2d(&tc, out_f32, C, M, out_split ? 2LL * N : N, ldc, out_f32 ? 32 : 64, 32);
    if (rc) return rc;
    // residual tile map of the TMA-fed + bf16 residual epilogue (same 32-row x 64-column boxes as the bf16 output)
    CUtensorMap tr = tc;
    if (residual && !(flags & T2S_GEMM_RES_F32) && !out_f32 && !out_split) {
        rc = make_tmap_2d(&tr, false, residual, M, N, ldr, 64, 32);
        if (rc) return rc;
    }
    GemmEpi ep{C, bias, residual, ldc, ldr, flags, N, 0};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int k_lo = x3 ? K : 0;
    const int cap = (flags >> T2S_GEMM_SM_CAP_SHIFT) & 0xff;
    switch (bn) {
        case 256: return pair == 2 ? launch_gemm<256, 2>(ta, tb, tc, tr, ep, M, N, K, k_lo, cap, st)
                       : pair == 1 ? launch_gemm<256, 1>(ta, tb, tc, tr, ep, M, N, K, k_lo, cap, st)
                                   : launch_gemm<256, 0>(ta, tb, tc, tr, ep, M, N, K, k_lo, cap, st);
        case 128: return launch_gemm<128, 0>(ta, tb, tc, tr, ep, M, N, K, k_lo, cap, st);
        case 64: return launch_gemm<64, 0>(ta, tb, tc, tr, ep, M, N, K, k_lo, cap, st);
        default: set_error("%s: block_n must be 0, 64, 128 or 256", who); return T2S_ERR_ARG;
    }
}

/* hits / misses of the tensor-map descriptor cache since the library was loaded (diagnostics, tests) */
extern "C" long long t2s_tmap_cache_stats(int which) {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    return which == 0 ? g_tmap_hits : g_tmap_misses;
}

extern "C" int t2s_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                             const void* residual, long long ldr, void* C, long long ldc, int M, int N, int K,
                             int flags, int block_n, void* stream) {
    return gemm_entry("gemm_bf16", false, A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, flags, block_n, stream);
}

extern "C" int t2s_gemm_bf16x3(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                               const void* residual, long long ldr, void* C, long long ldc, int M, int N, int K,
                               int flags, int block_n, void* stream) {
    return gemm_entry("gemm_bf16x3", true, A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, flags, block_n, stream);
}

extern "C" int t2s_gemm_wgrad_bf16(const void* G, long long ldg, const void* X, long long ldx, float* dW, long long ldd,
                                   int rows, int P, int Q, int splits, void* stream) {
    if (rows <= 0 || P <= 0 || Q <= 0) { set_error("gemm_wgrad: bad shape %d %d %d", rows, P, Q); return T2S_ERR_SHAPE; }
    if ((ldg % 8) || (ldx % 8) || (ldd % 4) || (reinterpret_cast<uintptr_t>(G) & 15) || (reinterpret_cast<uintptr_t>(X) & 15) ||
        (reinterpret_cast<uintptr_t>(dW) & 15)) {
        set_error("gemm_wgrad: operands need 16-byte aligned bases and row pitches (ldg %lld ldx %lld ldd %lld)", ldg, ldx, ldd);
        return T2S_ERR_ALIGN;
    }
    const int bn = Q >= 192 ? 256 : 64;
    const int tiles = ((P + GEMM_BM - 1) / GEMM_BM) * ((Q + bn - 1) / bn);
    const int kb_total = (rows + GEMM_BK - 1) / GEMM_BK;
    if (splits <= 0) {
        // fill the SMs about four times over, but keep >= 16 k-blocks (1024 rows) per item so that the fp32
        // reduce-add of the 128 x BN tile stays a small part of the item
        splits = (4 * num_sms()) / tiles;
        const int max_splits = (kb_total + 15) / 16;
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    if (splits > kb_total) splits = kb_total;
    int kb_per_split = (kb_total + splits - 1) / splits;
    splits = (kb_total + kb_per_split - 1) / kb_per_split;         // no empty items
    CUtensorMap tg, tx, td;
    int rc = make_tmap_2d(&tg, false, G, rows, P, ldg, 64, 64);
    if (rc) return rc;
    rc = make_tmap_2d(&tx, false, X, rows, Q, ldx, 64, 64);
    if (rc) return rc;
    rc = make_tmap_2d(&td, true, dW, P, Q, ldd, 32, 32);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // pair form (one cta_group::2 MMA of M = 256 per CTA pair) when the P tiles pair up; T2S_WGRAD_PAIR=0 turns it off
    static int wpair = -1;
    if (wpair < 0) {
        const char* e = getenv("T2S_WGRAD_PAIR");
        wpair = (e && e[0] == '0') ? 0 : 1;
    }
    if (bn == 256) {
        if (wpair && ((P + GEMM_BM - 1) / GEMM_BM) % 2 == 0)
            return launch_wgrad<256, true>(tg, tx, td, rows, P, Q, splits, kb_per_split, st);
        return launch_wgrad<256, false>(tg, tx, td, rows, P, Q, splits, kb_per_split, st);
    }
    return launch_wgrad<64, false>(tg, tx, td, rows, P, Q, splits, kb_per_split, st);
}
