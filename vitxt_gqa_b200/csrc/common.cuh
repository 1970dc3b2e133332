// Shared device/host helpers for libt2s_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#define T2S_OK 0
#define T2S_ERR_SHAPE (-1)
#define T2S_ERR_ALIGN (-2)
#define T2S_ERR_DRIVER (-3)
#define T2S_ERR_ARG (-4)

namespace t2s {

void set_error(const char* fmt, ...);
int num_sms();

// returns 0 or the (positive) cudaError_t of the launch just issued
inline int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return T2S_OK;
}

constexpr int kWarp = 32;

// ---------------------------------------------------------------- programmatic dependent launch (greedy decode chain)
// The 26 kernels of a greedy decode step depend on each other one by one and each runs ~10 us: launch latency and the
// set-up of the next kernel (barrier init, TMEM allocation, descriptor prefetch) are a visible part of the chain.  A kernel
// launched through launch_pdl() may become resident while its predecessor in the stream still runs; it executes
// pdl_wait() before it touches anything the predecessor reads or writes (griddepcontrol.wait returns when the preceding
// grid has completed and its writes are visible) and then calls pdl_release() so that ITS successor may come in -- in
// that order, so that at most one grid sits waiting.  Both instructions do nothing in a normally launched kernel.
// Only the small launches of the decode chain ask for it (`early`): a waiting grid of tens of thousands of CTAs would sit
// on the thread slots the other stream needs.  T2S_PDL=0 in the environment launches everything the ordinary way.
bool pdl_enabled();
constexpr int PDL_MAX_ROWS = 4096;      // row-wise kernels: launches of at most this many rows come in early
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch_pdl(bool early, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                       Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (early && pdl_enabled()) ? 1 : 0;        // early = false: an ordinary launch
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);       // errors surface through launch_status()
}

// index into a table of `n` rows, forced into range (bad feedback / teacher-forcing indices must not read out of bounds)
__device__ __forceinline__ long long clamp_index(long long i, long long n) {
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

// ---------------------------------------------------------------- dropout (training step only)
// Counter-based masks: nothing is stored, the backward recomputes the mask of a site from (seed, site, element).
// One 32-bit hash serves TWO adjacent elements (its 16-bit halves against thr = round(p * 65536)), the pair being
// (even column, odd column) of a row -- the accumulator layout of mma.sync and the per-row sweeps of the softmax both
// own such pairs.  Kept values are scaled by 65536 / (65536 - thr), the inverse of the exact keep probability.
//   row-major [rows, H] sites:    x = row * (H / 2) + col / 2          y = site << 20
//   attention probabilities:      x = (query << 16) | (key >> 1)       y = (site << 20) | (sample * heads + head)
// (query / key = positions in the virtual sequence of csrc/attn_bwd.cu: encoder rows then decoder rows; keys by their
// position in the compacted key list, then the decoder keys).
struct DropCfg {
    uint32_t s0, s1;      // seed
    uint32_t site;        // < 4096
    uint32_t thr;         // 0 = dropout off
    float scale;
};
inline DropCfg make_drop(float p, unsigned long long seed, unsigned site) {
    DropCfg d;
    d.s0 = (uint32_t)seed;
    d.s1 = (uint32_t)(seed >> 32);
    d.site = site & 0xfffu;
    long t = lrintf(p * 65536.0f);
    d.thr = p > 0.f ? (uint32_t)(t < 1 ? 1 : (t > 65535 ? 65535 : t)) : 0u;
    d.scale = 65536.0f / (65536.0f - (float)d.thr);
    return d;
}
__device__ __forceinline__ uint32_t drop_hash(uint32_t s0, uint32_t s1, uint32_t x, uint32_t y) {
    uint32_t h = x ^ s0;
    h *= 0x9E3779B1u;
    h ^= h >> 15;
    h += y * 0x85EBCA6Bu + s1;
    h *= 0xC2B2AE35u;
    h ^= h >> 13;
    h *= 0x27D4EB2Fu;
    h ^= h >> 16;
    h *= 0x165667B1u;
    h ^= h >> 15;
    return h;
}
// multipliers (0 or scale) of the element pair (x, y)
__device__ __forceinline__ void drop_pair(const DropCfg& d, uint32_t x, uint32_t y, float& m0, float& m1) {
    const uint32_t h = drop_hash(d.s0, d.s1, x, y);
    m0 = (h & 0xffffu) >= d.thr ? d.scale : 0.f;
    m1 = (h >> 16) >= d.thr ? d.scale : 0.f;
}
// four consecutive elements starting at the even column `col` of row `row` of a [rows, H] site
__device__ __forceinline__ float4 drop_mask4(const DropCfg& d, int row, int H, int col) {
    const uint32_t x = (uint32_t)row * (uint32_t)(H >> 1) + (uint32_t)(col >> 1);
    float4 m;
    drop_pair(d, x, d.site << 20, m.x, m.y);
    drop_pair(d, x + 1, d.site << 20, m.z, m.w);
    return m;
}
__device__ __forceinline__ uint32_t drop_attn_y(const DropCfg& d, int bh) { return (d.site << 20) | (uint32_t)bh; }
__device__ __forceinline__ uint32_t drop_attn_x(int query, int key) { return ((uint32_t)query << 16) | ((uint32_t)key >> 1); }

// ---------------------------------------------------------------- warp / block reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Block-wide sum; `red` is >= 33 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        float t = lane < nw ? red[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        float t = lane < nw ? red[lane] : -INFINITY;
        t = warp_max(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// exact-erf GELU as in pytorch_transformers modeling_bert.gelu
//   x * 0.5 * (1 + erf(x / sqrt 2))
// erf(t) = 1 - 2^(-t Q(t)) on t = |x| / sqrt 2 clamped to 4 (erfc(4) = 1.5e-8), Q a degree-6 least-squares fit of
// -log2(erfc(t)) / t: |erf error| <= 1.9e-7 in fp32 arithmetic, |GELU error| <= 1.3e-7 absolute -- the size of the
// fp32 rounding of erff itself -- in 13 branch-free instructions (6 FFMA + 1 MUFU.EX2) instead of erff's two
// divergent polynomial branches (~35), which made the GELU epilogue of the 128x256 tile longer than its main loop.
__device__ __forceinline__ float gelu_erf(float x) {
    const float h = 0.5f * x, ah = fabsf(h);
    const float t = fminf(fabsf(x) * 0.70710678118654752440f, 4.0f);
    float q = -8.592197123e-05f;
    q = fmaf(q, t, 3.653188699e-04f);
    q = fmaf(q, t, 2.547933478e-03f);
    q = fmaf(q, t, -2.975212620e-02f);
    q = fmaf(q, t, 1.491437337e-01f);
    q = fmaf(q, t, 9.182796953e-01f);
    q = fmaf(q, t, 1.627918195e+00f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-t * q));
    return fmaf(-ah, e, h + ah);      // h + |h| erf(t) = h (1 + sign(x) erf(t))
}

// d/dx of the GELU above: Phi(x) + x phi(x), same erf approximation (used by the dgrad epilogue of the FFN)
__device__ __forceinline__ float gelu_grad(float x) {
    const float t = fminf(fabsf(x) * 0.70710678118654752440f, 4.0f);
    float q = -8.592197123e-05f;
    q = fmaf(q, t, 3.653188699e-04f);
    q = fmaf(q, t, 2.547933478e-03f);
    q = fmaf(q, t, -2.975212620e-02f);
    q = fmaf(q, t, 1.491437337e-01f);
    q = fmaf(q, t, 9.182796953e-01f);
    q = fmaf(q, t, 1.627918195e+00f);
    float e, g;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-t * q));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(-0.72134752044448170368f * x * x));     // exp(-x^2 / 2)
    const float erf_abs = 1.0f - e;                                   // erf(|x| / sqrt 2)
    const float cdf = 0.5f + copysignf(0.5f * erf_abs, x);
    return fmaf(x * 0.39894228040143267794f, g, cdf);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// ---------------------------------------------------------------- PTX: smem / mbarrier / TMA
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(desc) : "memory");
}
// 2D tiled TMA load: coordinates (c0 = innermost element index, c1 = row index)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// the same load delivered to the same shared-memory offset (and mbarrier offset) of every CTA in cta_mask
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1,
                                                      uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
// thread-block cluster helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// 2D tiled TMA store of a shared-memory box (bulk-group completion); clips to the tensor bounds
__device__ __forceinline__ void tma_store_2d(const void* desc, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(desc),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
// same box, but added into global memory (fp32 red.add performed at L2; used by the split-K weight-gradient GEMM)
__device__ __forceinline__ void tma_reduce_add_2d(const void* desc, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(desc),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N bulk groups of this thread have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- PTX: tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in tensor memory (lane = row, one 32-bit column = two consecutive K elements): D (+)= A[tmem] . B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// the same arrive delivered to the mbarrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}
// ---- cta_group::2: the two CTAs of a cluster pair run ONE tcgen05.mma (M = 256) issued by the leader (cluster rank 0).
// Each CTA keeps its own 128 rows of A and its own half of the B tile in shared memory (same offsets in both CTAs: the
// leader's descriptors are applied to both) and receives its 128 accumulator rows in its own tensor memory.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta_rank) {      // shared::cta -> shared::cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are counted on an mbarrier of the pair's leader
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* desc, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(desc), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier at this offset in every CTA of cta_mask once the pair's MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base_lane + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// same wait, but tied to the 32 destination registers of one tcgen05.ld: their consumers depend on this statement,
// so a second tcgen05.ld can be issued behind it without the compiler moving reads of `r` above the wait
__device__ __forceinline__ void tmem_ld_wait_on(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
// 32 lanes x 16 consecutive fp32 columns (half the registers of the x32 form: two of these double-buffer a sweep)
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_on(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
// registers -> 32 lanes x 32 consecutive 32-bit TMEM columns (thread i writes lane base_lane + i)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major bf16 tile stored by TMA with the
// 128-byte swizzle: rows are 128 B apart, 8-row groups 1024 B apart (SBO), LBO unused.
// Bit layout (cute/arch/mma_sm100_desc.hpp SmemDescriptor): [0,14) addr>>4, [16,30) LBO>>4,
// [32,46) SBO>>4, [46,48) version=1, [61,64) layout type (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// MN-major bf16 operand (rows = K index, 64 contiguous M/N elements = one 128-byte row, 128B swizzle -- what a TMA
// box of [rows x 64 columns] of a row-major matrix produces): 8-row K groups `SBO` = 1024 B apart, 64-element
// atoms along M/N `lbo_bytes` apart (cute: Swizzle<3,4,3> o ((8,8,m),(8,k)):((1,8,LBO),(64,SBO))).
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc_lbo(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor, kind::f16: c=F32 (bit4), a=b=BF16 (bits 7,10), K-major both,
// N>>3 at [17,23), M>>4 at [24,29)  (cute/arch/mma_sm100_desc.hpp InstrDescriptor).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------- cp.async (LDGSTS)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace t2s
