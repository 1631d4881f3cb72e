// CTA-pair variant of the fused TensorDense TT-matvec (csrc/ttdense.cu has the single-CTA kernel and the full description of the algorithm):
// a thread-block cluster of two CTAs on the two SMs of a TPC works on ONE sample at a time with tcgen05.mma.cta_group::2.
//
// Why: in the single-CTA kernel the two CTAs that share a sample (one per half of o2) each need ALL of T1, so step 1 and the T1 re-layout
// are done twice, each CTA issues 160 small MMAs per sample from one or two warps, and only three 16 KB operand slots fit beside half of
// G2.  With cta_group::2 the MMA is M = 256 x N = 256 across the pair: CTA r owns the M-half "o3 in half r" of every accumulator (its own
// TMEM) and supplies half of every B operand from its own shared memory (the o2-half r of G2: the same 128 KB image), so
//   * each CTA computes and re-lays out only ITS half of T1 (one 8 KB operand slot per chunk instead of 16 KB: eight slots fit),
//   * the leader CTA issues 96 MMAs per sample for both (step 1: 16 x 2 of 256x16x8, step 2: 16 x 2 of 256x256x8, step 3: 8 x 4 of 256x16x8),
//   * the B-operand shared-memory reads per MMA are halved.
//
//   step 1   D1[(o3l,b2) | CTA r][i1]        = G3^T tile r (128 x 16) . x[(i1), i3]^T          x rows i1 split: CTA r loads i1 in [8r, 8r+8)
//   step 2   D2[(o3l,i1) | CTA r][(osrc,b1,o2l)] += T1 chunk (128 x 16, slot of CTA r) . G2[(b1,o2l) | CTA osrc][(i2,b2)]^T
//   step 3   D3[(o2,o3l) | CTA r][o1]        += T2 chunk (128 x 32, two slots of CTA r) . G1[o1 | CTA osrc = o1 / 8][(b1,i1)]^T
//
// Synchronisation: every "consumer-side" barrier the issuing warps wait on lives in the LEADER's shared memory and counts arrivals from
// the epilogue threads of both CTAs (the peer arrives through the cluster address, mapa); every "producer-side" barrier (accumulator
// ready, slot free) is signalled in BOTH CTAs by one tcgen05.commit.multicast.  The pair is launched with __cluster_dims__(2, 1, 1).
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"

namespace syn {
namespace ttp {

constexpr int GROUPS = 2;                                       // epilogue groups of four warps (per CTA)
constexpr int THREADS = 64 + 128 * GROUPS;                      // TMA + step 1, steps 2 / 3, the epilogue groups
constexpr int UNITS = 8;                                        // ring of 8 KB operand units: a T1 chunk takes one, a T2 chunk an aligned pair
constexpr int NB1 = 4;                                          // step-1 accumulator buffers
constexpr uint32_t UNIT_BYTES = 8192;
// shared-memory map (offsets from a 1024-byte aligned base; identical in both CTAs)
constexpr uint32_t OFF_B2 = 0;                                  // this CTA's o2-half of G2: 8 atoms x [128 rows x 128 B]  (SW128)
constexpr uint32_t OFF_UNIT = OFF_B2 + 131072;                  // UNITS x 8 KB
constexpr uint32_t OFF_A1 = OFF_UNIT + UNITS * UNIT_BYTES;      // this CTA's o3-half tile of G3^T: [128 rows x 64 B]      (SW64)
constexpr uint32_t OFF_B3 = OFF_A1 + 8192;                      // this CTA's o1-half of G1: 8 atoms x [8 rows x 128 B]     (SW128)
constexpr uint32_t OFF_X = OFF_B3 + 8192;                       // this CTA's i1-half of the sample: 128 rows (i2, i1l) x 64 B (SW64, TMA)
constexpr uint32_t OFF_BAR = OFF_X + 8192;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 512 + 1024;
// packed weight image (floats), shared with the single-CTA kernel: [A1 16 KB][B3 16 KB][B2 half 0][B2 half 1][B3 halves for the pair 16 KB]
constexpr size_t IMG_A1 = 0, IMG_B2 = 8192, IMG_B3P = 8192 + 2 * 32768;
constexpr uint32_t TM_D1 = 0, TM_D2 = 16 * NB1, TM_D3 = TM_D2 + 256, TM_COLS = 512;
static_assert(TM_D3 + 16 <= TM_COLS, "TMEM columns");
// a sample uses 16 + 2 * 8 = 32 units and 16 D1 buffers: positions and barrier parities must not depend on the sample index
static_assert(16 % UNITS == 0 && (16 / UNITS) % 2 == 0 && UNITS % 2 == 0, "unit ring: every unit is used an even number of times by each phase of a sample");
static_assert(16 % NB1 == 0 && (16 / NB1) % 2 == 0, "D1 ring");

enum Bar { W_FULL = 0, X_FULL, X_PEER, X_EMPTY, D1_FULL0, D1_EMPTY0 = D1_FULL0 + NB1, 
           // the unit ring has separate barriers for its T1 uses (uses 0 / 1 of a sample: re-layout 1 -> step 2) and its T2 uses (uses 2 / 3:
           // re-layout 2 -> step 3), so that each issuing warp sees the phases of its own barriers strictly in order
           U1_FULL0 = D1_EMPTY0 + NB1, U2_FULL0 = U1_FULL0 + UNITS, U1_EMPTY0 = U2_FULL0 + UNITS, U2_EMPTY0 = U1_EMPTY0 + UNITS,
           D2_FULL = U2_EMPTY0 + UNITS, D2_EMPTY, D3_FULL, D3_EMPTY, NUM_BARS };
static_assert(NUM_BARS * 8 + 8 <= 512, "barrier area");

__host__ __device__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t kbyte) {
    return (row >> 3) * 1024u + (row & 7u) * 128u + ((((kbyte >> 4) ^ (row & 7u)) & 7u) << 4) + (kbyte & 15u);
}

// ---- PTX wrappers -------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  Default semantics, the form CUTLASS's
// ClusterBarrier::arrive(cta_id) uses: an explicit .release.cluster compiled to MEMBAR.ALL.GPU (+ CCTL.IVALL behind every .acquire.cluster wait)
// and cost ~1000 cycles per hand-over (measured with the SYN_TT_DEBUG timeline).  Nothing the waiter consumes travels through generic
// memory: the operands go writer -> fence.proxy.async -> tensor core, the accumulators through tcgen05 fences.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// ONE arrival per warp on the leader's barrier: the warp's lanes have finished (and fenced) their part, __syncwarp orders their accesses
// before lane 0's arrival -- 8 arrivals per hand-over instead of 256, and only 4 of them cross the cluster network
__device__ __forceinline__ void warp_arrive_leader(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_cluster(bar, 0);
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {      // barriers that receive remote arrivals
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both CTAs] * B[smem of both CTAs], M = 256 across the pair; issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma2_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` of BOTH CTAs when they complete
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __noinline__ void wait_timeout(int id, uint32_t parity) {
    printf("tt_dense3_tf32_pair_kernel: block %d thread %d starved on barrier %d (parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, id, parity);
    __trap();
}
// bounded waits (a pipeline bug prints the starving barrier instead of hanging the GPU); cluster-scope acquire: the peer arrives remotely
__device__ __forceinline__ void bwait(uint64_t* bars, int id, uint32_t parity) {
    if (mbar_try_wait_cluster(&bars[id], parity)) return;
    const uint64_t t0 = global_ns();
    for (unsigned spin = 1; !mbar_try_wait_cluster(&bars[id], parity); ++spin)
        if ((spin & 255u) == 0 && global_ns() - t0 > 2000000000ull) wait_timeout(id, parity);
}
// barriers signalled by tcgen05.commit or TMA only (never by a remote thread): the plain CTA-scope wait of common.cuh
__device__ __forceinline__ void lwait(uint64_t* bars, int id, uint32_t parity) {
    if (mbar_try_wait(&bars[id], parity)) return;
    const uint64_t t0 = global_ns();
    for (unsigned spin = 1; !mbar_try_wait(&bars[id], parity); ++spin)
        if ((spin & 255u) == 0 && global_ns() - t0 > 2000000000ull) wait_timeout(id, parity);
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)layout_type << 61);
}
constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_SW64 = 4;
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

#ifdef SYN_TT_DEBUG
// timeline trace of the first pair (compile with -DSYN_TT_DEBUG): every mark appends (value, clock) to [cta][warp][PT_TRACE_LEN][2]
__device__ uint32_t* g_ttp_dbg = nullptr;
constexpr int PT_TRACE_LEN = 4096;
#define PT_MARK(tag_, j_, c_)                                                                                        \
    do {                                                                                                             \
        if (g_ttp_dbg && blockIdx.x < 2 && (threadIdx.x & 31) == 0 && pt_cnt_ < PT_TRACE_LEN) {                      \
            uint32_t* w_ = g_ttp_dbg + (((size_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5)) * PT_TRACE_LEN + pt_cnt_) * 2; \
            w_[0] = ((uint32_t)(tag_) << 12) + ((uint32_t)(j_) & 15u) * 256u + (uint32_t)(c_);                        \
            w_[1] = (uint32_t)clock64();                                                                             \
            ++pt_cnt_;                                                                                               \
        }                                                                                                            \
    } while (0)
#else
#define PT_MARK(tag_, j_, c_) do { } while (0)
#endif

// the pair's extra piece of the weight image: B3 halves -- [r][atom a][row = o1l (8)][kk = (b1 & 1)*16 + i1], b1 = 2a + (kk >> 4) = G1[i1, 8r + o1l, b1]
__global__ void __launch_bounds__(256) tt_pack_pair_kernel(const float* __restrict__ G1, float* __restrict__ img) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 4096; e += gridDim.x * blockDim.x) {
        const int kk = e & 31, o1l = (e >> 5) & 7, a = (e >> 8) & 7, r = e >> 11;
        const int b1 = 2 * a + (kk >> 4), i1 = kk & 15, o1 = 8 * r + o1l;
        img[IMG_B3P + (size_t)r * 2048 + (a * 1024 + sw128_off(o1l, kk * 4)) / 4] = G1[(i1 * 16 + o1) * 16 + b1];
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tt_dense3_tf32_pair_kernel(const __grid_constant__ CUtensorMap xmap, const float* __restrict__ img, const float* __restrict__ bias,
                           float* __restrict__ y, int batch, int relu) {
    extern __shared__ uint8_t ttp_smem_raw[];
    const uint32_t raw = smem_u32(ttp_smem_raw);
    uint8_t* smem = ttp_smem_raw + ((1024u - (raw & 1023u)) & 1023u);        // the same offset in both CTAs (same kernel, same static layout)
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NUM_BARS * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef SYN_TT_DEBUG
    uint32_t pt_cnt_ = 0;
#endif
    const uint32_t rank = cluster_rank();                                    // 0 = leader: issues every MMA; r owns o3-half r, supplies o2-half r
    const int pair = blockIdx.x >> 1, pairs = gridDim.x >> 1;
    const int my_samples = pair < batch ? (batch - pair + pairs - 1) / pairs : 0;

    if (threadIdx.x == 0) {
        mbar_init(&bars[W_FULL], 1);
        mbar_init(&bars[X_FULL], 1);
        mbar_init(&bars[X_PEER], 1);
        mbar_init(&bars[X_EMPTY], 1);
        for (int b = 0; b < NB1; b++) {
            mbar_init(&bars[D1_FULL0 + b], 1);
            mbar_init(&bars[D1_EMPTY0 + b], 8);                              // one arrival per epilogue warp of the group, both CTAs
        }
        for (int u = 0; u < UNITS; u++) {
            mbar_init(&bars[U1_FULL0 + u], 8);
            mbar_init(&bars[U2_FULL0 + u], 8);
            mbar_init(&bars[U1_EMPTY0 + u], 1);
            mbar_init(&bars[U2_EMPTY0 + u], 1);
        }
        mbar_init(&bars[D2_FULL], 1);
        mbar_init(&bars[D2_EMPTY], 8 * GROUPS);
        mbar_init(&bars[D3_FULL], 1);
        mbar_init(&bars[D3_EMPTY], 8);
        fence_async_smem();
    }
    __syncthreads();
    if (warp == 1) tmem_alloc2(tmem_ptr, TM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                                      // both CTAs: barriers initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer (both CTAs: own halves) + step-1 issuer (leader) ==================================================================
        constexpr uint32_t ID_S1 = idesc_tf32(256, 16);
        const uint64_t d64 = smem_desc(0, 512, LAYOUT_SW64);
        const uint64_t a1_desc = d64 + ((sbase + OFF_A1) >> 4), x_desc = d64 + ((sbase + OFF_X) >> 4);
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
            mbar_expect_tx(&bars[W_FULL], 8192u + 8192u + 131072u);
            bulk_g2s(smem + OFF_A1, img + IMG_A1 + (size_t)rank * 2048, 8192u, &bars[W_FULL]);
            bulk_g2s(smem + OFF_B3, img + IMG_B3P + (size_t)rank * 2048, 8192u, &bars[W_FULL]);
            for (int a = 0; a < 8; a++)
                bulk_g2s(smem + OFF_B2 + a * 16384, img + IMG_B2 + (size_t)rank * 32768 + (size_t)a * 4096, 16384u, &bars[W_FULL]);
        }
        __syncwarp();
        lwait(bars, W_FULL, 0);
        // weights of BOTH CTAs must be in place before the leader's first MMA reads them: the peer reports through X_PEER (below), which it
        // only signals after its own W_FULL and X_FULL
        for (int j = 0; j < my_samples; j++) {
            lwait(bars, X_EMPTY, (uint32_t)(j & 1) ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(&bars[X_FULL], 8192u);
                tma_load_4d(smem + OFF_X, &xmap, 0, 8 * (int)rank, 0, pair + j * pairs, &bars[X_FULL]);
            }
            __syncwarp();
            lwait(bars, X_FULL, (uint32_t)(j & 1));
            if (rank != 0) {
                if (elect_one()) mbar_arrive_cluster(&bars[X_PEER], 0);      // tell the leader: my half of x (and my weights) are in shared memory
                __syncwarp();
                continue;
            }
            PT_MARK(0x10, j, 0);
            bwait(bars, X_PEER, (uint32_t)(j & 1));
#pragma unroll
            for (int c = 0; c < 16; c++) {
                const uint32_t b = c % NB1, use = c / NB1;
                PT_MARK(0x11, j, c);
                bwait(bars, D1_EMPTY0 + b, (use & 1u) ^ 1u);
                PT_MARK(0x12, j, c);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t xd = x_desc + (uint64_t)(c * 32);         // chunk c: 8 rows x 64 B = 512 B
#pragma unroll
                    for (int k = 0; k < 2; k++) umma2_tf32(tmem + TM_D1 + b * 16, a1_desc + (uint64_t)(k * 2), xd + (uint64_t)(k * 2), ID_S1, k);
                    umma2_commit(&bars[D1_FULL0 + b]);
                    if (c == 15) umma2_commit(&bars[X_EMPTY]);
                }
                __syncwarp();
                PT_MARK(0x13, j, c);
            }
        }
    } else if (warp == 1) {
        // ===== step-2 / step-3 issuer (leader only; a separate step-3 warp measured slower: 7.24 vs 6.65 ms) ===========================================================================================
        if (rank == 0) {
            constexpr uint32_t ID_S2 = idesc_tf32(256, 256), ID_S3 = idesc_tf32(256, 16);
            const uint64_t d64 = smem_desc(0, 512, LAYOUT_SW64), d128 = smem_desc(0, 1024, LAYOUT_SW128);
            const uint64_t b2_desc = d128 + ((sbase + OFF_B2) >> 4), b3_desc = d128 + ((sbase + OFF_B3) >> 4);
            const uint64_t unit64_desc = d64 + ((sbase + OFF_UNIT) >> 4), unit128_desc = d128 + ((sbase + OFF_UNIT) >> 4);
            for (int j = 0; j < my_samples; j++) {
#pragma unroll
                for (int cc = 0; cc < 16; cc++) {
                    const uint32_t unit = cc % UNITS, par = (cc / UNITS) & 1u;                 // this unit's use 0 / 1 of the sample
                    PT_MARK(0x20, j, cc);
                    bwait(bars, U1_FULL0 + unit, par);
                    PT_MARK(0x21, j, cc);
                    if (cc == 0) bwait(bars, D2_EMPTY, (uint32_t)(j & 1) ^ 1u);
                    PT_MARK(0x22, j, cc);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad = unit64_desc + (uint64_t)(unit * (UNIT_BYTES >> 4));
                        const uint64_t bd = b2_desc + (uint64_t)((cc >> 1) * 1024 + (cc & 1) * 4);
#pragma unroll
                        for (int k = 0; k < 2; k++)
                            umma2_tf32(tmem + TM_D2, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), ID_S2, (cc > 0 || k > 0) ? 1u : 0u);
                        umma2_commit(&bars[U1_EMPTY0 + unit]);
                        if (cc == 15) umma2_commit(&bars[D2_FULL]);
                    }
                    __syncwarp();
                    PT_MARK(0x23, j, cc);
                }
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    // a T2 chunk occupies the aligned unit pair (unit, unit + 1); BOTH units carry the full / empty protocol, so that every
                    // unit sees four uses per sample (two T1 chunks, two T2 chunks) and the next T1 chunk in either of them waits for this MMA
                    const uint32_t unit = (2 * p) % UNITS, par = (p / (UNITS / 2)) & 1u;         // uses 2 / 3 of the sample
                    PT_MARK(0x30, j, p);
                    bwait(bars, U2_FULL0 + unit, par);
                    bwait(bars, U2_FULL0 + unit + 1, par);
                    PT_MARK(0x31, j, p);
                    if (p == 0) bwait(bars, D3_EMPTY, (uint32_t)(j & 1) ^ 1u);
                    PT_MARK(0x32, j, p);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad = unit128_desc + (uint64_t)(unit * (UNIT_BYTES >> 4));
                        const uint64_t bd = b3_desc + (uint64_t)(p * 64);                      // atom p of this CTA's G1 half: 1 KB
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            umma2_tf32(tmem + TM_D3, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), ID_S3, (p > 0 || k > 0) ? 1u : 0u);
                        umma2_commit(&bars[U2_EMPTY0 + unit]);
                        umma2_commit(&bars[U2_EMPTY0 + unit + 1]);
                        if (p == 7) umma2_commit(&bars[D3_FULL]);
                    }
                    __syncwarp();
                    PT_MARK(0x33, j, p);
                }
            }
        }
    } else {
        // ===== epilogue warps (both CTAs, own TMEM half): re-layouts and the output ============================================================
        const int grp = (warp - 2) >> 2;
        const int q = warp & 3;
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        const int hb = lane >> 4, lo = lane & 15;
        // re-layout 1: D1 lane = (o3l = 2q + hb, b2 = lo), column i1  ->  unit row m = q*32 + i1*2 + hb, k = b2   (SW64, 128 rows x 64 B)
        const uint32_t r1_base = (uint32_t)(q * 4) * 512u + (uint32_t)hb * 64u + (uint32_t)(lo & 3) * 4u;
        // re-layout 2: D2 lane L = i1*2 + hb' within the quadrant (o3l = 2q + hb'), column = osrc*128 + b1*8 + o2l
        //   ->  unit pair row m' = (osrc*8 + o2l)*8 + hb'*4 + q, k = b1l*16 + i1                                      (SW128, 128 rows x 128 B)
        const int hb2 = lane & 1, i1_2 = lane >> 1;
        for (int j = 0; j < my_samples; j++) {
            const int s = pair + j * pairs;
#pragma unroll
            for (int c = 0; c < 16; c++) {
                if ((16 * j + c) % GROUPS != grp) continue;
                const uint32_t b = c % NB1, use = c / NB1;
                const uint32_t unit = c % UNITS, par = (c / UNITS) & 1u;
                PT_MARK(0x40, j, c);
                lwait(bars, D1_FULL0 + b, use & 1u);
                PT_MARK(0x41, j, c);
                tc_fence_after();
                uint32_t r0[16];
                tmem_ld16(tq + TM_D1 + b * 16, r0);
                tmem_ld_wait();
                tc_fence_before();
                warp_arrive_leader(&bars[D1_EMPTY0 + b]);
                PT_MARK(0x42, j, c);
                lwait(bars, (par == 0 ? U2_EMPTY0 : U1_EMPTY0) + unit, par ^ 1u);    // use 0 follows use 3 of the previous sample, use 1 follows use 0
                PT_MARK(0x43, j, c);
                uint8_t* dst = smem + OFF_UNIT + unit * UNIT_BYTES + r1_base;
#pragma unroll
                for (int i1 = 0; i1 < 16; i1++) {
                    const uint32_t off = (uint32_t)(i1 >> 2) * 512u + (uint32_t)(i1 & 3) * 128u + ((((uint32_t)(lo >> 2)) ^ (uint32_t)(i1 & 3)) << 4);
                    *reinterpret_cast<uint32_t*>(dst + off) = r0[i1];
                }
                fence_async_smem();
                warp_arrive_leader(&bars[U1_FULL0 + unit]);
                PT_MARK(0x44, j, c);
            }
            PT_MARK(0x50, j, 0);
            lwait(bars, D2_FULL, (uint32_t)(j & 1));
            PT_MARK(0x51, j, 0);
            tc_fence_after();
            int last_p = -1;
            for (int p = 0; p < 8; p++)
                if ((8 * j + p) % GROUPS == grp) last_p = p;
#pragma unroll
            for (int p = 0; p < 8; p++) {
                if ((8 * j + p) % GROUPS != grp) continue;
                const uint32_t unit = (2 * p) % UNITS, par = (p / (UNITS / 2)) & 1u;
                uint32_t r0[16], r1[16];
                tmem_ld16(tq + TM_D2 + p * 16, r0);                            // osrc = 0: columns (2p + b1l)*8 + o2l
                tmem_ld16(tq + TM_D2 + 128 + p * 16, r1);                      // osrc = 1
                tmem_ld_wait();
                if (p == last_p) { tc_fence_before(); warp_arrive_leader(&bars[D2_EMPTY]); }
                PT_MARK(0x52, j, p);
                lwait(bars, (par == 0 ? U1_EMPTY0 : U2_EMPTY0) + unit, par ^ 1u);        // use 2 follows use 1 (step 2), use 3 follows use 2 (step 3)
                lwait(bars, (par == 0 ? U1_EMPTY0 : U2_EMPTY0) + unit + 1, par ^ 1u);
                PT_MARK(0x53, j, p);
                uint8_t* dst = smem + OFF_UNIT + unit * UNIT_BYTES;
#pragma unroll
                for (int cidx = 0; cidx < 16; cidx++) {
                    const int b1l = cidx >> 3, o2l = cidx & 7;
                    const uint32_t kbyte = (uint32_t)(b1l * 16 + i1_2) * 4u;
                    const uint32_t row0 = (uint32_t)(o2l * 8 + hb2 * 4 + q);
                    *reinterpret_cast<uint32_t*>(dst + sw128_off(row0, kbyte)) = r0[cidx];
                    *reinterpret_cast<uint32_t*>(dst + sw128_off(row0 + 64, kbyte)) = r1[cidx];
                }
                fence_async_smem();
                warp_arrive_leader(&bars[U2_FULL0 + unit]);
                warp_arrive_leader(&bars[U2_FULL0 + unit + 1]);
                PT_MARK(0x54, j, p);
            }
            // output: D3 lane m' = o2*8 + hb*4 + q' (o3 = 8 rank + 2q' + hb), column o1; the groups take the samples' outputs in turn
            if (j % GROUPS != grp) continue;
            PT_MARK(0x60, j, 0);
            lwait(bars, D3_FULL, (uint32_t)(j & 1));
            PT_MARK(0x61, j, 0);
            tc_fence_after();
            uint32_t acc[16];
            tmem_ld16(tq + TM_D3, acc);
            tmem_ld_wait();
            tc_fence_before();
            warp_arrive_leader(&bars[D3_EMPTY]);
            const int mrow = q * 32 + lane;
            const int o2 = mrow >> 3, o3 = 8 * (int)rank + 2 * (mrow & 3) + ((mrow >> 2) & 1);
            float* yo = y + (size_t)s * 4096 + o2 * 16 + o3;
            const float* bo = bias ? bias + o2 * 16 + o3 : nullptr;
#pragma unroll
            for (int o1 = 0; o1 < 16; o1++) {
                float v = __uint_as_float(acc[o1]) + (bo ? __ldg(bo + o1 * 256) : 0.0f);
                if (relu) v = fmaxf(v, 0.0f);
                yo[o1 * 256] = v;
            }
            PT_MARK(0x62, j, 0);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                                      // the peer's shared memory and TMEM are read by the leader's MMAs until here
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc2(tmem, TM_COLS);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int tt_dense3_tf32_pair(const float* x, const float* img, const float* bias, float* y, int batch, int relu, cudaStream_t st) {
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<EncodeTiledFn>(p);
    }
    SYN_REQUIRE(enc != nullptr, "syn_tt_dense3_tf32: cuTensorMapEncodeTiled is not available from this driver");
    // x[s][i1][i2][i3] as a rank-4 tensor (i3, i1, i2, s): one box = the i1-half of one sample, landing as 128 rows (i2, i1l) of 64 bytes
    CUtensorMap map;
    const cuuint64_t dims[4] = {16, 16, 16, (cuuint64_t)batch};
    const cuuint64_t strides[3] = {1024, 64, 16384};
    const cuuint32_t box[4] = {16, 8, 16, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SYN_REQUIRE(cr == CUDA_SUCCESS, "syn_tt_dense3_tf32: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    static PerDevice configured;
    const int dev_ = current_device();
    if (!configured.get(dev_)) {
        SYN_CUDA(cudaFuncSetAttribute(tt_dense3_tf32_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        configured.set(dev_);
    }
    int grid = sm_count() & ~1;
    if (const char* e = getenv("SYN_TT_GRID")) { int g = atoi(e) & ~1; if (g >= 2 && g < grid) grid = g; }
    if (grid > 2 * batch) grid = 2 * batch;
    tt_dense3_tf32_pair_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(map, img, bias, y, batch, relu);
    return launch_status("tt_dense3_tf32_pair_kernel");
}

int tt_pack_pair(const float* G1, float* img, cudaStream_t st) {
    tt_pack_pair_kernel<<<16, 256, 0, st>>>(G1, img);
    return launch_status("tt_pack_pair_kernel");
}

}  // namespace ttp
}  // namespace syn

#ifdef SYN_TT_DEBUG
extern "C" int syn_ttp_debug_buffer(uint32_t* device_ptr) {
    return cudaMemcpyToSymbol(syn::ttp::g_ttp_dbg, &device_ptr, sizeof(uint32_t*)) == cudaSuccess ? 0 : 1;
}
#endif
