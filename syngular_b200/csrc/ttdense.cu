// Fused TT-matvec of the MPO-compressed dense layer (TensorDense forward, reference layers/TensorDense.py:74-142) on the 5th-generation
// tensor cores: tcgen05.mma.kind::tf32 with TMEM accumulators, TMA-staged input, all intermediates on chip, bias + ReLU in the epilogue.
//
//   y[s,o1,o2,o3] = act( sum_{i1,i2,i3,b1,b2} x[s,i1,i2,i3] G1[i1,o1,b1] G2[i2,o2,b1,b2] G3[i3,o3,b2] + bias[o1,o2,o3] )
//
// for BASELINE configs[4]: three cores, every mode and bond = 16 (a 4096 -> 4096 layer held in 3 * 4096... 73,728 parameters).  Per sample
// the contraction is three GEMMs whose intermediates are 16x the sample (65,536 floats each): written to HBM they would cost 34 GB per
// 65,536-sample batch against 2 GB of input + output, so the whole chain runs inside one CTA per (sample, half of o2):
//
//   step 1   T1[(o3,b2), i1 | i2]        = sum_i3      G3^T[(o3,b2), i3]   x[(i2,i1), i3]             M=2x128  N=16   K=16   per i2
//   step 2   T2[(o3,i1), (b1,o2)]       += sum_(i2,b2) T1[(o3,i1), (i2,b2)] G2[(b1,o2), (i2,b2)]      M=2x128  N=128  K=256  (16 chunks of K=16)
//   step 3   Y [(o2,o3), o1]             = sum_(b1,i1) T2[(o2,o3), (b1,i1)] G1[o1, (b1,i1)]           M=128    N=16   K=256  (8 chunks of K=32)
//
// (the chain is contracted from the i3 end so that every operand, the input included, is K-major in memory).  In each step the
// contracted index pairs a bond produced on the weight side with a mode of the data side, so the accumulator of one step (TMEM:
// lane = M index, column = N index) is not the operand layout of the next: T1 and T2 are re-laid out TMEM -> registers -> shared
// memory by the four epilogue warps.  The index orders are chosen so that in BOTH re-layouts the contracted index of the next step is
// the TMEM lane: a warp's 32 lanes then write 32 consecutive words of one or two operand rows (conflict-free, one wavefront), in the
// UMMA canonical K-major swizzled layout the next tcgen05.mma reads through its shared-memory descriptor.
//
// Warp roles (64 + 128 GROUPS threads, one CTA per SM, persistent over its samples):
//   warp 0     TMA producer: weights once (pre-swizzled images, cp.async.bulk), then one 16 KB tensor-map load per sample
//              (cp.async.bulk.tensor.4d, 64B swizzle, box = the whole sample re-ordered to rows (i2,i1)); it also issues the step-1
//              MMAs, which run ahead of step 2 through NB1 accumulator buffers
//   warp 1     TMEM allocation + the step-2 / step-3 MMA issue (an independent instruction stream: a wait on an operand slot
//              never holds step 1 back)
//   warps 2..  epilogue, in GROUPS groups of four warps (one per TMEM lane quadrant) that take alternate chunks: tcgen05.ld -> st.shared
//              re-layouts, final tcgen05.ld + bias + ReLU + coalesced global stores
// Pipelines (mbarriers): x full/empty, D1 buffers full/empty, a ring of three 16 KB operand slots full/empty shared by the
// T1 and T2 re-layouts, D2 full/empty, D3 full/empty.
//
// Shared memory (one CTA per SM): G2 half image 128 KB + 3 operand slots 48 KB + G3 image 16 KB + G1 image 16 KB + x 16 KB = 224 KB.
// TMEM: D1 NB1 x 2 x 16 columns, D2 2 x 128, D3 16 -> 400 of 512 columns.
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"

namespace syn {

namespace tt {

constexpr int GROUPS = 2;                                       // epilogue groups of four warps
constexpr int THREADS = 64 + 128 * GROUPS + 64;                 // 2 issuer warps + epilogue warps + 2 more issuer warps
constexpr int WARP_S1B = 2 + 4 * GROUPS, WARP_S23B = WARP_S1B + 1;
constexpr int SLOTS = 3;
constexpr int NB1 = 4;                                          // D1 (step-1 accumulator) buffers: step 1 runs up to NB1 chunks ahead
constexpr uint32_t SLOT_BYTES = 16384;
// shared-memory map (offsets from a 1024-byte aligned base)
constexpr uint32_t OFF_B2 = 0;                                  // 8 atoms x [128 rows x 128 B]   (SW128, K-major)
constexpr uint32_t OFF_SLOT = OFF_B2 + 131072;                  // 3 x 16 KB
constexpr uint32_t OFF_A1 = OFF_SLOT + SLOTS * SLOT_BYTES;      // 2 tiles x [128 rows x 64 B]     (SW64)
constexpr uint32_t OFF_B3 = OFF_A1 + 16384;                     // 8 atoms x [16 rows x 128 B]     (SW128)
constexpr uint32_t OFF_X = OFF_B3 + 16384;                      // 256 rows x 64 B                 (SW64, written by TMA)
constexpr uint32_t OFF_BAR = OFF_X + 16384;                     // mbarriers + TMEM base pointer
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256 + 1024;           // + slack for the 1024-byte alignment of the base
// packed weight image in global memory (floats): [A1 16 KB][B3 16 KB][B2 half 0 128 KB][B2 half 1 128 KB]
constexpr size_t IMG_A1 = 0, IMG_B3 = 4096, IMG_B2 = 8192, IMG_FLOATS = 8192 + 2 * 32768 + 4096;      // + the G1 halves of the CTA-pair kernel
// TMEM columns
constexpr uint32_t TM_D1 = 0, TM_D2 = 32 * NB1, TM_D3 = TM_D2 + 256, TM_COLS = 512;      // D3: two partial accumulators of 16 columns
static_assert(TM_D3 + 32 <= TM_COLS, "TMEM columns");

enum Bar { W_FULL = 0, X_FULL, X_EMPTY, D1_FULL0, D1_EMPTY0 = D1_FULL0 + NB1, SLOT_FULL0 = D1_EMPTY0 + NB1, SLOT_EMPTY0 = SLOT_FULL0 + SLOTS,
           D2_FULL = SLOT_EMPTY0 + SLOTS, D2_EMPTY, D3_FULL, D3_EMPTY, NUM_BARS };
static_assert(NB1 % GROUPS == 0, "a D1 buffer must always be drained by the same epilogue group");
// mbarrier waits are by parity: a waiter may run at most one phase ahead.  An epilogue group takes every GROUPS-th chunk and its wait on
// a slot proves that the chunk SLOTS uses earlier was consumed, so it stays within one phase only if GROUPS <= SLOTS.
static_assert(GROUPS <= SLOTS, "epilogue groups could run two phases ahead of an operand slot");
// The issue loops are fully unrolled so that every ring position is a compile-time constant (the issuing warp's instruction stream,
// not the tensor pipe, paces this kernel: a group of four tcgen05.mma with run-time descriptors costs ~460 cycles of R2UR moves): a
// sample must use every D1 buffer and every operand slot an EVEN number of times, so that neither the position nor the barrier parity
// depends on the sample index.
static_assert(16 % NB1 == 0 && (16 / NB1) % 2 == 0, "D1 buffer position / parity must not depend on the sample");
static_assert(24 % SLOTS == 0 && (24 / SLOTS) % 2 == 0, "operand slot position / parity must not depend on the sample");

// byte offset of element (row, kbyte) in a K-major operand block with 128-byte rows / 128B swizzle, resp. 64-byte rows / 64B swizzle
__host__ __device__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t kbyte) {
    return (row >> 3) * 1024u + (row & 7u) * 128u + ((((kbyte >> 4) ^ (row & 7u)) & 7u) << 4) + (kbyte & 15u);
}
__host__ __device__ __forceinline__ uint32_t sw64_off(uint32_t row, uint32_t kbyte) {
    return (row >> 3) * 512u + (row & 7u) * 64u + ((((kbyte >> 4) ^ ((row & 7u) >> 1)) & 3u) << 4) + (kbyte & 15u);
}

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one lane of a converged warp (the same one every time): the warp stays uniform around the single-thread tcgen05 instructions, so
// the compiler keeps descriptors and addresses in uniform registers instead of moving them there lane by lane
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// all tcgen05 operations issued so far by this thread arrive on `bar` when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[N]) {
    static_assert(N == 16 || N == 8 || N == 4, "tmem_ld: 4, 8 or 16 columns");
    if constexpr (N == 16) {
        tmem_ld16(taddr, r);
    } else if constexpr (N == 4) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr) : "memory");
    }
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}

// bounded wait with a diagnosis: a pipeline bug prints which barrier starved instead of hanging the GPU
__device__ __noinline__ void tt_wait_timeout(int id, uint32_t parity) {
    printf("tt_dense3_tf32_kernel: block %d thread %d starved on barrier %d (parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, id, parity);
    __trap();
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tt_wait(uint64_t* bars, int id, uint32_t parity) {
    if (mbar_try_wait(&bars[id], parity)) return;
    const uint64_t t0 = global_ns();
    for (unsigned spin = 1; !mbar_try_wait(&bars[id], parity); ++spin)
        if ((spin & 255u) == 0 && global_ns() - t0 > 2000000000ull) tt_wait_timeout(id, parity);      // 2 s: far beyond any legitimate wait
}

#ifdef SYN_TT_DEBUG
// timeline trace of CTA 0 (compile with -DSYN_TT_DEBUG): every mark appends (value, clock) to its warp's lane of a device buffer
__device__ uint32_t* g_tt_dbg = nullptr;      // [warp][TT_TRACE_LEN][2]
constexpr int TT_TRACE_LEN = 8192;
#define TT_TRACE_DECL uint32_t tt_cnt_ = 0;
#define TT_MARK(slot_, val_)                                                                            \
    do {                                                                                                \
        if (g_tt_dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && tt_cnt_ < TT_TRACE_LEN) {         \
            uint32_t* w_ = g_tt_dbg + ((size_t)(threadIdx.x >> 5) * TT_TRACE_LEN + tt_cnt_) * 2;        \
            w_[0] = (uint32_t)(val_);                                                                   \
            w_[1] = (uint32_t)clock64();                                                                \
            ++tt_cnt_;                                                                                  \
        }                                                                                               \
    } while (0)
#else
#define TT_TRACE_DECL
#define TT_MARK(slot_, val_) do { } while (0)
#endif

// K-major shared-memory matrix descriptor (sm_100 "version 1"): start address, stride between 8-row groups, swizzle mode
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)layout_type << 61);
}
constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_SW64 = 4;
// instruction descriptor of kind::tf32: FP32 accumulate, TF32 A and B, both K-major, shape M x N (K = 8 per instruction)
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- weight packing: the three cores -> the pre-swizzled operand images ------------------------------------------------------------
// G1 (i1,o1,b1), G2 (i2,o2,b1,b2), G3 (i3,o3,b2): the reference's layouts (layers/TensorDense.py:50-71), every extent 16.
__global__ void __launch_bounds__(256) tt_pack_kernel(const float* __restrict__ G1, const float* __restrict__ G2, const float* __restrict__ G3,
                                                      float* __restrict__ img) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    // A1[t][row = o3l*16 + b2][k = i3] = G3[i3, 8t + o3l, b2]                                      (SW64; 2 tiles of 8 KB)
    for (int e = tid; e < 4096; e += stride) {
        const int i3 = e & 15, row = (e >> 4) & 127, t = e >> 11;
        const int o3 = 8 * t + (row >> 4), b2 = row & 15;
        img[IMG_A1 + (t * 8192 + sw64_off(row, i3 * 4)) / 4] = G3[(i3 * 16 + o3) * 16 + b2];
    }
    // B3[atom a][row = o1][kk = (b1 & 1)*16 + i1], b1 = 2a + (kk >> 4)  = G1[i1, o1, b1]              (SW128; 8 atoms of 2 KB)
    for (int e = tid; e < 4096; e += stride) {
        const int kk = e & 31, o1 = (e >> 5) & 15, a = e >> 9;
        const int b1 = 2 * a + (kk >> 4), i1 = kk & 15;
        img[IMG_B3 + (a * 2048 + sw128_off(o1, kk * 4)) / 4] = G1[(i1 * 16 + o1) * 16 + b1];
    }
    // B2[h][atom a][row n = b1*8 + o2l][kk = (i2 & 1)*16 + b2], i2 = 2a + (kk >> 4)  = G2[i2, 8h + o2l, b1, b2]   (SW128; 8 atoms of 16 KB)
    for (int e = tid; e < 65536; e += stride) {
        const int kk = e & 31, n = (e >> 5) & 127, a = (e >> 12) & 7, h = e >> 15;
        const int i2 = 2 * a + (kk >> 4), b2 = kk & 15, b1 = n >> 3, o2 = 8 * h + (n & 7);
        img[IMG_B2 + (size_t)h * 32768 + (a * 16384 + sw128_off(n, kk * 4)) / 4] = G2[((i2 * 16 + o2) * 16 + b1) * 16 + b2];
    }
}

// ---- the fused forward kernel ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 1)
tt_dense3_tf32_kernel(const __grid_constant__ CUtensorMap xmap, const float* __restrict__ img, const float* __restrict__ bias,
                      float* __restrict__ y, int batch, int relu) {
    extern __shared__ uint8_t tt_smem_raw[];
    // 1024-byte aligned base: SW128 atoms and the descriptors' swizzle phases assume it
    const uint32_t raw = smem_u32(tt_smem_raw);
    uint8_t* smem = tt_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NUM_BARS * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TT_TRACE_DECL
    const int h = blockIdx.x & 1;                                   // which half of o2 this CTA computes
    const int first = blockIdx.x >> 1, step = gridDim.x >> 1;       // its samples: first, first + step, ...
    const int my_samples = first < batch ? (batch - first + step - 1) / step : 0;

    if (threadIdx.x == 0) {
        mbar_init(&bars[W_FULL], 1);
        mbar_init(&bars[X_FULL], 1);
        mbar_init(&bars[X_EMPTY], 2);
        for (int b = 0; b < NB1; b++) {
            mbar_init(&bars[D1_FULL0 + b], 2);
            mbar_init(&bars[D1_EMPTY0 + b], 128);
        }
        for (int s = 0; s < SLOTS; s++) {
            mbar_init(&bars[SLOT_FULL0 + s], 128);
            mbar_init(&bars[SLOT_EMPTY0 + s], 2);
        }
        mbar_init(&bars[D2_FULL], 2);
        mbar_init(&bars[D2_EMPTY], 128 * GROUPS);
        mbar_init(&bars[D3_FULL], 2);
        mbar_init(&bars[D3_EMPTY], 128);
        fence_async_smem();
    }
    TT_MARK(2, 1);
    if (warp == 1) tmem_alloc(tmem_ptr, TM_COLS);
    TT_MARK(2, 2);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 0 || warp == WARP_S1B) {
        // ===== step-1 issuers: warp 0 (tile 0; it is also the TMA producer) and warp WARP_S1B (tile 1).  Each issuing warp's instruction
        // stream (descriptor moves into uniform registers, ~100 cycles per tcgen05.mma), not the tensor pipe, bounds the issue rate, so
        // every MMA group is split over two warps by accumulator tile.  The whole warp runs the loop; one elected lane issues. =========
        const int tile = warp == 0 ? 0 : 1;
        constexpr uint32_t ID_S1 = idesc_tf32(128, 16);
        const uint64_t d64 = smem_desc(0, 512, LAYOUT_SW64);
        const uint64_t a1_desc = d64 + ((sbase + OFF_A1) >> 4), x_desc = d64 + ((sbase + OFF_X) >> 4);
        if (warp == 0 && elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
            // weights: A1 (16 KB), B3 (16 KB), this CTA's half of B2 (128 KB in 8 copies)
            mbar_expect_tx(&bars[W_FULL], 16384u + 16384u + 131072u);
            bulk_g2s(smem + OFF_A1, img + IMG_A1, 16384u, &bars[W_FULL]);
            bulk_g2s(smem + OFF_B3, img + IMG_B3, 16384u, &bars[W_FULL]);
            for (int a = 0; a < 8; a++)
                bulk_g2s(smem + OFF_B2 + a * 16384, img + IMG_B2 + (size_t)h * 32768 + (size_t)a * 4096, 16384u, &bars[W_FULL]);
        }
        __syncwarp();
        tt_wait(bars, W_FULL, 0);
        for (int j = 0; j < my_samples; j++) {
            TT_MARK(0, 0x100 + j);
            if (warp == 0) {
                tt_wait(bars, X_EMPTY, (uint32_t)(j & 1) ^ 1u);      // a fresh barrier passes a wait on the "previous" phase
                if (elect_one()) {
                    mbar_expect_tx(&bars[X_FULL], 16384u);
                    tma_load_4d(smem + OFF_X, &xmap, 0, 0, 0, first + j * step, &bars[X_FULL]);
                }
                __syncwarp();
            }
            tt_wait(bars, X_FULL, (uint32_t)(j & 1));
            // step 1, chunk c (= i2): D1[c % NB1][t] = A1[t] (128 x 16) . x[(i2 = c, i1), i3]^T (16 x 16); runs ahead of step 2 as far as
            // the D1 buffers allow
#pragma unroll
            for (int c = 0; c < 16; c++) {
                const uint32_t b = c % NB1, use = c / NB1;           // position and parity are the same in every sample (static_asserts above)
                TT_MARK(0, 0x10000 + j * 256 + c);
                tt_wait(bars, D1_EMPTY0 + b, (use & 1u) ^ 1u);
                TT_MARK(0, 0x11000 + j * 256 + c);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t xd = x_desc + (uint64_t)(c * 64);
#pragma unroll
                    for (int k = 0; k < 2; k++)
                        umma_tf32(tmem + TM_D1 + b * 32 + tile * 16, a1_desc + (uint64_t)(tile * 512 + k * 2), xd + (uint64_t)(k * 2), ID_S1, k);
                    umma_commit(&bars[D1_FULL0 + b]);
                    if (c == 15) umma_commit(&bars[X_EMPTY]);
                }
                __syncwarp();
                TT_MARK(0, 0x12000 + j * 256 + c);
            }
        }
    } else if (warp == 1 || warp == WARP_S23B) {
        // ===== step-2 / step-3 issuers: warp 1 takes accumulator tile 0 of step 2 and the first two k-steps of every step-3 chunk, warp
        // WARP_S23B tile 1 and the last two k-steps (two partial D3 accumulators, added by the epilogue) ======================================
        const int tile = warp == 1 ? 0 : 1;
        // Every ring position has a running use count u (24 operand-slot uses per sample: 16 T1 chunks, 8 T2 chunks):
        // slot = u % SLOTS, and the parity to wait for is (u / SLOTS) & 1 on a "full" barrier, the opposite on an "empty" one.
        {
            constexpr uint32_t ID_S2 = idesc_tf32(128, 128), ID_S3 = idesc_tf32(128, 16);
            // descriptor templates: everything but the start address; a k-step of 8 TF32 (32 bytes) adds 2 to the address field
            const uint64_t d64 = smem_desc(0, 512, LAYOUT_SW64), d128 = smem_desc(0, 1024, LAYOUT_SW128);
            const uint64_t b2_desc = d128 + ((sbase + OFF_B2) >> 4), b3_desc = d128 + ((sbase + OFF_B3) >> 4);
            const uint64_t slot64_desc = d64 + ((sbase + OFF_SLOT) >> 4), slot128_desc = d128 + ((sbase + OFF_SLOT) >> 4);
            tt_wait(bars, W_FULL, 0);
            for (int j = 0; j < my_samples; j++) {
                // step 2, chunk cc: D2[t] += T1 chunk (slot: 2 tiles of 128 x 16) . G2 half [(b1,o2l), (i2 = cc, b2)]^T
#pragma unroll
                for (int cc = 0; cc < 16; cc++) {
                    const uint32_t u = (uint32_t)cc, slot = u % SLOTS;          // compile-time ring position and parity
                    TT_MARK(0, 0x20000 + j * 256 + cc);
                    tt_wait(bars, SLOT_FULL0 + slot, (u / SLOTS) & 1u);
                    if (cc == 0) tt_wait(bars, D2_EMPTY, (uint32_t)(j & 1) ^ 1u);
                    TT_MARK(0, 0x21000 + j * 256 + cc);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad = slot64_desc + (uint64_t)(slot * (SLOT_BYTES >> 4));
                        const uint64_t bd = b2_desc + (uint64_t)((cc >> 1) * 1024 + (cc & 1) * 4);
#pragma unroll
                        for (int k = 0; k < 2; k++)
                            umma_tf32(tmem + TM_D2 + tile * 128, ad + (uint64_t)(tile * 512 + k * 2), bd + (uint64_t)(k * 2), ID_S2,
                                      (cc > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&bars[SLOT_EMPTY0 + slot]);
                        if (cc == 15) umma_commit(&bars[D2_FULL]);
                    }
                    __syncwarp();
                    TT_MARK(0, 0x22000 + j * 256 + cc);
                }
                // step 3, chunk p (= b1 pair): D3 += T2 chunk (slot: 128 x 32) . G1 [o1, (b1, i1)]^T
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    const uint32_t u = (uint32_t)(16 + p), slot = u % SLOTS;
                    TT_MARK(0, 0x30000 + j * 256 + p);
                    tt_wait(bars, SLOT_FULL0 + slot, (u / SLOTS) & 1u);
                    if (p == 0) tt_wait(bars, D3_EMPTY, (uint32_t)(j & 1) ^ 1u);
                    TT_MARK(0, 0x31000 + j * 256 + p);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad = slot128_desc + (uint64_t)(slot * (SLOT_BYTES >> 4));
                        const uint64_t bd = b3_desc + (uint64_t)(p * 128);
#pragma unroll
                        for (int k = 0; k < 2; k++)
                            umma_tf32(tmem + TM_D3 + tile * 16, ad + (uint64_t)((2 * tile + k) * 2), bd + (uint64_t)((2 * tile + k) * 2), ID_S3,
                                      (p > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&bars[SLOT_EMPTY0 + slot]);
                        if (p == 7) umma_commit(&bars[D3_FULL]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue warps: re-layouts and the output ===================================================================================
        // GROUPS groups of four warps (one warp per TMEM lane quadrant); group g takes the chunks c = g (mod GROUPS) of both re-layouts, so
        // the tcgen05.ld -> st.shared -> fence latency chains of consecutive chunks overlap, and GROUPS-th of the output columns.
        const int grp = (warp - 2) >> 2;
        const int q = warp & 3;                                      // TMEM lane quadrant this warp may access
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        const int hb = lane >> 4, lo = lane & 15;
        // re-layout 1: D1 lane = (o3l = 2q + hb, b2 = lo), column i1  ->  A2 slot, tile t, row m = q*32 + i1*2 + hb, k = b2        (SW64)
        //   row m: group (m >> 3) = q*4 + (i1 >> 2), in-group (m & 7) = (i1 & 3)*2 + hb, swizzle phase = i1 & 3
        const uint32_t r1_base = (uint32_t)(q * 4) * 512u + (uint32_t)hb * 64u + (uint32_t)(lo & 3) * 4u;
        // re-layout 2: D2 lane = (o3l = 2q + hb', i1), lane index L = i1*2 + hb' within the quadrant; column = b1l*8 + o2l of the chunk
        //   ->  A3 slot row m' = o2l*16 + t*8 + hb'*4 + q, k = b1l*16 + i1                                                         (SW128)
        const int hb2 = lane & 1, i1_2 = lane >> 1;
        for (int j = 0; j < my_samples; j++) {
            const int s = first + j * step;
#pragma unroll
            for (int c = 0; c < 16; c++) {
                if ((16 * j + c) % GROUPS != grp) continue;          // running chunk count: the groups take the chunks in turn
                const uint32_t b = c % NB1, use = c / NB1;
                const uint32_t u = (uint32_t)c, slot = u % SLOTS;
                TT_MARK(0, 0x40000 + j * 256 + c);
                tt_wait(bars, D1_FULL0 + b, use & 1u);
                TT_MARK(0, 0x41000 + j * 256 + c);
                tc_fence_after();
                uint32_t r0[16], r1[16];
                tmem_ld16(tq + TM_D1 + b * 32, r0);
                tmem_ld16(tq + TM_D1 + b * 32 + 16, r1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bars[D1_EMPTY0 + b]);
                TT_MARK(0, 0x50000 + j * 256 + c);
                tt_wait(bars, SLOT_EMPTY0 + slot, ((u / SLOTS) & 1u) ^ 1u);
                TT_MARK(0, 0x51000 + j * 256 + c);
                uint8_t* dst = smem + OFF_SLOT + slot * SLOT_BYTES + r1_base;
#pragma unroll
                for (int i1 = 0; i1 < 16; i1++) {
                    const uint32_t off = (uint32_t)(i1 >> 2) * 512u + (uint32_t)(i1 & 3) * 128u + ((((uint32_t)(lo >> 2)) ^ (uint32_t)(i1 & 3)) << 4);
                    *reinterpret_cast<uint32_t*>(dst + off) = r0[i1];
                    *reinterpret_cast<uint32_t*>(dst + 8192 + off) = r1[i1];
                }
                TT_MARK(0, 0x52000 + j * 256 + c);
                fence_async_smem();
                mbar_arrive(&bars[SLOT_FULL0 + slot]);
                TT_MARK(0, 0x53000 + j * 256 + c);
            }
            TT_MARK(0, 0x60000 + j * 256);
            tt_wait(bars, D2_FULL, (uint32_t)(j & 1));
            TT_MARK(0, 0x61000 + j * 256);
            tc_fence_after();
            int last_p = -1;                                         // this group's last T2 chunk of the sample: its last read of D2
            for (int p = 0; p < 8; p++)
                if ((uint32_t)(8 * j + p) % GROUPS == (uint32_t)grp) last_p = p;
#pragma unroll
            for (int p = 0; p < 8; p++) {
                if ((uint32_t)(8 * j + p) % GROUPS != (uint32_t)grp) continue;
                const uint32_t u = (uint32_t)(16 + p), slot = u % SLOTS;
                uint32_t r0[16], r1[16];
                tmem_ld16(tq + TM_D2 + p * 16, r0);
                tmem_ld16(tq + TM_D2 + 128 + p * 16, r1);
                tmem_ld_wait();
                if (p == last_p) { tc_fence_before(); mbar_arrive(&bars[D2_EMPTY]); }
                TT_MARK(0, 0x70000 + j * 256 + p);
                tt_wait(bars, SLOT_EMPTY0 + slot, ((u / SLOTS) & 1u) ^ 1u);
                uint8_t* dst = smem + OFF_SLOT + slot * SLOT_BYTES;
#pragma unroll
                for (int cidx = 0; cidx < 16; cidx++) {
                    const int b1l = cidx >> 3, o2l = cidx & 7;
                    const uint32_t kbyte = (uint32_t)(b1l * 16 + i1_2) * 4u;
                    const uint32_t row0 = (uint32_t)(o2l * 16 + hb2 * 4 + q);
                    *reinterpret_cast<uint32_t*>(dst + sw128_off(row0, kbyte)) = r0[cidx];
                    *reinterpret_cast<uint32_t*>(dst + sw128_off(row0 + 8, kbyte)) = r1[cidx];
                }
                fence_async_smem();
                mbar_arrive(&bars[SLOT_FULL0 + slot]);
            }
            // output: D3 lane m' = o2l*16 + t*8 + hb*4 + q', column o1; the groups take the samples' outputs in turn
            if (j % GROUPS != grp) continue;
            TT_MARK(0, 0x80000 + j * 256);
            tt_wait(bars, D3_FULL, (uint32_t)(j & 1));
            TT_MARK(0, 0x81000 + j * 256);
            tc_fence_after();
            uint32_t acc[16], acc2[16];
            tmem_ld16(tq + TM_D3, acc);
            tmem_ld16(tq + TM_D3 + 16, acc2);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars[D3_EMPTY]);
            const int mrow = q * 32 + lane;
            const int o2 = 8 * h + (mrow >> 4), o3 = 8 * ((mrow >> 3) & 1) + 2 * (mrow & 3) + ((mrow >> 2) & 1);
            float* yo = y + (size_t)s * 4096 + o2 * 16 + o3;
            const float* bo = bias ? bias + o2 * 16 + o3 : nullptr;
#pragma unroll
            for (int o1 = 0; o1 < 16; o1++) {
                float v = __uint_as_float(acc[o1]) + __uint_as_float(acc2[o1]) + (bo ? __ldg(bo + o1 * 256) : 0.0f);
                if (relu) v = fmaxf(v, 0.0f);
                yo[o1 * 256] = v;
            }
        }
    }
    TT_MARK(0, 0x90000);
    tc_fence_before();
    __syncthreads();
    TT_MARK(0, 0xA0000);
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem, TM_COLS);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;              // the driver entry point is process-wide, not per device
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int tt_dense3_tf32(const float* x, const float* img, const float* bias, float* y, int batch, int relu, cudaStream_t st) {
    SYN_REQUIRE(x && img && y && batch >= 0, "syn_tt_dense3_tf32: null argument");
    if (batch == 0) return 0;
    SYN_REQUIRE(((((uintptr_t)x) | ((uintptr_t)img) | ((uintptr_t)y)) & 15) == 0, "syn_tt_dense3_tf32: x, weights and y must be 16-byte aligned");
    EncodeTiledFn enc = encode_tiled_fn();
    SYN_REQUIRE(enc != nullptr, "syn_tt_dense3_tf32: cuTensorMapEncodeTiled is not available from this driver");
    // x[s][i1][i2][i3] seen as a rank-4 tensor with dimensions (i3, i1, i2, s): one box = one sample, landing as rows (i2, i1) of 64 bytes
    CUtensorMap map;
    const cuuint64_t dims[4] = {16, 16, 16, (cuuint64_t)batch};
    const cuuint64_t strides[3] = {1024, 64, 16384};               // bytes, dimensions 1..3
    const cuuint32_t box[4] = {16, 16, 16, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SYN_REQUIRE(cr == CUDA_SUCCESS, "syn_tt_dense3_tf32: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    static PerDevice configured;
    const int dev_ = current_device();
    if (!configured.get(dev_)) {
        SYN_CUDA(cudaFuncSetAttribute(tt_dense3_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        configured.set(dev_);
    }
    int grid = sm_count() & ~1;                                     // CTA pairs (o2 halves) share the samples
    if (const char* e = getenv("SYN_TT_GRID")) { int g = atoi(e) & ~1; if (g >= 2 && g < grid) grid = g; }   // experiment knob
    if (grid > 2 * batch) grid = 2 * batch;
    tt_dense3_tf32_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(map, img, bias, y, batch, relu);
    return launch_status("tt_dense3_tf32_kernel");
}

}  // namespace tt
namespace ttp {
int tt_dense3_tf32_pair(const float* x, const float* img, const float* bias, float* y, int batch, int relu, cudaStream_t st);
int tt_pack_pair(const float* G1, float* img, cudaStream_t st);
}  // namespace ttp
}  // namespace syn

#ifdef SYN_TT_DEBUG
extern "C" int syn_tt_debug_buffer(uint32_t* host_mapped_device_ptr) {
    return cudaMemcpyToSymbol(syn::tt::g_tt_dbg, &host_mapped_device_ptr, sizeof(uint32_t*)) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" size_t syn_tt_dense3_packed_floats(void) { return syn::tt::IMG_FLOATS; }

extern "C" int syn_tt_dense3_pack_tf32(const float* G1, const float* G2, const float* G3, float* packed, void* stream) {
    using namespace syn;
    SYN_REQUIRE(G1 && G2 && G3 && packed, "syn_tt_dense3_pack_tf32: null argument");
    tt::tt_pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(G1, G2, G3, packed);
    if (int rc = launch_status("tt_pack_kernel")) return rc;
    return ttp::tt_pack_pair(G1, packed, (cudaStream_t)stream);
}

static int tt_check_args(const float* x, const float* packed, const float* y) {
    using namespace syn;
    SYN_REQUIRE(x && packed && y, "syn_tt_dense3_tf32: null argument");
    SYN_REQUIRE(((((uintptr_t)x) | ((uintptr_t)packed) | ((uintptr_t)y)) & 15) == 0, "syn_tt_dense3_tf32: x, weights and y must be 16-byte aligned");
    return 0;
}

extern "C" int syn_tt_dense3_tf32(const float* x, const float* packed, const float* bias, float* y, int batch, int relu, void* stream) {
    return syn::tt::tt_dense3_tf32(x, packed, bias, y, batch, relu, (cudaStream_t)stream);
}

// the same layer on CTA pairs (tcgen05 cta_group::2, csrc/ttdense_pair.cu): same arguments, same results to rounding order
extern "C" int syn_tt_dense3_tf32_pair(const float* x, const float* packed, const float* bias, float* y, int batch, int relu, void* stream) {
    if (batch <= 0) return 0;
    if (int rc = tt_check_args(x, packed, y)) return rc;
    return syn::ttp::tt_dense3_tf32_pair(x, packed, bias, y, batch, relu, (cudaStream_t)stream);
}
