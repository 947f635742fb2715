// sgmm_spec256.cu -- H = 256 policy rollout on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// BASELINE.json config 4 ("wider MLP hidden sizes (256x256) on tensor cores").  Per individual the
// sequential rollout is a chain of 256x256 GEMVs -- no batch dimension, the tensor cores would
// idle.  The inventory is a 5-valued integer ({-2..2}, Env/market_env.py:14-15,34-35), so the
// action at bar t is a function of (t, inv) only: evaluate the policy for ALL five
// inventories of every bar (time-parallel speculation) and the hidden layer becomes a real GEMM
//     D[(t,inv), j] = sum_k h1[(t,inv), k] * W2[j, k]        M = 5*T rows, N = K = 256
// followed by a trivially cheap walk of the 5-state automaton.  5x the hidden-layer FLOPs, on a
// pipe ~30x faster than the FP32 cores, and no step-to-step latency chain.
//
// One persistent CTA per SM, one individual at a time, warp-specialised:
//   warps 0-3  epilogue : tcgen05.ld of the fp32 accumulator (lane = row), +b2, ReLU, layer 3
//                         (256->2, thread-local), x5 + round-half-even, and the SPECULATIVE env step of
//                         the row's (bar, inventory): fills, next inventory, fp64 reward
//                         (Env/market_env.py:30-58) -> 32-byte table entry
//   warps 4-7  producer : layer 1 (3->256) in fp32 SGMM order for the 5 inventories of 25 bars,
//                         cvt.rn.relu.bf16x2, 128B-swizzled K-major A tile (4 k-blocks of 16 KB, ring)
//   warp  8    MMA      : one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//                         (M=128, N=256, K=16) x16 per tile, bf16 x bf16 -> fp32 in TMEM
//                         (2 x 256 columns, double-buffered); tcgen05.commit -> mbarriers
//   warp  9    walker   : inv <- next[t][inv], total += reward[t][inv] (fp64, reference order),
//                         trades; fitness / trade count out (Env/drl_engine.py:54-67)
// W2 (bf16, 128 KB, swizzled K-major B operand) stays resident in shared memory for the episode.
//
// Precision: h1 and W2 are rounded to bf16 (fp32 accumulate), so policy outputs differ from the
// fp32 oracle by ~1e-3 of a tick; tests/test_gpu_spec256.py states the tolerance, checks that no
// rounding flips where the oracle's margin exceeds it, and checks that GIVEN the kernel's actions
// every integer and fp64 quantity is bit-identical to the oracle (teacher-forced replay).
#include <cuda_bf16.h>
#include "sgmm_internal.h"
#include "sgmm_rng.cuh"
#include "sgmm_step_core.h"

namespace sgmm {

namespace s256 {

constexpr int H = 256;
constexpr int TILE_ROWS = 128;
constexpr int TILE_BARS = 25;                 // 25 bars x 5 inventories = 125 rows (+3 idle rows)
constexpr int KBLK = 64;                      // bf16 elements per 128-byte swizzle row
constexpr int NKB = H / KBLK;                 // 4 k-blocks
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8, NUM_PROD_WARPS = 4;      // epilogue: 2 column halves x 4 TMEM lane quarters
constexpr int WARP_MMA = 12, WARP_WALK = 13;
constexpr int NUM_THREADS = 448;
constexpr uint32_t TMEM_COLS = 512;
constexpr int64_t G = (int64_t)H * H + 7 * H + 2;       // 67330

struct __align__(32) TableEntry {
    double reward;          // reward of taking this row's action from this inventory (market_env.py:58)
    int32_t ka, kb;         // quantised offsets (drl_engine.py:39)
    int32_t next;           // inventory index (inv+2) after the step
    int32_t traded;         // 1 if any side filled (drl_engine.py:60-61)
    float raw_a, raw_b;     // policy outputs (audit)
};

struct Smem {
    uint8_t b_tile[NKB][H * 128];              // W2 bf16, [k-block][n][64] swizzled, 128 KB
    uint8_t a_tile[NKB][TILE_ROWS * 128];      // h1 bf16, [k-block][row][64] swizzled, 64 KB
    float w1x[H], w1y[H], w1i[H], b1[H], b2[H], w3a[H], w3b[H];
    float b3[4];
    TableEntry table[2][TILE_ROWS];
    float2 part[2][TILE_ROWS];                 // layer-3 partial sums of the upper column half
    uint64_t a_full[NKB], a_empty[NKB], d_full[2], d_empty[2], t_full[2], t_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE_%=;\n"
        "bra LAB_WAIT_%=;\n"
        "LAB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major) | [32,46) SBO>>4 = 1024 B (8 rows)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 (bit 4), a=b=bf16 (bits 7,10), K-major both,
// n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29)
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(TILE_ROWS >> 4) << 24);

__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// two fp32 -> packed bf16x2 with ReLU; `lo` lands in the low half (lower address)
__device__ __forceinline__ uint32_t pack_relu_bf16(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// byte offset of 16-byte chunk `c` (8 bf16) of row `r` inside a [rows][128 B] SWIZZLE_128B slab
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

struct Args {
    const BarSig* sig; const BarPx* px; int64_t T;
    double tick, phi, fee;
    PopArgs mm;
    double* fitness; int32_t* trades;
    float* raw_table;        // optional audit output [P][T][5][2]
    int32_t* act_trace;      // optional audit output [P][T][2] : actions actually taken
};

template <bool FEE>
__global__ void __launch_bounds__(NUM_THREADS, 1) spec256_kernel(const Args a)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B operands need 1024-byte aligned slabs (descriptor base_offset = 0)
    const uint32_t misalign = smem_u32(smem_raw) & 1023u;
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((1024u - misalign) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t T = a.T;
    const int64_t ntiles = (T + TILE_BARS - 1) / TILE_BARS;

    if (tid == 0) {
        for (int i = 0; i < NKB; ++i) { mbar_init(&sm.a_full[i], NUM_PROD_WARPS); mbar_init(&sm.a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.d_full[i], 1); mbar_init(&sm.d_empty[i], NUM_EPI_WARPS);
            mbar_init(&sm.t_full[i], 4); mbar_init(&sm.t_empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // rows 125..127 of the A tile are never produced: keep them finite
    for (int i = tid; i < NKB * TILE_ROWS * 128 / 16; i += NUM_THREADS)
        reinterpret_cast<uint4*>(&sm.a_tile[0][0])[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    const PopArgs pop = a.mm;
    uint32_t gt = 0;                               // tiles processed so far by this CTA (pipeline phase source)

    for (int64_t ind = blockIdx.x; ind < pop.count; ind += gridDim.x) {
        // ---------------- stage the individual's weights (all warps) ----------------------------
        {
            GenomeSource src;
            src.seeded = (pop.genomes == nullptr);
            src.row = src.seeded ? pop.master : pop.genomes + ind * G;
            src.sigma = pop.sigma; src.seed = pop.seed; src.generation = pop.generation;
            src.individual = (uint64_t)(pop.first_index + ind);
            const int64_t offW2 = 4 * H;                      // W1[H,3] | b1[H] | W2[H,H] | b2 | W3[2,H] | b3[2]
            for (int q = tid; q < H * H / 8; q += NUM_THREADS) {        // 8 consecutive k of one row j
                const int j = q / (H / 8), kc = q % (H / 8);           // kc: 16-byte chunk along k (0..31)
                float v[4], w[4];
                src.at4(offW2 + (int64_t)j * H + kc * 8, v);
                src.at4(offW2 + (int64_t)j * H + kc * 8 + 4, w);
                uint4 o;
                o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
                o.z = pack_bf16(w[0], w[1]); o.w = pack_bf16(w[2], w[3]);
                *reinterpret_cast<uint4*>(&sm.b_tile[kc >> 3][swz(j, kc & 7)]) = o;
            }
            for (int j = tid; j < H; j += NUM_THREADS) {
                sm.w1x[j] = src.at(3 * j); sm.w1y[j] = src.at(3 * j + 1); sm.w1i[j] = src.at(3 * j + 2);
                sm.b1[j] = src.at(3 * H + j);
                sm.b2[j] = src.at(4 * H + (int64_t)H * H + j);
                sm.w3a[j] = src.at(5 * H + (int64_t)H * H + j);
                sm.w3b[j] = src.at(6 * H + (int64_t)H * H + j);
            }
            if (tid < 2) sm.b3[tid] = src.at(7 * H + (int64_t)H * H + tid);
        }
        fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncthreads();

        if (warp < NUM_EPI_WARPS) {
            // =========================== EPILOGUE ==============================================
            const int half = warp >> 2, quarter = warp & 3;      // column half, TMEM lane quarter (= warp % 4)
            const int row = quarter * 32 + lane;
            const int tl = row / 5, iv = row % 5;          // bar within the tile, inventory index (inv+2)
            const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
            const int col0 = half * (H / 2);
            for (int64_t it = 0; it < ntiles; ++it) {
                const uint32_t g = gt + (uint32_t)it, buf = g & 1u, use = g >> 1;
                // bar data of this row's step: issued before the accumulator is even ready so that the
                // L2 latency hides behind the MMA wait and the column loop
                const int64_t t = it * TILE_BARS + tl;
                const bool valid = (row < TILE_BARS * 5) && (t < T);
                const int64_t tc_ = valid ? t : 0;
                int2 kth = make_int2(0, 0); double2 ab = make_double2(0., 0.); double mid = 0.0;
                if (half == 0 && T > 0) {
                    kth = __ldg(reinterpret_cast<const int2*>(&a.sig[tc_].ka1));
                    ab = __ldg(reinterpret_cast<const double2*>(&a.px[tc_].ask));
                    mid = __ldg(&a.px[tc_].mid_next);
                }

                mbar_wait(&sm.d_full[buf], use & 1u);
                tc_fence_after();
                float2 accA[4], accB[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { accA[q] = make_float2(0.f, 0.f); accB[q] = make_float2(0.f, 0.f); }
                const uint32_t tbase = lane_addr + buf * 256u + (uint32_t)col0;
                uint32_t v[2][32];
                tmem_ld32(tbase, v[0]);
#pragma unroll
                for (int cc = 0; cc < H / 64; ++cc) {
                    tmem_ld_wait();                                        // chunk cc has landed
                    if (cc + 1 < H / 64) tmem_ld32(tbase + (uint32_t)((cc + 1) * 32), v[(cc + 1) & 1]);   // prefetch
                    const uint32_t* w = v[cc & 1];
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        const float4 bb = *reinterpret_cast<const float4*>(&sm.b2[col0 + cc * 32 + c]);
                        const float4 wa = *reinterpret_cast<const float4*>(&sm.w3a[col0 + cc * 32 + c]);
                        const float4 wb = *reinterpret_cast<const float4*>(&sm.w3b[col0 + cc * 32 + c]);
                        float2 x0 = __fadd2_rn(make_float2(__uint_as_float(w[c]), __uint_as_float(w[c + 1])), make_float2(bb.x, bb.y));
                        float2 x1 = __fadd2_rn(make_float2(__uint_as_float(w[c + 2]), __uint_as_float(w[c + 3])), make_float2(bb.z, bb.w));
                        x0.x = fmaxf(x0.x, 0.f); x0.y = fmaxf(x0.y, 0.f); x1.x = fmaxf(x1.x, 0.f); x1.y = fmaxf(x1.y, 0.f);
                        const int q = (c >> 2) & 1;
                        accA[2 * q] = __ffma2_rn(x0, make_float2(wa.x, wa.y), accA[2 * q]);
                        accB[2 * q] = __ffma2_rn(x0, make_float2(wb.x, wb.y), accB[2 * q]);
                        accA[2 * q + 1] = __ffma2_rn(x1, make_float2(wa.z, wa.w), accA[2 * q + 1]);
                        accB[2 * q + 1] = __ffma2_rn(x1, make_float2(wb.z, wb.w), accB[2 * q + 1]);
                    }
                }
                accA[0] = __fadd2_rn(accA[0], accA[2]); accA[1] = __fadd2_rn(accA[1], accA[3]);
                accB[0] = __fadd2_rn(accB[0], accB[2]); accB[1] = __fadd2_rn(accB[1], accB[3]);
                // the accumulator has been read: hand the TMEM buffer back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.d_empty[buf]);

                // combine the two column halves: the upper half hands its partial sums over
                float pa = __fadd_rn(__fadd_rn(accA[0].x, accA[1].x), __fadd_rn(accA[0].y, accA[1].y));
                float pb = __fadd_rn(__fadd_rn(accB[0].x, accB[1].x), __fadd_rn(accB[0].y, accB[1].y));
                if (half == 1) sm.part[g & 1u][row] = make_float2(pa, pb);
                asm volatile("bar.sync 1, 256;" ::: "memory");            // the 8 epilogue warps only
                if (half == 1) continue;
                const float2 other = sm.part[g & 1u][row];
                const float ra = __fadd_rn(__fadd_rn(pa, other.x), sm.b3[0]);
                const float rb = __fadd_rn(__fadd_rn(pb, other.y), sm.b3[1]);
                const int ka = __float2int_rn(__fmul_rn(ra, 5.0f));          // drl_engine.py:39
                const int kb = __float2int_rn(__fmul_rn(rb, 5.0f));
                // speculative env step of (bar t, inventory iv-2)  (market_env.py:30-58)
                TableEntry e;
                e.ka = ka; e.kb = kb; e.raw_a = ra; e.raw_b = rb; e.reward = 0.0; e.next = iv; e.traded = 0;
                if (valid) {
                    const int inv = iv - 2;
                    const bool fb = (inv < 2) && (kb < kth.y);               // :34,:37
                    const bool fs = (inv > -2) && (ka < kth.x);              // :35,:38
                    const double my_ask = add_rn(ab.x, mul_rn((double)ka, a.tick));
                    const double my_bid = sub_rn(ab.y, mul_rn((double)kb, a.tick));
                    double leg_b = sub_rn(mid, my_bid), leg_s = sub_rn(my_ask, mid);
                    if (FEE) {
                        leg_b = sub_rn(leg_b, mul_rn(my_bid, a.fee));
                        leg_s = sub_rn(leg_s, mul_rn(my_ask, a.fee));
                    }
                    double pnl = 0.0;
                    pnl = fb ? add_rn(pnl, leg_b) : pnl;
                    pnl = fs ? add_rn(pnl, leg_s) : pnl;
                    const int ninv = inv + (fb ? 1 : 0) - (fs ? 1 : 0);
                    const int ai = ninv < 0 ? -ninv : ninv;
                    e.reward = sub_rn(pnl, mul_rn(a.phi, (double)ai));        // :57-58
                    e.next = ninv + 2;
                    e.traded = (fb || fs) ? 1 : 0;
                    if (a.raw_table) {
                        float* o = a.raw_table + (((int64_t)ind * T + t) * 5 + iv) * 2;
                        o[0] = ra; o[1] = rb;
                    }
                }
                // publish the tile's table
                mbar_wait(&sm.t_empty[buf], (use & 1u) ^ 1u);
                sm.table[buf][row] = e;
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.t_full[buf]);
            }
        } else if (warp < NUM_EPI_WARPS + NUM_PROD_WARPS) {
            // =========================== PRODUCER ==============================================
            const int ptid = tid - NUM_EPI_WARPS * 32;     // 0..127
            const int c = ptid & 7;                        // this thread's 16-byte chunk (8 k) in every k-block
            for (int64_t it = 0; it < ntiles; ++it) {
                const uint32_t g = gt + (uint32_t)it;
                const int64_t t0 = it * TILE_BARS;
                // the thread's (up to) two bars of this tile
                float2 z[2]; bool ok[2]; int tl[2];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const int q = ptid + 128 * m;          // task id: (bar, chunk)
                    tl[m] = q >> 3;
                    ok[m] = (tl[m] < TILE_BARS);
                    const int64_t t = t0 + tl[m] < T ? t0 + tl[m] : T - 1;
                    z[m] = ok[m] ? *reinterpret_cast<const float2*>(&a.sig[t].z1) : make_float2(0.f, 0.f);
                }
#pragma unroll 1
                for (int kb = 0; kb < NKB; ++kb) {
                    const int k0 = kb * KBLK + c * 8;
                    float2 wx[4], wy[4], wi[4], bb[4];
                    {
                        const float4 p0 = *reinterpret_cast<const float4*>(&sm.w1x[k0]), p1 = *reinterpret_cast<const float4*>(&sm.w1x[k0 + 4]);
                        wx[0] = make_float2(p0.x, p0.y); wx[1] = make_float2(p0.z, p0.w); wx[2] = make_float2(p1.x, p1.y); wx[3] = make_float2(p1.z, p1.w);
                        const float4 q0 = *reinterpret_cast<const float4*>(&sm.w1y[k0]), q1 = *reinterpret_cast<const float4*>(&sm.w1y[k0 + 4]);
                        wy[0] = make_float2(q0.x, q0.y); wy[1] = make_float2(q0.z, q0.w); wy[2] = make_float2(q1.x, q1.y); wy[3] = make_float2(q1.z, q1.w);
                        const float4 r0 = *reinterpret_cast<const float4*>(&sm.w1i[k0]), r1 = *reinterpret_cast<const float4*>(&sm.w1i[k0 + 4]);
                        wi[0] = make_float2(r0.x, r0.y); wi[1] = make_float2(r0.z, r0.w); wi[2] = make_float2(r1.x, r1.y); wi[3] = make_float2(r1.z, r1.w);
                        const float4 s0 = *reinterpret_cast<const float4*>(&sm.b1[k0]), s1 = *reinterpret_cast<const float4*>(&sm.b1[k0 + 4]);
                        bb[0] = make_float2(s0.x, s0.y); bb[1] = make_float2(s0.z, s0.w); bb[2] = make_float2(s1.x, s1.y); bb[3] = make_float2(s1.z, s1.w);
                    }
                    // slot kb of the ring: wait until the MMA of the previous tile has consumed it
                    if (g > 0) mbar_wait(&sm.a_empty[kb], (g - 1) & 1u);
                    uint8_t* slab = &sm.a_tile[kb][0];
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        if (ok[m]) {
                            float2 A[4];
#pragma unroll
                            for (int p = 0; p < 4; ++p)          // SGMM-F32 layer 1: b1, +W1[.,0] z1, +W1[.,1] z2
                                A[p] = __ffma2_rn(wy[p], make_float2(z[m].y, z[m].y), __ffma2_rn(wx[p], make_float2(z[m].x, z[m].x), bb[p]));
#pragma unroll
                            for (int iv = 0; iv < 5; ++iv) {
                                const float inv2 = (float)(iv - 2) * 0.5f;       // drl_engine.py:35
                                uint4 o;
                                float2 v0 = __ffma2_rn(wi[0], make_float2(inv2, inv2), A[0]);
                                float2 v1 = __ffma2_rn(wi[1], make_float2(inv2, inv2), A[1]);
                                float2 v2 = __ffma2_rn(wi[2], make_float2(inv2, inv2), A[2]);
                                float2 v3 = __ffma2_rn(wi[3], make_float2(inv2, inv2), A[3]);
                                o.x = pack_relu_bf16(v0.x, v0.y); o.y = pack_relu_bf16(v1.x, v1.y);
                                o.z = pack_relu_bf16(v2.x, v2.y); o.w = pack_relu_bf16(v3.x, v3.y);
                                const int r = tl[m] * 5 + iv;
                                *reinterpret_cast<uint4*>(slab + swz(r, c)) = o;
                            }
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.a_full[kb]);
                }
            }
        } else if (warp == WARP_MMA) {
            // =========================== MMA ISSUER ============================================
            if (lane == 0) {
                for (int64_t it = 0; it < ntiles; ++it) {
                    const uint32_t g = gt + (uint32_t)it, buf = g & 1u, use = g >> 1;
                    mbar_wait(&sm.d_empty[buf], (use & 1u) ^ 1u);          // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d = tmem_base + buf * 256u;
                    for (int kb = 0; kb < NKB; ++kb) {
                        mbar_wait(&sm.a_full[kb], g & 1u);
                        tc_fence_after();
                        const uint64_t ad = make_desc(smem_u32(&sm.a_tile[kb][0]));
                        const uint64_t bd = make_desc(smem_u32(&sm.b_tile[kb][0]));
#pragma unroll
                        for (int k = 0; k < KBLK / UMMA_K; ++k)              // +32 B per K=16 step inside the swizzle atom
                            umma(d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), (kb | k) != 0 ? 1u : 0u);
                        umma_commit(&sm.a_empty[kb]);                        // slot free once these MMAs retire
                    }
                    umma_commit(&sm.d_full[buf]);                            // accumulator complete
                }
            }
            __syncwarp();
        } else {
            // =========================== WALKER ================================================
            if (lane == 0) {
                int iv = 2, trades = 0;                                       // inventory 0
                double total = 0.0;                                           // drl_engine.py:26
                for (int64_t it = 0; it < ntiles; ++it) {
                    const uint32_t g = gt + (uint32_t)it, buf = g & 1u, use = g >> 1;
                    mbar_wait(&sm.t_full[buf], use & 1u);
                    const int64_t t0 = it * TILE_BARS;
                    const int n = (int)(T - t0 < TILE_BARS ? T - t0 : TILE_BARS);
                    for (int s = 0; s < n; ++s) {
                        const TableEntry& e = sm.table[buf][s * 5 + iv];
                        total = add_rn(total, e.reward);                      // drl_engine.py:54
                        trades += e.traded;
                        if (a.act_trace) { a.act_trace[((int64_t)ind * T + t0 + s) * 2] = e.ka; a.act_trace[((int64_t)ind * T + t0 + s) * 2 + 1] = e.kb; }
                        iv = e.next;
                    }
                    mbar_arrive(&sm.t_empty[buf]);
                }
                if (trades == 0) total = sub_rn(total, 50.0);                 // drl_engine.py:64-65
                a.fitness[ind] = total; a.trades[ind] = trades;
            }
            __syncwarp();
        }
        gt += (uint32_t)ntiles;
        tc_fence_before();
        __syncthreads();                            // every role is done with this individual's weights
        tc_fence_after();
    }

    if (warp == WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace s256

int launch_spec256(const sgmm_bundle* b, const PopArgs& mm, double phi, double fee, double* fitness, int32_t* trades,
                   float* raw_table, int32_t* act_trace, cudaStream_t st)
{
    using namespace s256;
    if (mm.count == 0) return SGMM_OK;
    Args a;
    a.sig = b->sig; a.px = b->px; a.T = b->T; a.tick = b->tick; a.phi = phi; a.fee = fee;
    a.mm = mm; a.fitness = fitness; a.trades = trades; a.raw_table = raw_table; a.act_trace = act_trace;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device);
    const int grid = (int)(mm.count < sms ? mm.count : sms);
    const size_t smem = sizeof(Smem) + 1024;
    static bool configured[2] = {false, false};
    const bool has_fee = fee != 0.0;
    auto kern = has_fee ? spec256_kernel<true> : spec256_kernel<false>;
    if (!configured[has_fee]) {
        if (int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                "cudaFuncSetAttribute(spec256 smem)")) return rc;
        configured[has_fee] = true;
    }
    kern<<<grid, NUM_THREADS, smem, st>>>(a);
    return check_cuda(cudaGetLastError(), "spec256_kernel launch");
}

}  // namespace sgmm
