// sgmm_internal.h -- shared declarations between the translation units of libsgmm_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <mutex>
#include <vector>
#include "../../include/sgmm.h"

namespace sgmm {

// One bar as the kernels see it.  16 B signal/threshold record + 32 B price record.
struct __align__(32) BarSig {
    float z1, z2;        // normalised SGU1 / SGU2 signals (drl_engine.py:33-34), fp32
    float tha, thb;      // float fill thresholds on q = raw*5 (BEFORE rounding):
                         //   fill_sell <=> q_a < tha, fill_buy <=> q_b < thb   (-inf never, +inf always)
                         // tha = Ka + 0.5 nudged one ulp up when Ka is even, so that the strict
                         // compare reproduces round-half-even followed by off <= Ka
    int32_t ka1, kb1;    // integer thresholds + 1: fill_sell <=> off_a < ka1 (used when the adversary
    int32_t pad0, pad1;  // displaces the rounded offsets); INT32_MIN never, INT32_MAX always
};
struct __align__(32) BarPx {
    double ask, bid, mid_next, pad;
};

// step code of the deferred fp64 accounting (sgmm_account.cu): this value in an offset field = that side did not fill
#define SGMM_CODE_NOFILL INT32_MIN
#define SGMM_CODE_NOFILL_F ((int)0xFFFFFFFF)     // the same marker when the field carries the fp32 q = raw*5 (a NaN)
constexpr int32_t K_NEVER = INT32_MIN;
constexpr int32_t K_ALWAYS = INT32_MAX;
constexpr int32_t K_CLAMP = 1 << 30;
constexpr int32_t K_FLOAT_EXACT = 1 << 22;   // |threshold| below which Ka+0.5 is exact in fp32

}  // namespace sgmm

#define SGMM_HOST_SLOTS 3          // 1 synchronous + 2 pipelined host-buffer workspaces per bundle

struct sgmm_bundle {
    int device = 0;
    int64_t T = 0;
    double tick = 0.0;
    sgmm::BarSig* sig = nullptr;       // [T]
    sgmm::BarPx* px = nullptr;         // [T]
    double* bmax = nullptr;            // [T] raw bounds (trace kernel runs the literal step core)
    double* smin = nullptr;
    uint8_t* a1 = nullptr;             // [ceil(T/25)][4096] layer-1 A operand tiles of the tensor-core H=32 path (sgmm_tc32.cu)
    // grow-only workspaces of the *_host entry points: slot 0 = the synchronous entry (caller's stream), slots 1.. = the
    // pipelined entry (sgmm_rollout_population_host_async), each with its own stream so that the H2D copy of one batch
    // overlaps the kernel of the previous one
    std::mutex ws_mutex;
    void* ws[SGMM_HOST_SLOTS] = {};
    size_t ws_bytes[SGMM_HOST_SLOTS] = {};
    cudaStream_t slot_stream[SGMM_HOST_SLOTS] = {};
    uint64_t async_next = 0;
    // [count][T] 64-bit step codes of the tensor-core rollouts (sgmm_account.cu), grow-only
    std::mutex codes_mutex;
    uint64_t* codes = nullptr;
    size_t codes_cap = 0;
    std::vector<uint64_t*> codes_retired;
    // leg tables of the tensor-core H=32 rollout (sgmm_tc32.cu): one per fee rate seen, never freed before the bundle
    struct LegTable { double fee; void* buf; cudaEvent_t ready; cudaStream_t stream; };
    std::mutex legs_mutex;
    std::vector<LegTable> legs;
    // ordering of the users of `codes` across streams: the last user's stream and an event recorded after its accounting
    // kernel; a rollout arriving on another stream waits for it (outside stream capture)
    cudaStream_t codes_stream = nullptr;
    cudaEvent_t codes_done = nullptr;
    bool codes_busy = false;
    bool codes_captured = false;       // a stream capture has used `codes`: its address may live in a graph
};

namespace sgmm {

// device-side population descriptor
struct PopArgs {
    const float* genomes;   // [count,G] or nullptr
    const float* master;    // [G]
    const int64_t* first_index_dev;   // optional device scalar added to first_index (GA: best child)
    float sigma;
    const float* sigma_dev;           // optional device scalar overriding sigma (GA sigma decay)
    uint64_t seed;
    uint64_t generation;
    const int32_t* generation_dev;    // optional device scalar overriding generation
    int64_t first_index;
    int64_t count;
};

struct RolloutArgs {
    const BarSig* sig;
    const BarPx* px;
    int64_t T;
    double tick, phi, fee;
    PopArgs mm, adv;
    double* fitness;
    int32_t* trades;
    uint64_t* codes;     // [count][T] step codes (sgmm_account.cu)
};

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: opt in once per (kernel, device).
// `mask` is one static word per kernel instantiation, bit d = configured on device d (the calling thread's current
// device).  Devices >= 64 simply re-apply the attribute on every launch.  Safe from concurrent host threads: a lost
// race only repeats an idempotent call.
template <typename Kernel>
inline int opt_in_smem(Kernel kern, size_t bytes, std::atomic<uint64_t>& mask, const char* what)
{
    int dev = -1;
    if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
    const bool tracked = dev >= 0 && dev < 64;
    if (tracked && (mask.load(std::memory_order_acquire) >> dev & 1ull)) return SGMM_OK;
    if (int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), what)) return rc;
    if (tracked) mask.fetch_or(1ull << dev, std::memory_order_release);
    return SGMM_OK;
}

int launch_rollout(const sgmm_bundle* b, const PopArgs& mm, const PopArgs* adv, int hidden,
                   double phi, double fee, int units_per_lane, int warps_per_cta,
                   double* fitness, int32_t* trades, cudaStream_t st);
// small populations, fast (sgmm_one.cu): policy table for every (bar, inventory) + automaton scan + reference-order sum
constexpr int64_t SMALL_POP_MAX = 400;      // measured break-even against the sequential kernel: ~420 individuals (profiles/r2_small_population_path.log)
constexpr int64_t SMALL_POP_MAX_ADV = 148;  // with the adversary (one 512-thread CTA per individual and SM)
int launch_rollout_small(const sgmm_bundle* b, const PopArgs& mm, const PopArgs* adv, double phi, double fee, double* fitness, int32_t* trades,
                         cudaStream_t st);
int launch_trace(const sgmm_bundle* b, const float* mm_genome, int hidden, const float* adv_genome,
                 const int32_t* forced, const int32_t* table, double phi, double fee, const sgmm_trace* tr,
                 double* fitness, int32_t* trades, cudaStream_t st);
int launch_spec256(const sgmm_bundle* b, const PopArgs& mm, double phi, double fee, double* fitness, int32_t* trades,
                   float* raw_table, int32_t* act_trace, cudaStream_t st);
int launch_tc32(const sgmm_bundle* b, const PopArgs& mm, const PopArgs* adv, double phi, double fee, int group, double* fitness, int32_t* trades,
                float* raw_table, int32_t* act_trace, cudaStream_t st, int mode = 0);   // mode: 0 bf16, 1 tf32, 2 f16
inline int tc32_mode_of(int precision) { return precision == SGMM_PRECISION_TF32 ? 1 : (precision == SGMM_PRECISION_F16 ? 2 : 0); }
int launch_tc32_prologue(sgmm_bundle* b, cudaStream_t st);
int launch_account(const sgmm_bundle* b, const uint64_t* codes, int64_t count, double phi, double fee, double* fitness,
                   int32_t* trades, cudaStream_t st, bool float_offsets = false);
int reserve_codes(const sgmm_bundle* b, int64_t count, cudaStream_t st, uint64_t** out);
int release_codes(const sgmm_bundle* b, cudaStream_t st);      // after the last kernel that reads the codes was enqueued
size_t tc32_a1_bytes(int64_t T);
int launch_prologue(sgmm_bundle* b, const float* z1, const float* z2, const double* mid,
                    const double* ask, const double* bid, cudaStream_t st);

int launch_bundle_windows(int64_t E, const double* ask1, const double* bid1, const double* pmax, const double* pmin,
                          int64_t step, int64_t n, double* mid_next, double* best_ask, double* best_bid,
                          double* buy_max, double* sell_min, cudaStream_t st);
int launch_analytics(int64_t B, int64_t T, const double* wealth, const double* cash, const double* mid,
                     const int32_t* inventory, const uint8_t* is_trade, double* scratch, double* out, cudaStream_t st);

inline int64_t genome_len(int H) { return (int64_t)H * H + 7 * (int64_t)H + 2; }

}  // namespace sgmm
