// sgmm_account.cu -- the fp64 half of the tensor-core rollouts, run AFTER the rollout kernel.
//
// Measured on B200 (tools/walk_bench.cu, profiles/r1_fp64_under_mma.txt): while tcgen05.mma instructions execute,
// an FP64 instruction of the same SM waits ~740 cycles instead of 9 (100 dependent DADDs: 921 cycles on an idle SM,
// 74 146 with back-to-back N=256 MMAs) -- integer, shared-memory and shuffle instructions are unaffected.  The
// reference's accounting is fp64 by definition (Env/market_env.py:30-58, Env/drl_engine.py:54), so the tensor-core
// kernels do only the INTEGER half of the env step (offsets, fills, inventory) and record, per bar actually visited,
// a 64-bit code; this kernel turns the codes into rewards and sums them in the reference's order:
//     code = ka (24 bit, signed) | kb (24 bit, signed) << 24 | fill_buy << 48 | fill_sell << 49 | |inv'| << 50
// One warp per individual: 32 bars' rewards in parallel (un-fused fp64 in the reference's op order), then the
// running sum bar by bar through shuffles (the only serial part: one DADD per bar).
#include "sgmm_internal.h"
#include "sgmm_step_core.h"

namespace sgmm {

namespace {

constexpr int ACCOUNT_WARPS = 4;

template <bool FEE>
__global__ void __launch_bounds__(ACCOUNT_WARPS * 32) account_kernel(const uint64_t* __restrict__ codes, const BarPx* __restrict__ px,
                                                                     int64_t T, int64_t count, double tick, double phi, double fee,
                                                                     double* __restrict__ fitness, int32_t* __restrict__ trades)
{
    const int lane = threadIdx.x & 31;
    const int64_t ind = (int64_t)blockIdx.x * ACCOUNT_WARPS + (threadIdx.x >> 5);
    if (ind >= count) return;
    const uint64_t* c = codes + ind * T;
    double total = 0.0;                                              // drl_engine.py:26
    int ntr = 0;
    for (int64_t t0 = 0; t0 < T; t0 += 32) {
        const int64_t t = t0 + lane;
        const bool valid = t < T;
        double rew = 0.0;
        bool traded = false;
        if (valid) {
            const uint64_t code = __ldcs(c + t);
            const int ka = ((int)(uint32_t)(code << 8)) >> 8;                          // sign-extend 24 bits
            const int kb = ((int)(uint32_t)((code >> 24) << 8)) >> 8;
            const bool fb = (code >> 48) & 1u, fs = (code >> 49) & 1u;
            const int ai = (int)((code >> 50) & 3u);
            const double2 ab = __ldg(reinterpret_cast<const double2*>(&px[t].ask));
            const double mid = __ldg(&px[t].mid_next);
            const double my_ask = add_rn(ab.x, mul_rn((double)ka, tick));               // market_env.py:30
            const double my_bid = sub_rn(ab.y, mul_rn((double)kb, tick));               // :31
            double leg_b = sub_rn(mid, my_bid), leg_s = sub_rn(my_ask, mid);
            if (FEE) {
                leg_b = sub_rn(leg_b, mul_rn(my_bid, fee));                             // :46,:48
                leg_s = sub_rn(leg_s, mul_rn(my_ask, fee));                             // :52,:54
            }
            double pnl = 0.0;                                                           // :40
            pnl = fb ? add_rn(pnl, leg_b) : pnl;
            pnl = fs ? add_rn(pnl, leg_s) : pnl;
            rew = sub_rn(pnl, mul_rn(phi, (double)ai));                                 // :57-58
            traded = fb || fs;
        }
        ntr += __popc(__ballot_sync(0xffffffffu, traded));                              // drl_engine.py:60-61
        const int n = (int)(T - t0 < 32 ? T - t0 : 32);
        const int rlo = __double2loint(rew), rhi = __double2hiint(rew);
        if (n == 32) {
#pragma unroll
            for (int s = 0; s < 32; ++s)
                total = add_rn(total, __hiloint2double(__shfl_sync(0xffffffffu, rhi, s), __shfl_sync(0xffffffffu, rlo, s)));   // drl_engine.py:54
        } else {
            for (int s = 0; s < n; ++s)
                total = add_rn(total, __hiloint2double(__shfl_sync(0xffffffffu, rhi, s), __shfl_sync(0xffffffffu, rlo, s)));
        }
    }
    if (lane == 0) {
        if (ntr == 0) total = sub_rn(total, 50.0);                                      // drl_engine.py:64-65
        fitness[ind] = total;
        trades[ind] = ntr;
    }
}

}  // namespace

int launch_account(const sgmm_bundle* b, const uint64_t* codes, int64_t count, double phi, double fee, double* fitness,
                   int32_t* trades, cudaStream_t st)
{
    if (count == 0) return SGMM_OK;
    const unsigned grid = (unsigned)((count + ACCOUNT_WARPS - 1) / ACCOUNT_WARPS);
    if (fee != 0.0) account_kernel<true><<<grid, ACCOUNT_WARPS * 32, 0, st>>>(codes, b->px, b->T, count, b->tick, phi, fee, fitness, trades);
    else account_kernel<false><<<grid, ACCOUNT_WARPS * 32, 0, st>>>(codes, b->px, b->T, count, b->tick, phi, fee, fitness, trades);
    return check_cuda(cudaGetLastError(), "account_kernel launch");
}

// The code buffer of a bundle: [count][T] 64-bit codes, grow-only.  It is scratch of ONE rollout at a time: rollouts of
// the tensor-core kernels on the same bundle must be ordered on one stream (or serialised by the caller).
int reserve_codes(const sgmm_bundle* cb, int64_t count, cudaStream_t st, uint64_t** out)
{
    sgmm_bundle* b = const_cast<sgmm_bundle*>(cb);
    std::lock_guard<std::mutex> lock(b->codes_mutex);
    const size_t want = (size_t)count * (size_t)b->T;
    if (b->codes_cap < want) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (st) cudaStreamIsCapturing(st, &cap);
        if (cap != cudaStreamCaptureStatusNone) {
            set_error("the code buffer of the bundle holds %zu codes, this rollout needs %zu: run one rollout of this size outside "
                      "the stream capture first", b->codes_cap, want);
            return SGMM_ERR_INVALID;
        }
        if (b->codes) { cudaDeviceSynchronize(); cudaFree(b->codes); b->codes = nullptr; b->codes_cap = 0; }
        if (int rc = check_cuda(cudaMalloc(&b->codes, want * sizeof(uint64_t)), "cudaMalloc(code buffer)")) return rc;
        b->codes_cap = want;
    }
    *out = b->codes;
    return SGMM_OK;
}

}  // namespace sgmm
