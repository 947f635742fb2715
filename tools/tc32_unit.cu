// tc32_unit.cu -- building-block checks for the H=32 tensor-core rollout (sgmm_tc32.cu), run once on a B200:
//   (1) tcgen05.mma kind::f16 SS with SWIZZLE_NONE K-major operands (canonical 8x16B core matrices),
//       M=128 N=32 K=16, both LBO/SBO assignments tried
//   (2) tcgen05.st of packed bf16 pairs + tcgen05.mma TS (A operand in TMEM), K=48 in three K=16 steps,
//       B advanced by 2*LBO per step, A advanced by 8 columns per step
//   (3) N=16 TS MMA
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc32_unit tools/tc32_unit.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nLAB_WAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LAB_DONE_%=;\nbra LAB_WAIT_%=;\nLAB_DONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// SWIZZLE_NONE K-major descriptor: start>>4 | LBO>>4 @16 | SBO>>4 @32 | version 1 @46 | layout 0 @61
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bd, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }

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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// canonical SWIZZLE_NONE K-major offset of element (r, k) in a [rows][K] bf16 tile: 8x(16 B) core matrices,
// k-chunks contiguous (128 B apart), 8-row groups (K/8)*128 B apart
__host__ __device__ inline uint32_t canon(int r, int k, int K) { return (uint32_t)(((r >> 3) * (K >> 3) + (k >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2); }

struct Out { float d1a[128 * 32], d1b[128 * 32], d2[128 * 32], d3[128 * 16], d2b[128 * 32], d3b[128 * 16]; };

__global__ void __launch_bounds__(128, 1) unit_kernel(const __nv_bfloat16* A1, const __nv_bfloat16* B1, const __nv_bfloat16* A2,
                                                       const __nv_bfloat16* B2, const __nv_bfloat16* B3, Out* out)
{
    __shared__ __align__(1024) uint8_t sA1[128 * 16 * 2];
    __shared__ __align__(1024) uint8_t sB1[32 * 16 * 2];
    __shared__ __align__(1024) uint8_t sB2[32 * 48 * 2];
    __shared__ __align__(1024) uint8_t sB3[16 * 48 * 2];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * 16; i += 128) *reinterpret_cast<__nv_bfloat16*>(sA1 + canon(i / 16, i % 16, 16)) = A1[i];
    for (int i = tid; i < 32 * 16; i += 128) *reinterpret_cast<__nv_bfloat16*>(sB1 + canon(i / 16, i % 16, 16)) = B1[i];
    for (int i = tid; i < 32 * 48; i += 128) *reinterpret_cast<__nv_bfloat16*>(sB2 + canon(i / 48, i % 48, 48)) = B2[i];
    for (int i = tid; i < 16 * 48; i += 128) *reinterpret_cast<__nv_bfloat16*>(sB3 + canon(i / 48, i % 48, 48)) = B3[i];
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;
    const uint32_t lane_addr = tb + ((uint32_t)(warp * 32) << 16);
    // TMEM map: D1a @0 (32), D1b @32 (32), A2 @64 (24), D2 @96 (32), D2b @128 (32), D3 @160 (16), D3b @176 (16)
    uint32_t ph = 0;
    // ---- (1) SS, both descriptor conventions
    if (tid == 0) {
        umma_ss(tb + 0, make_desc(smem_u32(sA1), 128, 256), make_desc(smem_u32(sB1), 128, 256), idesc_bf16(128, 32), 0);
        umma_ss(tb + 32, make_desc(smem_u32(sA1), 256, 128), make_desc(smem_u32(sB1), 256, 128), idesc_bf16(128, 32), 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, ph); ph ^= 1;
    tc_fence_after();
    {
        uint32_t v[32];
        tmem_ld32(lane_addr + 0, v); tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out->d1a[tid * 32 + j] = __uint_as_float(v[j]);
        tmem_ld32(lane_addr + 32, v); tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out->d1b[tid * 32 + j] = __uint_as_float(v[j]);
    }
    // ---- (2) A2 -> TMEM (packed bf16 pairs, k even in the low half), TS MMA K=48
    {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(A2 + (size_t)tid * 48);
        for (int c = 0; c < 3; ++c) {
            uint32_t p[8];
            for (int j = 0; j < 8; ++j) p[j] = row[c * 8 + j];
            tmem_st8(lane_addr + 64 + c * 8, p);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        for (int k = 0; k < 3; ++k)
            umma_ts(tb + 96, tb + 64 + k * 8, make_desc(smem_u32(sB2) + k * 256, 128, 768), idesc_bf16(128, 32), k > 0);
        for (int k = 0; k < 3; ++k)
            umma_ts(tb + 160, tb + 64 + k * 8, make_desc(smem_u32(sB3) + k * 256, 128, 768), idesc_bf16(128, 16), k > 0);
        for (int k = 0; k < 3; ++k)
            umma_ts(tb + 128, tb + 64 + k * 8, make_desc(smem_u32(sB2) + k * 256, 768, 128), idesc_bf16(128, 32), k > 0);
        for (int k = 0; k < 3; ++k)
            umma_ts(tb + 176, tb + 64 + k * 8, make_desc(smem_u32(sB3) + k * 256, 768, 128), idesc_bf16(128, 16), k > 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, ph); ph ^= 1;
    tc_fence_after();
    {
        uint32_t v[32];
        tmem_ld32(lane_addr + 96, v); tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out->d2[tid * 32 + j] = __uint_as_float(v[j]);
        uint32_t w[16];
        tmem_ld16(lane_addr + 160, w); tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out->d3[tid * 16 + j] = __uint_as_float(w[j]);
        tmem_ld32(lane_addr + 128, v); tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out->d2b[tid * 32 + j] = __uint_as_float(v[j]);
        tmem_ld16(lane_addr + 176, w); tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out->d3b[tid * 16 + j] = __uint_as_float(w[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u) : "memory");
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main()
{
    srand(1);
    auto rnd = [] { return bf((float)rand() / RAND_MAX * 2.f - 1.f); };
    std::vector<float> A1(128 * 16), B1(32 * 16), A2(128 * 48), B2(32 * 48), B3(16 * 48);
    for (auto& x : A1) x = rnd(); for (auto& x : B1) x = rnd(); for (auto& x : A2) x = rnd();
    for (auto& x : B2) x = rnd(); for (auto& x : B3) x = rnd();
    auto up = [](const std::vector<float>& h) {
        std::vector<__nv_bfloat16> t(h.size()); for (size_t i = 0; i < h.size(); ++i) t[i] = __float2bfloat16(h[i]);
        __nv_bfloat16* d; cudaMalloc(&d, t.size() * 2); cudaMemcpy(d, t.data(), t.size() * 2, cudaMemcpyHostToDevice); return d; };
    Out* dout; cudaMalloc(&dout, sizeof(Out)); cudaMemset(dout, 0, sizeof(Out));
    unit_kernel<<<1, 128>>>(up(A1), up(B1), up(A2), up(B2), up(B3), dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    Out* o = new Out; cudaMemcpy(o, dout, sizeof(Out), cudaMemcpyDeviceToHost);
    auto check = [&](const char* name, const float* got, const std::vector<float>& A, const std::vector<float>& B, int N, int K) {
        double worst = 0;
        for (int r = 0; r < 128; ++r) for (int n = 0; n < N; ++n) {
            double s = 0; for (int k = 0; k < K; ++k) s += (double)A[r * K + k] * B[n * K + k];
            worst = fmax(worst, fabs(s - got[r * N + n]));
        }
        printf("%-40s max|err| = %.3g  %s\n", name, worst, worst < 1e-4 ? "OK" : "MISMATCH");
        return worst < 1e-4;
    };
    bool a = check("SS no-swizzle LBO=128(K) SBO=256(M)", o->d1a, A1, B1, 32, 16);
    bool b = check("SS no-swizzle LBO=256 SBO=128 (swapped)", o->d1b, A1, B1, 32, 16);
    bool c = check("TS A-in-TMEM K=48 N=32", o->d2, A2, B2, 32, 48);
    bool d = check("TS A-in-TMEM K=48 N=16", o->d3, A2, B3, 16, 48);
    check("TS K=48 N=32 swapped LBO/SBO", o->d2b, A2, B2, 32, 48);
    check("TS K=48 N=16 swapped LBO/SBO", o->d3b, A2, B3, 16, 48);
    printf("RESULT ss=%s ts32=%d ts16=%d\n", a ? "lbo_k" : (b ? "lbo_m" : "none"), (int)c, (int)d);
    return 0;
}
