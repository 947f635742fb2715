// tc32_unit2.cu -- probe: tcgen05.mma kind::f16 with an F16 accumulator (c_format = 0): where do the 16-bit results live in TMEM?
#include <cuda_fp16.h>
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
    for (;;) { uint32_t ok; asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory"); if (ok) return; }
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__host__ __device__ inline uint32_t canon(int r, int k, int K) { return (uint32_t)(((r >> 3) * (K >> 3) + (k >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2); }
// a = b = f16 (format 0), c = f16 (format 0) or f32 (1)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int cfmt) { return ((uint32_t)cfmt << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__global__ void __launch_bounds__(128, 1) k(const __half* A, const __half* B, uint32_t* out, uint32_t* out2)
{
    __shared__ __align__(1024) uint8_t sA[128 * 16 * 2];
    __shared__ __align__(1024) uint8_t sB[32 * 16 * 2];
    __shared__ uint64_t bar; __shared__ uint32_t tb_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 16; i += 128) *reinterpret_cast<__half*>(sA + canon(i / 16, i % 16, 16)) = A[i];
    for (int i = tid; i < 32 * 16; i += 128) *reinterpret_cast<__half*>(sB + canon(i / 16, i % 16, 16)) = B[i];
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb_s)), "r"(128u) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tb_s, la = tb + ((uint32_t)(warp * 32) << 16);
    // poison 64 columns
    { uint32_t z = 0xDEADBEEFu; for (int c = 0; c < 64; ++c) asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(la + c), "r"(z) : "memory"); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        uint64_t ad = make_desc(smem_u32(sA), 128, 256), bd = make_desc(smem_u32(sB), 128, 256);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tb), "l"(ad), "l"(bd), "r"(idesc_f16(128, 32, 0)), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 64; ++c) { uint32_t v; asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(la + c) : "memory"); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); out[tid * 64 + c] = v; }
    // packed-read variant: .pack::16b reads two 16-bit columns into one register
    for (int c = 0; c < 16; ++c) { uint32_t v; asm volatile("tcgen05.ld.sync.aligned.32x32b.pack::16b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(la + 2 * c) : "memory"); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); out2[tid * 16 + c] = v; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(128u) : "memory");
}
int main()
{
    srand(2);
    std::vector<float> A(128 * 16), B(32 * 16);
    for (auto& x : A) x = (float)(rand() % 17 - 8) / 8.f;
    for (auto& x : B) x = (float)(rand() % 17 - 8) / 8.f;
    std::vector<__half> hA(A.size()), hB(B.size());
    for (size_t i = 0; i < A.size(); ++i) hA[i] = __float2half(A[i]);
    for (size_t i = 0; i < B.size(); ++i) hB[i] = __float2half(B[i]);
    __half *dA, *dB; uint32_t *dO, *dO2;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * 64 * 4); cudaMalloc(&dO2, 128 * 16 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    k<<<1, 128>>>(dA, dB, dO, dO2);
    printf("kernel: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    std::vector<uint32_t> O(128 * 64), O2(128 * 16);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(O2.data(), dO2, O2.size() * 4, cudaMemcpyDeviceToHost);
    for (int r : {0, 37}) {
        printf("row %d expected:", r);
        for (int n = 0; n < 8; ++n) { float s = 0; for (int kk = 0; kk < 16; ++kk) s += A[r * 16 + kk] * B[n * 16 + kk]; printf(" %.3f", s); }
        printf("\n  cols (lo,hi as f16):");
        for (int c = 0; c < 36; ++c) { uint32_t v = O[r * 64 + c]; __half lo = __ushort_as_half((unsigned short)(v & 0xffff)), hi = __ushort_as_half((unsigned short)(v >> 16)); if (v == 0xDEADBEEFu) printf(" [P]"); else printf(" (%.3f,%.3f)", __half2float(lo), __half2float(hi)); }
        printf("\n  pack::16b reads:");
        for (int c = 0; c < 8; ++c) { uint32_t v = O2[r * 16 + c]; printf(" (%.3f,%.3f)", __half2float(__ushort_as_half((unsigned short)(v & 0xffff))), __half2float(__ushort_as_half((unsigned short)(v >> 16)))); }
        printf("\n");
    }
    return 0;
}
