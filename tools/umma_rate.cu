// umma_rate.cu -- how long does one tcgen05.mma take?  M = 128, K = 16 (kind::f16), N swept, operands SS (A and B in
// shared memory, SWIZZLE_128B K-major) or TS (A in tensor memory).  One warp per CTA issues NMMA back-to-back MMAs and
// one commit; clock64 from the first issue to the completion of the commit's mbarrier phase.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu ; run: tools/umma_rate [ctas]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    for (;;) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
    }
}
template <bool TS>
__device__ __forceinline__ void umma(uint32_t d, uint64_t ad, uint32_t a_tmem, uint64_t bd, uint32_t idesc, uint32_t acc)
{
    if (TS)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

template <bool TS>
__global__ void __launch_bounds__(512, 1) rate_kernel(int N, int nmma, int cfmt, int bg, long long* out)
{
    extern __shared__ __align__(1024) uint8_t smem[];          // A: 128 rows x 128 B (16 KB), B: 256 rows x 128 B (32 KB)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop;
    __shared__ __align__(16) uint8_t scratch[8 * 4096];
    if (threadIdx.x == 0) stop = 0;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (48 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    if (warp == 0) {
        const uint32_t fmt = cfmt ? 1u : 0u;                                       // operands: bf16 with fp32 D, f16 with f16 D
        const uint32_t idesc = ((uint32_t)cfmt << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint64_t ad = make_desc(smem_u32(smem)), bd = make_desc(smem_u32(smem + 16384));
        uint32_t par = 0;
        for (int rep = 0; rep < 3; ++rep) {
            __syncwarp();
            const long long t0 = clock64();
            uint32_t leader;
            asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(leader));
            if (leader) {
                for (int i = 0; i < nmma; ++i)
                    umma<TS>(tb + (uint32_t)((bg == 5 ? 0 : (i & 1)) * 256), ad + (uint64_t)((i & 3) * 2), tb + 256u + (uint32_t)((i & 3) * 8), bd + (uint64_t)((i & 3) * 2), idesc, (uint32_t)(i > 1));
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
            const long long t1 = clock64();
            mbar_wait(smem_u32(&bar), par);
            par ^= 1u;
            const long long t2 = clock64();
            if (threadIdx.x == 0 && blockIdx.x == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
        }
        __syncwarp();
        if (threadIdx.x == 0) stop = 1;
    } else if (bg == 1 && warp >= 8) {
        // background: conflict-free STS.128 streams (512 B per warp instruction), 8 warps
        uint4* dst = reinterpret_cast<uint4*>(scratch + (warp - 8) * 4096) + (threadIdx.x & 31);
        uint4 v = make_uint4(warp, 1, 2, 3);
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i * 32] = v;
        }
    } else if (bg == 2 && warp >= 4 && warp < 8) {
        // background: tcgen05.ld of 32 columns per instruction from this warp's lane quarter (columns 256..)
        const uint32_t la = tb + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
        uint32_t acc = 0;
        while (!stop) {
            uint32_t r[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(la) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += r[0] ^ r[31];
        }
        if (acc == 0x12345678u) out[7] = acc;
    } else if (bg == 3 && warp >= 8) {
        // background: broadcast LDS.128 (one wavefront per instruction), 8 warps
        const uint4* src = reinterpret_cast<const uint4*>(scratch + (warp - 8) * 4096);
        uint32_t acc = 0;
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(src + i))); acc += v.x; }
        }
        if (acc == 0x12345678u) out[7] = acc;
    } else if (bg == 4 && warp >= 8) {
        // background: FFMA2-heavy register work (power / issue pressure, no memory)
        float2 a = make_float2(1.0f, 2.0f), b = make_float2(0.5f, 0.25f), c = make_float2(warp, 1.0f);
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 32; ++i) c = __ffma2_rn(a, b, c);
        }
        if (c.x == 0.12345f) out[7] = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

int main(int argc, char** argv)
{
    const int ctas = argc > 1 ? atoi(argv[1]) : 1;
    long long* d; cudaMalloc(&d, 64);
    long long h[6];
    cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024 + 1024);
    cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024 + 1024);
    const int nmma = 64;
    const int bgmax = argc > 2 ? atoi(argv[2]) : 0;
    const int bgmin = argc > 3 ? atoi(argv[3]) : 0;
    const char* bgname[] = {"idle", "STS.128 x8 warps", "tcgen05.ld.x32 x4 warps", "LDS.128 broadcast x8 warps", "FFMA2 x8 warps", "idle, ONE accumulator (dependent chain)"};
    for (int bg = bgmin; bg <= bgmax; ++bg)
    for (int ts = 0; ts < 2; ++ts)
        for (int cfmt = 1; cfmt >= (bgmax ? 1 : 0); --cfmt)
            for (int N : {256, 128, 64, 32, 16}) {
                if (bgmax && bgmax < 5 && N != 256 && N != 32) continue;
                if (ts) rate_kernel<true><<<ctas, 512, 48 * 1024 + 1024>>>(N, nmma, cfmt, bg, d);
                else rate_kernel<false><<<ctas, 512, 48 * 1024 + 1024>>>(N, nmma, cfmt, bg, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
                printf("[%s] %s D=%s N=%3d ctas=%d: issue %5.1f clk/MMA, complete %6.1f clk/MMA  (%.0f%% of 8192 FLOP/clk)\n", bgname[bg], ts ? "TS" : "SS", cfmt ? "f32" : "f16", N, ctas,
                       (double)h[4] / nmma, (double)h[5] / nmma, 100.0 * (2.0 * 128 * N * 16) / ((double)h[5] / nmma) / 8192.0);
            }
    return 0;
}
