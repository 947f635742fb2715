// walk_bench.cu -- cycles per 25-bar tile of the table walk (sgmm_spec256.cu / sgmm_tc32.cu), alone on an SM.
//   variant 0: two-phase walk by ONE lane (automaton on the byte table, then the fp64 reward sum), fully unrolled
//   variant 1: the same by all 32 lanes (tc32: lane = individual; here all lanes walk the same table)
//   variant 2: entry chasing (load the 8-byte record, pick the byte of the current inventory, next record ...), rolled
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/walk_bench tools/walk_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int TILE_BARS = 25;
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
template <bool FULL>
__device__ __forceinline__ void walk_tile(const uint8_t* nb, const double* rb, int n, int& iv, int& trades, double& total)
{
    uint32_t es[TILE_BARS], ivs[TILE_BARS];
    uint32_t w = (uint32_t)iv;
#pragma unroll
    for (int s = 0; s < TILE_BARS; ++s) {
        ivs[s] = w; es[s] = 0;
        if (FULL || s < n) {
            const uint2 x = *reinterpret_cast<const uint2*>(nb + s * 8);
            es[s] = __byte_perm(x.x, x.y, w);
            w = es[s] & 7u;
        }
    }
    iv = (int)w;
#pragma unroll
    for (int s = 0; s < TILE_BARS; ++s) {
        if (FULL || s < n) {
            trades += (int)(es[s] & 8u);
            total = add_rn(total, rb[s * 5 + ivs[s]]);
        }
    }
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// the same walk while warp 1 keeps the tensor pipe busy with back-to-back M=128, N=256, K=16 MMAs (mma = 1) or
// while warps 2.. run FFMA2 / HFMA2 loops (mma = 2 / 3)
__global__ void __launch_bounds__(320, 1) walk_under_load(int mma, int fg, int tiles, long long* out, double* sink)
{
    extern __shared__ __align__(1024) uint8_t opnd[];          // A 16 KB + B 32 KB, zero
    __shared__ double tab_r[2][128];
    __shared__ __align__(8) uint8_t tab_n[2][TILE_BARS * 8];
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(opnd)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) stop = 0;
    if (warp == 0) {
        for (int i = lane; i < 256; i += 32) (&tab_r[0][0])[i] = 1e-3 * (i % 17);
        for (int i = lane; i < 2 * TILE_BARS * 8; i += 32) (&tab_n[0][0])[i] = (uint8_t)(((i * 7 + 3) % 5) | ((i & 3) == 0 ? 8 : 0));
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    if (warp == 0) {
        int iv = 2, trades = 0; double total = 0.0;
        const long long t0 = clock64();
        if (fg == 0) { if (lane == 0) for (int it = 0; it < tiles; ++it) walk_tile<true>(tab_n[it & 1], tab_r[it & 1], 25, iv, trades, total); }
        else if (fg == 1) {                                      // 100 dependent LDS (pointer chase through the byte table)
            uint32_t p = lane;
            for (int it = 0; it < tiles; ++it)
#pragma unroll 4
                for (int k = 0; k < 100; ++k) p = (&tab_n[0][0])[(p * 8 + k) % 400] & 7u;
            iv = (int)p;
        } else if (fg == 2) {                                    // 100 dependent SHFL
            uint32_t p = lane;
            for (int it = 0; it < tiles; ++it)
#pragma unroll 4
                for (int k = 0; k < 100; ++k) p = __shfl_sync(0xffffffffu, p, (p + k) & 31);
            iv = (int)p;
        } else if (fg == 3) {                                    // 100 dependent DADD
            for (int it = 0; it < tiles; ++it)
#pragma unroll 4
                for (int k = 0; k < 100; ++k) total = add_rn(total, 1e-3);
        } else if (fg == 4) {                                    // 100 dependent LOP/IADD
            uint32_t p = lane;
            for (int it = 0; it < tiles; ++it)
#pragma unroll 4
                for (int k = 0; k < 100; ++k) p = (p ^ (p >> 3)) + k;
            iv = (int)p;
        } else if (fg == 6 || fg == 7) {                         // 100 x (tcgen05.ld.x32 + wait [+ tcgen05.st.x16 + wait]) on columns 384..
            uint32_t r[32];
            const uint32_t la = tb + 384u;                       // warp 0: lanes 0..31
            for (int it = 0; it < tiles; ++it)
                for (int k = 0; k < 100; ++k) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(la) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (fg == 7) {
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                            ::"r"(la), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    }
                    iv += (int)(r[0] & 1u);
                }
        } else if (fg == 5) {                                    // 100 independent conflict-free STS.128
            uint4 v = make_uint4(lane, 1, 2, 3);
            for (int it = 0; it < tiles; ++it)
#pragma unroll 4
                for (int k = 0; k < 100; ++k) reinterpret_cast<uint4*>(opnd + 48 * 1024 - 512)[lane] = v;
        }
        __syncwarp();
        const long long t1 = clock64();
        if (lane == 0) { out[0] = (t1 - t0) / tiles; sink[0] = total + trades + iv; stop = 1; }
    } else if (warp == 1 && (mma == 1 || mma >= 4)) {
        // mma 1: SS N=256; 4: TS N=256 (A from tensor memory); 5: SS N=128; 6: SS N=64
        const int N = mma == 5 ? 128 : (mma == 6 ? 64 : 256);
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint64_t ad = make_desc(smem_u32(opnd)), bd = make_desc(smem_u32(opnd + 16384));
        uint32_t leader;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(leader));
        if (leader) {
            while (!stop) {
                for (int i = 0; i < 16; ++i) {
                    if (mma == 4)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                                     ::"r"(tb), "r"(tb + 256u + (uint32_t)((i & 3) * 8)), "l"(bd + (uint64_t)((i & 3) * 2)), "r"(idesc), "r"(1u) : "memory");
                    else
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                     ::"r"(tb), "l"(ad + (uint64_t)((i & 3) * 2)), "l"(bd + (uint64_t)((i & 3) * 2)), "r"(idesc), "r"(1u) : "memory");
                }
            }
        }
        __syncwarp();
    } else if (warp >= 2 && mma == 2) {
        float2 a = make_float2(1.0f, 2.0f), b = make_float2(0.5f, 0.25f), c = make_float2(warp, 1.0f);
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 32; ++i) c = __ffma2_rn(a, b, c);
        }
        if (c.x == 0.12345f) out[15] = 1;
    } else if (warp >= 2 && mma == 3) {
        uint32_t c = warp, a = 0x3C003C00u, b = 0x38003800u;
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 32; ++i) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(c) : "r"(a), "r"(b));
        }
        if (c == 0x12345u) out[15] = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}
__global__ void __launch_bounds__(32, 1) walk_kernel(int variant, int tiles, long long* out, double* sink)
{
    __shared__ double tab_r[2][128];
    __shared__ __align__(8) uint8_t tab_n[2][TILE_BARS * 8];
    const int lane = threadIdx.x;
    for (int i = lane; i < 256; i += 32) (&tab_r[0][0])[i] = 1e-3 * (i % 17);
    for (int i = lane; i < 2 * TILE_BARS * 8; i += 32) (&tab_n[0][0])[i] = (uint8_t)(((i * 7 + 3) % 5) | ((i & 3) == 0 ? 8 : 0));
    __syncwarp();
    int iv = 2, trades = 0; double total = 0.0;
    const long long t0 = clock64();
    if (variant == 0) {
        if (lane == 0) for (int it = 0; it < tiles; ++it) walk_tile<true>(tab_n[it & 1], tab_r[it & 1], 25, iv, trades, total);
    } else if (variant == 1) {
        for (int it = 0; it < tiles; ++it) walk_tile<true>(tab_n[it & 1], tab_r[it & 1], 25, iv, trades, total);
    } else {
        if (lane == 0) for (int it = 0; it < tiles; ++it) {
            const uint8_t* nb = tab_n[it & 1]; const double* rb = tab_r[it & 1];
#pragma unroll 1
            for (int s = 0; s < 25; ++s) {
                const uint32_t e = nb[s * 8 + iv];
                total = add_rn(total, rb[s * 5 + iv]);
                trades += e & 8; iv = e & 7;
            }
        }
    }
    __syncwarp();
    const long long t1 = clock64();
    if (lane == 0) { out[variant] = (t1 - t0) / tiles; sink[variant] = total + trades + iv; }
}
int main(int argc, char** argv)
{
    long long* d; double* s; cudaMalloc(&d, 128); cudaMalloc(&s, 128);
    cudaFuncSetAttribute(walk_under_load, cudaFuncAttributeMaxDynamicSharedMemorySize, 49 * 1024);
    for (int v = 0; v < 3; ++v) walk_kernel<<<1, 32>>>(v, 200, d, s);
    cudaDeviceSynchronize();
    long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("cycles per 25-bar tile: two-phase one lane %lld, two-phase 32 lanes %lld, entry chasing (rolled) %lld\n", h[0], h[1], h[2]);
    const char* nm[] = {"idle SM", "SS N=256 MMAs back to back", "8 warps of FFMA2", "8 warps of HFMA2", "TS N=256 MMAs (A in TMEM)", "SS N=128 MMAs", "SS N=64 MMAs"};
    const char* fn[] = {"two-phase walk of a tile (one lane)", "100 dependent LDS", "100 dependent SHFL", "100 dependent DADD", "100 dependent LOP+IADD", "100 STS.128", "100 x (tcgen05.ld.x32 + wait)", "100 x (ld.x32 + wait + st.x16 + wait)"};
    for (int f = (argc > 1 ? 6 : 0); f < 8; ++f)
        for (int m = 0; m < 7; ++m) {
            walk_under_load<<<1, 320, 49 * 1024>>>(m, f, 50, d, s);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long v; cudaMemcpy(&v, d, 8, cudaMemcpyDeviceToHost);
            printf("%-38s | %-28s: %7lld cycles\n", fn[f], nm[m], v);
        }
    return 0;
}
