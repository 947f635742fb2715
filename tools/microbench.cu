// microbench.cu -- per-instruction latency (1 warp, dependent chain) and throughput (full chip)
// of the operations the rollout kernel's step is made of.  nvcc -arch=sm_100a microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N 4096
template <int OP>
__global__ void lat_kernel(float* out, long long* cyc, float a, double da, int ia)
{
    float x = a + threadIdx.x; float2 x2 = make_float2(x, x + 1); double d = da + threadIdx.x; int k = ia + threadIdx.x;
    __shared__ float4 sm[64];
    sm[threadIdx.x] = make_float4(0, 0, 0, 0); sm[threadIdx.x + 32] = make_float4(0, 0, 0, 0);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = __fmaf_rn(x, 1.0000001f, 1e-7f);
        if (OP == 1) x2 = __ffma2_rn(x2, make_float2(1.0000001f, 0.9999999f), make_float2(1e-7f, 1e-7f));
        if (OP == 2) d = __dadd_rn(d, 1e-7);
        if (OP == 3) d = __dmul_rn(d, 1.0000001);
        if (OP == 4) { d = __dadd_rn((double)k, d); k = (int)d & 1023; }           // I2F.F64 + DADD + F2I
        if (OP == 5) { k = __float2int_rn(x); x = __fmul_rn((float)k, 0.5f) + 1.f; }  // F2I + I2F + FMUL + FADD
        if (OP == 6) x = __fadd_rn(x, __shfl_xor_sync(0xffffffffu, x, 1));
        if (OP == 7) { float4 v = sm[(k & 31)]; k = __float_as_int(v.x) + i; }     // LDS.128 pointer chase
        if (OP == 8) { sm[threadIdx.x].x = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31].x + 1.0f; __syncwarp(); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[OP] = t1 - t0;
    out[threadIdx.x] = x + x2.x + x2.y + (float)d + k;
}

template <int OP>
__global__ void tput_kernel(float* out, float a, double da)
{
    float x[8]; float2 x2[8]; double d[8]; float2 w2r[8]; float2 h2r = make_float2(a * 0.5f, a * 0.25f);
    for (int j = 0; j < 8; ++j) { x[j] = a + j + threadIdx.x; x2[j] = make_float2(x[j], x[j]); d[j] = da + j; w2r[j] = make_float2(1.0f + 1e-7f * (j + threadIdx.x), 1.0f - 1e-7f * j); }
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (OP == 0) x[j] = __fmaf_rn(x[j], 1.0000001f, 1e-7f);
            if (OP == 1) x2[j] = __ffma2_rn(x2[j], make_float2(1.0000001f, 0.9999999f), make_float2(1e-7f, 1e-7f));
            if (OP == 2) d[j] = __dadd_rn(d[j], 1e-7);
            if (OP == 3) d[j] = __dmul_rn(d[j], 1.0000001);
            if (OP == 9) { x2[j] = __ffma2_rn(x2[j], make_float2(x[j], x[(j + 1) & 7]), x2[(j + 3) & 7]); }
            // mixes: do FFMA and FFMA2 (or FFMA2 and DADD) run on separate pipes?  OP 10: one FFMA2 + one FFMA; OP 11: one FFMA2 + two FFMA;
            // OP 12: one FFMA2 + one DADD (counted in lane-FMAs resp. instructions by the caller)
            if (OP == 10) { x2[j] = __ffma2_rn(x2[j], make_float2(1.0000001f, 0.9999999f), make_float2(1e-7f, 1e-7f)); x[j] = __fmaf_rn(x[j], 1.0000001f, 1e-7f); }
            if (OP == 11) { x2[j] = __ffma2_rn(x2[j], make_float2(1.0000001f, 0.9999999f), make_float2(1e-7f, 1e-7f)); x[j] = __fmaf_rn(x[j], 1.0000001f, 1e-7f); d[j] = __hiloint2double(__float_as_int(__fmaf_rn(__int_as_float(__double2hiint(d[j])), 1.0000001f, 1e-7f)), __double2loint(d[j])); }
            // the hidden-layer pattern of sgmm_rollout.cu: 8 independent accumulators, weights and activations in registers
            if (OP == 13) { x2[j] = __ffma2_rn(w2r[j], h2r, x2[j]); }
            if (OP == 14) { x[j] = __fmaf_rn(w2r[j].x, h2r.x, x[j]); }
            if (OP == 12) { x2[j] = __ffma2_rn(x2[j], make_float2(1.0000001f, 0.9999999f), make_float2(1e-7f, 1e-7f)); d[j] = __dadd_rn(d[j], 1e-7); }
        }
    }
    float s = 0; for (int j = 0; j < 8; ++j) s += x[j] + x2[j].x + x2[j].y + (float)d[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> void run_lat(const char* name, float* out, long long* cyc)
{
    lat_kernel<OP><<<1, 32>>>(out, cyc, 1.0f, 1.0, 1);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("latency  %-28s %7.2f cycles/iter\n", name, (double)h[OP] / N);
}
template <int OP> void run_tput(const char* name, float* out, int per_thread_ops)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tput_kernel<OP><<<148 * 4, 512>>>(out, 1.0f, 1.0);
    cudaEventRecord(e0);
    tput_kernel<OP><<<148 * 4, 512>>>(out, 1.0f, 1.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_instrs = (double)148 * 4 * 16 * N * 8;
    double per_sm_per_clk = warp_instrs / 148.0 / (ms * 1e-3 * 1.965e9);
    printf("tput     %-28s %7.3f ms  %6.3f warp-instr/clk/SM (at 1965 MHz)  -> %.1f lane-ops/clk/SM\n", name, ms,
           per_sm_per_clk, per_sm_per_clk * 32 * per_thread_ops);
}
int main()
{
    float* out; long long* cyc; cudaMalloc(&out, 148 * 4 * 512 * 4); cudaMalloc(&cyc, 128);
    run_lat<0>("FFMA", out, cyc); run_lat<1>("FFMA2", out, cyc); run_lat<2>("DADD", out, cyc); run_lat<3>("DMUL", out, cyc);
    run_lat<4>("I2F.F64+DADD+F2I.F64", out, cyc); run_lat<5>("F2I+I2F+FMUL+FADD", out, cyc);
    run_lat<6>("SHFL.BFLY+FADD", out, cyc); run_lat<7>("LDS.128 chase", out, cyc); run_lat<8>("STS+sync+LDS+sync+FADD", out, cyc);
    run_tput<0>("FFMA", out, 1); run_tput<1>("FFMA2", out, 2); run_tput<2>("DADD", out, 1); run_tput<3>("DMUL", out, 1);
    run_tput<9>("FFMA2 (reg operands)", out, 2);
    run_tput<10>("1 FFMA2 + 1 FFMA (3 FMA/iter)", out, 3); run_tput<11>("1 FFMA2 + 2 FFMA (4 FMA/iter)", out, 4); run_tput<12>("1 FFMA2 + 1 DADD (2 instr)", out, 2);
    run_tput<13>("FFMA2 acc += w(reg) * h(reg)", out, 2); run_tput<14>("FFMA  acc += w(reg) * h(reg)", out, 1);
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
