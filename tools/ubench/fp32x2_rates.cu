// fp32x2_rates.cu -- issue/pipe rates of the instructions the streaming Jacobi kernel is made of, on sm_100a:
// scalar FADD vs packed FADD2 (add.rn.f32x2), SHFL, and the mixes the two kernel layouts would issue per row.
// Evidence for DESIGN.md section 5.1 (packed fp32x2 decision).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float add1(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

constexpr int kIters = 2048;
// MODE 0: 16 scalar FADD / iter (8 chains x 2); 1: 8 FADD2 / iter (same flops); 2: 16 FADD + 2 SHFL; 3: 8 FADD2 + 2 SHFL + 2 MOV-ish;
// 4: 8 FADD2 + 4 SHFL; 5: SHFL only (8 / iter)
template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, int lanesrc) {
    float f[16];
    u64 p[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = (float)(threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = ((u64)__float_as_uint((float)(threadIdx.x + i)) << 32) | __float_as_uint((float)i);
    float s0 = f[0], s1 = f[1], s2 = f[2], s3 = f[3];
    for (int it = 0; it < kIters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = add1(f[i], f[(i + 1) & 15]);
        }
        if (MODE == 1 || MODE == 3 || MODE == 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = add2(p[i], p[(i + 1) & 7]);
        }
        if (MODE == 2 || MODE == 3) {
            s0 = __shfl_up_sync(0xffffffffu, s0, 1);
            s1 = __shfl_down_sync(0xffffffffu, s1, 1);
        }
        if (MODE == 4) {
            s0 = __shfl_sync(0xffffffffu, s0, lanesrc);
            s1 = __shfl_sync(0xffffffffu, s1, lanesrc);
            s2 = __shfl_sync(0xffffffffu, s2, lanesrc);
            s3 = __shfl_sync(0xffffffffu, s3, lanesrc);
        }
        if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                s0 = __shfl_up_sync(0xffffffffu, s0, 1);
                s1 = __shfl_down_sync(0xffffffffu, s1, 1);
                s2 = __shfl_up_sync(0xffffffffu, s2, 1);
                s3 = __shfl_down_sync(0xffffffffu, s3, 1);
            }
        }
    }
    float acc = s0 + s1 + s2 + s3;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += f[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, double flop_per_iter, int ctas_per_sm, float* out) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<sms * ctas_per_sm, 128>>>(out, 3);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<sms * ctas_per_sm, 128>>>(out, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double warps_per_smsp = ctas_per_sm * 4 / 4.0;
    const double cyc = ms * 1e-3 * clk * 1e3;  // at the nominal max clock
    printf("{\"mode\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"cycles_per_iter_per_smsp\": %.2f, \"fp32_adds_per_clk_per_sm\": %.1f}\n",
           name, ctas_per_sm, ms, cyc / kIters / 1.0, flop_per_iter * 32 * warps_per_smsp * 4 * kIters / cyc);
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 128 * sizeof(float) * 4);
    for (int c : {1, 3, 4}) {
        run<0>("16xFADD", 16, c, out);
        run<1>("8xFADD2", 16, c, out);
        run<2>("16xFADD+2SHFL", 16, c, out);
        run<3>("8xFADD2+2SHFL", 16, c, out);
        run<4>("8xFADD2+4SHFL", 16, c, out);
        run<5>("8xSHFL", 0, c, out);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
