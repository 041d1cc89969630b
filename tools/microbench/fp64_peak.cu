// Microbenchmark: FP64 DFMA vs DMMA (mma.sync f64) register-resident peak on sm_100a.
// Used once to pick the roofline denominator for the dense contractions (MEASURED_PEAKS.json has no FP64 entry).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
    double a[16];
    double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.999999;
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = i * 0.5;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_dmma884(double* out, int iters) {
    double c[NACC][2];
    double a = 1.0 + threadIdx.x * 1e-6, b = 0.5;
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma1688(double* out, int iters) {
    double c[NACC][4];
    double a[4] = {1.0 + threadIdx.x * 1e-6, 0.5, 0.25, 0.125}, b[2] = {0.5, 0.25};
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma16816(double* out, int iters) {
    double c[NACC][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = 1.0 + threadIdx.x * 1e-6 * i;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = 0.5 + i;
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device %s sm_%d%d SMs=%d clock=%d kHz smem/block optin=%zu L2=%d MB\n", p.name, p.major, p.minor, p.multiProcessorCount,
           p.clockRate, p.sharedMemPerBlockOptin, p.l2CacheSize >> 20);
    double* out;
    cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        const int threads = warps * 32, blocks = p.multiProcessorCount * (warps >= 32 ? 1 : 2);
        float ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
        printf("DFMA   warps/CTA=%2d CTAs=%d : %.3f ms  %.2f TFLOP/s\n", warps, blocks, ms, 2.0 * 16 * iters * threads * (double)blocks / ms / 1e9);
        ms = timeit([&] { k_dmma884<8><<<blocks, threads>>>(out, iters); });
        printf("DMMA884  x8  warps/CTA=%2d : %.3f ms  %.2f TFLOP/s\n", warps, ms, 512.0 * 8 * iters * warps * (double)blocks / ms / 1e9);
        ms = timeit([&] { k_dmma884<16><<<blocks, threads>>>(out, iters); });
        printf("DMMA884  x16 warps/CTA=%2d : %.3f ms  %.2f TFLOP/s\n", warps, ms, 512.0 * 16 * iters * warps * (double)blocks / ms / 1e9);
        ms = timeit([&] { k_dmma1688<8><<<blocks, threads>>>(out, iters); });
        printf("DMMA1688 x8  warps/CTA=%2d : %.3f ms  %.2f TFLOP/s\n", warps, ms, 2048.0 * 8 * iters * warps * (double)blocks / ms / 1e9);
        ms = timeit([&] { k_dmma16816<8><<<blocks, threads>>>(out, iters); });
        printf("DMMA16816 x8 warps/CTA=%2d : %.3f ms  %.2f TFLOP/s\n", warps, ms, 4096.0 * 8 * iters * warps * (double)blocks / ms / 1e9);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
