// Microbenchmark: DMMA fed from shared memory with the contraction kernels' fragment pattern, no global traffic.
// Isolates what the LDS + DMMA inner loop can reach (ceiling for k_rho / k_contract) from the pipeline around it.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int LDN = 136, KT = 32;
template <int MT, int NT, bool SYNC, bool SCALE>
__global__ void __launch_bounds__(256, 1) k(double* out, int iters) {
    extern __shared__ double sm[];
    double* As = sm;             // [32][136]
    double* Bs = sm + KT * LDN;  // [32][136]
    double* ds = sm + 2 * KT * LDN;
    for (int i = threadIdx.x; i < 2 * KT * LDN + KT; i += 256) sm[i] = 1.0 + (i % 7) * 1e-3;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int wm = (MT == 4) ? (warp & 3) : warp, wn = (MT == 4) ? (warp >> 2) : 0;
    double acc[MT * NT][2];
#pragma unroll
    for (int t = 0; t < MT * NT; t++) acc[t][0] = acc[t][1] = 0.0;
    for (int it = 0; it < iters; it++) {
        if (SYNC) __syncthreads();
#pragma unroll
        for (int kk = 0; kk < KT; kk += 4) {
            double a[MT], b[NT];
            const double dv = SCALE ? ds[kk + q] : 1.0;
#pragma unroll
            for (int mt = 0; mt < MT; mt++) a[mt] = As[(kk + q) * LDN + wm * (MT * 8) + mt * 8 + g] * dv;
#pragma unroll
            for (int nt = 0; nt < NT; nt++) b[nt] = Bs[(kk + q) * LDN + wn * (NT * 8) + nt * 8 + g];
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
                for (int nt = 0; nt < NT; nt++) dmma884(acc[mt * NT + nt][0], acc[mt * NT + nt][1], a[mt], b[nt]);
        }
    }
    double s = 0;
#pragma unroll
    for (int t = 0; t < MT * NT; t++) s += acc[t][0] + acc[t][1];
    out[blockIdx.x * 256 + threadIdx.x] = s;
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
template <int MT, int NT, bool SYNC, bool SCALE>
void run(const char* name, double* out) {
    const int iters = 4000, smem = (2 * KT * LDN + KT) * 8;
    cudaFuncSetAttribute(k<MT, NT, SYNC, SCALE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float ms = timeit([&] { k<MT, NT, SYNC, SCALE><<<148, 256, smem>>>(out, iters); });
    double flops = 148.0 * 8 * iters * 8 * (MT * NT) * 512.0;
    printf("%-44s %.3f ms  %.2f TFLOP/s (%.1f%% of 37.05)\n", name, ms, flops / ms / 1e9, 100 * flops / ms / 1e9 / 37.05);
}
int main() {
    double* out;
    cudaMalloc(&out, 148 * 256 * 8);
    run<4, 8, false, false>("4x2 warps, 32x64 tile, no sync", out);
    run<4, 8, true, false>("4x2 warps, 32x64 tile, sync/32k", out);
    run<2, 16, false, false>("8x1 warps, 16x128 tile, no sync", out);
    run<2, 16, true, false>("8x1 warps, 16x128 tile, sync/32k", out);
    run<2, 16, true, true>("8x1 warps, 16x128 tile, sync + A scaling", out);
    run<4, 8, true, true>("4x2 warps, 32x64 tile, sync + A scaling", out);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
