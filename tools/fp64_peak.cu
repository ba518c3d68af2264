// FP64 pipe micro-benchmark for B200 (sm_100a): DMMA.8x8x4 peak, DFMA peak, and whether the two
// pipes overlap.  Output: one JSON object on stdout.  Used to set the FP64 roofline denominator
// (MEASURED_PEAKS.json has no fp64 entry).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__device__ __forceinline__ void dmma_loop(double* out, const double* in, int iters) {
    double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = 0; c[i][1] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[threadIdx.x] = s;
}

template <int ILP>
__device__ __forceinline__ void dfma_loop(double* out, const double* in, int iters) {
    double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i];
    if (s == 123.456) out[threadIdx.x] = s;
}

__device__ __forceinline__ void exp_loop(double* out, const double* in, int iters) {
    double x = in[threadIdx.x & 31] * 1e-3;
    double s = 0;
    for (int it = 0; it < iters; it++) { s += exp(x); x += 1e-7; }
    if (s == 123.456) out[threadIdx.x] = s;
}

// mode 0: all warps DMMA; 1: all warps DFMA; 2: even warps DMMA, odd warps DFMA; 3: exp
__global__ void k(double* out, const double* in, int iters_mma, int iters_fma, int mode) {
    int w = threadIdx.x >> 5;
    if (mode == 0 || (mode == 2 && (w >> 2 & 1) == 0)) dmma_loop<16>(out, in, iters_mma);
    else if (mode == 1 || mode == 2) dfma_loop<16>(out, in, iters_fma);
    else exp_loop(out, in, iters_fma);
}

static float run(int blocks, int threads, int im, int ifm, int mode, double* out, double* in) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<<<blocks, threads>>>(out, in, im / 8 + 1, ifm / 8 + 1, mode);   // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0));
        k<<<blocks, threads>>>(out, in, im, ifm, mode);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double *in, *out; CK(cudaMalloc(&in, 1024)); CK(cudaMalloc(&out, 1024 * 8)); CK(cudaMemset(in, 0, 1024));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);
    const int IT = 20000;
    for (int wps = 4; wps <= 32; wps *= 2) {   // warps per SM (1 block per SM)
        float ms = run(sms, wps * 32, IT, IT, 0, out, in);
        double fl = (double)sms * wps * IT * 16 * 512.0;
        printf(", \"dmma_tflops_w%d\": %.3f", wps, fl / ms * 1e-9);
    }
    for (int wps = 4; wps <= 32; wps *= 2) {
        float ms = run(sms, wps * 32, IT, IT * 8, 1, out, in);
        double fl = (double)sms * wps * (IT * 8.0) * 16 * 64.0;
        printf(", \"dfma_tflops_w%d\": %.3f", wps, fl / ms * 1e-9);
    }
    {   // overlap test: 16 warps/SM, warps 0-3,8-11 DMMA, 4-7,12-15 DFMA; each sized to take ~equal time alone
        float a = run(sms, 8 * 32, IT, IT, 0, out, in);          // 8 DMMA warps alone
        float b = run(sms, 8 * 32, IT, IT * 8, 1, out, in);      // 8 DFMA warps alone
        float c = run(sms, 16 * 32, IT, IT * 8, 2, out, in);     // both
        printf(", \"overlap\": {\"dmma8_ms\": %.3f, \"dfma8_ms\": %.3f, \"both16_ms\": %.3f}", a, b, c);
    }
    {
        float ms = run(sms, 16 * 32, IT, IT, 3, out, in);
        printf(", \"exp_per_ns\": %.3f", (double)sms * 16 * 32 * IT / (ms * 1e6));
    }
    printf("}\n");
    return 0;
}
