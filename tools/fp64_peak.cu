// FP64 pipe microbenchmarks for B200 (sm_100a): DFMA, DMMA m8n8k4 and m16n8k16 issue rates.
// Used once to choose between the DFMA and DMMA inner kernels and as a roofline denominator.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void dmma884_kernel(double* out, int iters) {
    double c[8][2];
    for (int k = 0; k < 8; ++k) { c[k][0] = 0; c[k][1] = 0; }
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma16816_kernel(double* out, int iters) {
    double c[4][4];
    for (int k = 0; k < 4; ++k) for (int j = 0; j < 4; ++j) c[k][j] = 0;
    double a[8], b[4];
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3 + j;
    for (int j = 0; j < 4; ++j) b[j] = 1.0 + threadIdx.x * 1e-6 + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[k][0]), "+d"(c[k][1]), "+d"(c[k][2]), "+d"(c[k][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
    for (int k = 0; k < 4; ++k) for (int j = 0; j < 4; ++j) s += c[k][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_it(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    for (int wps = 4; wps <= 32; wps *= 2) {   // warps per SM
        int threads = 32 * (wps > 8 ? 8 : wps), blocks = sms * (wps > 8 ? wps / 8 : 1);
        int iters = 20000;
        float ms = time_it([&] { dfma_kernel<<<blocks, threads>>>(out, iters); }, 5);
        double fl = 2.0 * 8 * iters * (double)threads * blocks;
        printf("DFMA       warps/SM %2d : %.2f TFLOP/s\n", wps, fl / ms * 1e-9);
        ms = time_it([&] { dmma884_kernel<<<blocks, threads>>>(out, iters / 4); }, 5);
        fl = 2.0 * 8 * 8 * 4 * 8 * (iters / 4) * (double)(threads / 32) * blocks;
        printf("DMMA 8x8x4 warps/SM %2d : %.2f TFLOP/s\n", wps, fl / ms * 1e-9);
        ms = time_it([&] { dmma16816_kernel<<<blocks, threads>>>(out, iters / 8); }, 5);
        fl = 2.0 * 16 * 8 * 16 * 4 * (iters / 8) * (double)(threads / 32) * blocks;
        printf("DMMA16x8x16 warps/SM %2d : %.2f TFLOP/s\n", wps, fl / ms * 1e-9);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
