// Hardware microbenchmarks that set the roofline denominators and the tiling rules for the
// FP64 sweep kernels (B200, sm_100a).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
//   1. DFMA peak (FP64 FMA pipe)           -> fp64 roofline denominator
//   2. DMMA m8n8k4 f64 rate                -> is the legacy FP64 tensor path any faster?
//   3. LDS wavefront cost of broadcast patterns (LDS.64 / LDS.128) -> register-tile shapes
//   4. pinned H2D / D2H bandwidth          -> e2e expectations
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA and DFMA interleaved (ratio 1 DMMA : NF DFMA): do the two share one FP64 datapath?
template <int NF>
__global__ void mixed_kernel(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[8][2], f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; f[i] = i * 0.5; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
            for (int j = 0; j < NF; j++) f[(i + j) & 7] = fma(f[(i + j) & 7], 1.0000001, 1e-9);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// LDS pattern benchmark: every lane loads from smem at an address derived from `mode`.
// We count SM cycles per warp-level LDS instruction with `nw` warps resident.
template <int VEC>  // 1 = LDS.64, 2 = LDS.128
__global__ void lds_kernel(double* out, int iters, int mode, long long* cycles) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    int lane = threadIdx.x & 31;
    int idx;
    switch (mode) {
        case 0: idx = 0; break;                         // all lanes same address
        case 1: idx = (lane & 7) * VEC; break;          // 8 unique, repeated in every quarter warp
        case 2: idx = (lane >> 2) * VEC; break;         // 8 unique, each shared by 4 adjacent lanes
        case 3: idx = (lane >> 3) * VEC; break;         // 4 unique, one per quarter warp
        case 4: idx = lane * VEC; break;                // all distinct, contiguous
        case 5: idx = (lane & 15) * VEC; break;         // 16 unique, repeated per half warp
        case 6: idx = (lane >> 1) * VEC; break;         // 16 unique, shared by lane pairs
        case 7: idx = (lane & 3) * VEC; break;          // 4 unique repeated 8x
        default: idx = 0;
    }
    double acc0 = 0, acc1 = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            int off = idx + ((it + j) & 7) * 64 * VEC;   // keep within 4096 doubles, varying address
            if (VEC == 1) {
                double v;
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(sm + off)));
                acc0 += v;
            } else {
                double v0, v1;
                asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v0), "=d"(v1) : "r"((unsigned)__cvta_generic_to_shared(sm + off)));
                acc0 += v0; acc1 += v1;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs %d smem/SM %zu clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerMultiprocessor, prop.clockRate);
    int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 64 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    // 1. DFMA peak
    for (int threads : {128, 256, 512, 1024}) {
        for (int rep = 0; rep < 2; rep++) {
            int iters = 20000; const int ILP = 8;
            int blocks = sms * (2048 / threads);
            dfma_kernel<ILP><<<blocks, threads>>>(out, 1000, 1.0000001, 1e-9);
            CK(cudaEventRecord(e0));
            dfma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            double flops = 2.0 * ILP * iters * (double)blocks * threads;
            if (rep) printf("DFMA threads/blk %4d blocks %5d: %.2f TFLOP/s (%.3f ms)\n", threads, blocks, flops / ms * 1e-9, ms);
        }
    }
    // long DFMA run (sustained, ~2 s) to see power-cap clocks
    {
        int threads = 256, blocks = sms * 8, iters = 2000000; const int ILP = 8;
        CK(cudaEventRecord(e0));
        dfma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * ILP * (double)iters * (double)blocks * threads;
        printf("DFMA sustained: %.2f TFLOP/s over %.1f ms\n", flops / ms * 1e-9, ms);
    }
    // 2. DMMA
    for (int threads : {128, 256, 512}) {
        int iters = 20000; int blocks = sms * (2048 / threads);
        dmma_kernel<<<blocks, threads>>>(out, 1000);
        CK(cudaEventRecord(e0));
        dmma_kernel<<<blocks, threads>>>(out, iters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 256 * 8 * (double)iters * (double)blocks * (threads / 32);
        printf("DMMA m8n8k4 threads/blk %4d: %.2f TFLOP/s (%.3f ms)\n", threads, flops / ms * 1e-9, ms);
    }
    // 2b. DMMA + DFMA mixed: time relative to DMMA alone tells whether the pipes are shared
    {
        int threads = 256, blocks = sms * 8, iters = 20000;
        auto run = [&](auto kern, const char* name, int nf) {
            kern<<<blocks, threads>>>(out, 1000);
            CK(cudaEventRecord(e0));
            kern<<<blocks, threads>>>(out, iters);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            double dm = 2.0 * 256 * 8 * (double)iters * (double)blocks * (threads / 32);
            double df = 2.0 * nf * 8 * (double)iters * (double)blocks * threads;
            printf("%s: %.3f ms  DMMA %.2f TF + DFMA %.2f TF = %.2f TF\n", name, ms, dm / ms * 1e-9, df / ms * 1e-9, (dm + df) / ms * 1e-9);
        };
        run(mixed_kernel<0>, "mixed 1 DMMA : 0 DFMA", 0);
        run(mixed_kernel<2>, "mixed 1 DMMA : 2 DFMA", 2);
        run(mixed_kernel<4>, "mixed 1 DMMA : 4 DFMA", 4);
        run(mixed_kernel<8>, "mixed 1 DMMA : 8 DFMA", 8);
    }
    // 3. LDS patterns: 1 block on 1 SM, nw warps
    long long* dcyc; CK(cudaMalloc(&dcyc, 8));
    CK(cudaFuncSetAttribute(lds_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(lds_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int vec = 1; vec <= 2; vec++) {
        for (int mode = 0; mode < 8; mode++) {
            for (int nw : {4, 16}) {
                int iters = 2000; long long cyc;
                if (vec == 1) lds_kernel<1><<<1, nw * 32, 65536>>>(out, iters, mode, dcyc);
                else lds_kernel<2><<<1, nw * 32, 65536>>>(out, iters, mode, dcyc);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
                printf("LDS.%d mode %d nw %2d: %.2f SM-cycles per warp-instruction\n", vec * 64, mode, nw, (double)cyc / ((double)iters * 16 * nw));
            }
        }
    }
    // 4. PCIe pinned bandwidth
    {
        size_t bytes = 1ull << 30; void *h, *d;
        CK(cudaMallocHost(&h, bytes)); CK(cudaMalloc(&d, bytes));
        CK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("H2D pinned: %.1f GB/s\n", bytes / ms * 1e-6);
        CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); printf("D2H pinned: %.1f GB/s\n", bytes / ms * 1e-6);
    }
    return 0;
}
