// DMMA (mma.sync.m8n8k4.f64) issue rate as a function of warps per SM sub-partition and of the number of
// independent accumulator tiles per warp: can one or two warps per scheduler keep the FP64 tensor pipe busy?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_occ dmma_occ.cu && ./dmma_occ
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int sms, double* out, double clk_ghz) {
    for (int wps = 1; wps <= 4; wps++) {               // warps per sub-partition
        int threads = 128 * wps, iters = 40000 / ILP * 8;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<ILP><<<sms, threads>>>(out, 100);
        cudaEventRecord(e0);
        k<ILP><<<sms, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double dmma_per_subpart = (double)ILP * iters * wps;
        double cyc = ms * 1e-3 * clk_ghz * 1e9;
        printf("ILP %2d  warps/sub-partition %d: %.2f cycles per DMMA per sub-partition, %.2f TFLOP/s\n", ILP, wps, cyc / dmma_per_subpart,
               512.0 * dmma_per_subpart * 4 * sms / (ms * 1e-3) * 1e-12);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, 1 << 24);
    double clk = p.clockRate * 1e-6;
    printf("%s SMs %d clock %.3f GHz\n", p.name, p.multiProcessorCount, clk);
    run<1>(p.multiProcessorCount, out, clk);
    run<2>(p.multiProcessorCount, out, clk);
    run<4>(p.multiProcessorCount, out, clk);
    run<8>(p.multiProcessorCount, out, clk);
    run<20>(p.multiProcessorCount, out, clk);
    return 0;
}
