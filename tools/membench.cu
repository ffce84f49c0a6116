// membench.cu -- how much HBM bandwidth does a "many-stream SoA" access pattern reach at a given
// occupancy?  NR read streams + NW write streams of doubles, one element per thread per stream.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/membench membench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int NR, int NW, int MINB>
__global__ void __launch_bounds__(256, MINB) k_stream(const double* __restrict__ in, double* __restrict__ out, long n, long stride) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v[NR];
#pragma unroll
    for (int k = 0; k < NR; k++) v[k] = __ldg(in + k * stride + i);
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < NR; k++) acc += v[k];
#pragma unroll
    for (int k = 0; k < NW; k++) out[k * stride + i] = acc + k;
}

// same, but the loads are consumed one after the other through a dependent chain of divisions
template <int NR, int NW, int MINB>
__global__ void __launch_bounds__(256, MINB) k_stream_dep(const double* __restrict__ in, double* __restrict__ out, long n, long stride) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = 1.0;
#pragma unroll 1
    for (int k = 0; k < NR; k++) acc = acc / (__ldg(in + k * stride + i) + 2.0);
#pragma unroll
    for (int k = 0; k < NW; k++) out[k * stride + i] = acc + k;
}

template <typename K>
void run(const char* name, K kern, int nr, int nw, const double* in, double* out, long n, long stride) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    int blocks = (int)((n + 255) / 256);
    for (int w = 0; w < 2; w++) kern<<<blocks, 256>>>(in, out, n, stride);
    cudaEventRecord(a);
    const int reps = 5;
    for (int r = 0; r < reps; r++) kern<<<blocks, 256>>>(in, out, n, stride);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    ms /= reps;
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, 0);
    printf("%-28s NR=%2d NW=%2d blocks/SM=%d  %.3f ms  %.0f GB/s  (%s)\n", name, nr, nw, nb, ms,
           (double)(nr + nw) * n * 8 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
    long n = argc > 1 ? atol(argv[1]) : 4096000;
    long stride = (n + 31) / 32 * 32;
    double *in, *out;
    cudaMalloc(&in, 64 * stride * 8);
    cudaMalloc(&out, 32 * stride * 8);
    cudaMemset(in, 0, 64 * stride * 8);
    run("stream<8,8> full occ", k_stream<8, 8, 1>, 8, 8, in, out, n, stride);
    run("stream<48,20> minb=1", k_stream<48, 20, 1>, 48, 20, in, out, n, stride);
    run("stream<48,20> minb=2", k_stream<48, 20, 2>, 48, 20, in, out, n, stride);
    run("stream<48,20> minb=4", k_stream<48, 20, 4>, 48, 20, in, out, n, stride);
    run("stream<48,20> minb=8", k_stream<48, 20, 8>, 48, 20, in, out, n, stride);
    run("stream<60,5> minb=2", k_stream<60, 5, 2>, 60, 5, in, out, n, stride);
    run("stream<16,8> minb=2", k_stream<16, 8, 2>, 16, 8, in, out, n, stride);
    run("stream<16,8> minb=8", k_stream<16, 8, 8>, 16, 8, in, out, n, stride);
    run("dep<48,20> minb=2", k_stream_dep<48, 20, 2>, 48, 20, in, out, n, stride);
    run("dep<48,20> minb=8", k_stream_dep<48, 20, 8>, 48, 20, in, out, n, stride);
    return 0;
}
