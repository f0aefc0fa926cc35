// Diagnostics: latency of dependent warp-level operations for ONE warp alone on an SM -- what the longest cascade of a launch
// pays per step of its dependency chain.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/lat scripts/lone_warp_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
template <int OP> __global__ void k(unsigned* out, long long* cyc, unsigned seed) {
    unsigned v = seed + threadIdx.x;
    __shared__ unsigned sm[64];
    sm[threadIdx.x & 63] = v;
    __syncwarp();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        if (OP == 0) v = __ballot_sync(0xffffffffu, v & 1) + i;
        else if (OP == 1) v = __shfl_sync(0xffffffffu, v, (v + 1) & 31) + i;
        else if (OP == 2) v = __reduce_add_sync(0xffffffffu, v) + i;
        else if (OP == 3) v = __reduce_max_sync(0xffffffffu, v) + i;
        else if (OP == 4) v = v * 2654435761u + i;
        else if (OP == 5) v = sm[v & 63] + i;
        else if (OP == 6) v = __shfl_down_sync(0xffffffffu, v, 1) + i;
        else if (OP == 7) v = __popc(v) + (v << 3) + i;
        else if (OP == 8) { v = __umulhi(v, 0xD2511F53u) ^ (v * 0xCD9E8D57u); }
        else if (OP == 9) { __syncwarp(); v += i; }
        else if (OP == 10) { v = __reduce_or_sync(0xffffffffu, v) + i; }
        else if (OP == 11) { v = __ffs(v) + (v >> 1) + i; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[OP] = t1 - t0;
}
int main() {
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 128);
    const char* names[] = {"ballot", "shfl idx", "redux add", "redux max", "imad", "lds", "shfl_down", "popc+shift", "umulhi^mul", "syncwarp", "redux or", "ffs"};
#define RUN(OP) k<OP><<<1, 32>>>(out, cyc, 1); k<OP><<<1, 32>>>(out, cyc, 1); cudaDeviceSynchronize(); printf("%-12s %.1f cycles per dependent iteration\n", names[OP], (double)cyc[OP] / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11)
    return 0;
}
