// Instruction-cache capacity probe (B200): a loop whose body is KB kilobytes of straight-line SASS (independent FFMAs, 16 B
// each), run by one warp per CTA, `warps` CTAs per SM.  Prints cycles per instruction against body size: the knees are the
// SM-level instruction cache (ICC) and the GPC-level L1.5 (GCC).  Used to size the closed-loop kernel's hot footprint.
#include <cstdio>
#include <cuda_runtime.h>
#define R4(x) x x x x
#define R16(x) R4(R4(x))
#define R64(x) R4(R16(x))
#define BODY64 R64(a = fmaf(a, b, c); c = fmaf(c, b, a);)   /* 128 FFMA = 2 KB */
template <int KB2>   // body = KB2 * 2 KB
__global__ void probe(float *out, int iters, long long *cyc) {
    float a = threadIdx.x, b = 1.0001f, c = 0.5f;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < KB2; ++r) { BODY64 }
    }
    long long t1 = clock64();
    out[blockIdx.x * 32 + threadIdx.x] = a + c;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int KB2>
void run(int warps_per_sm, int nsm) {
    float *out; long long *cyc;
    int grid = warps_per_sm * nsm;
    cudaMalloc(&out, grid * 32 * 4); cudaMalloc(&cyc, grid * 8);
    int iters = 4096 / KB2 + 8;
    probe<KB2><<<grid, 32>>>(out, 4, cyc);
    probe<KB2><<<grid, 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long *h = new long long[grid];
    cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < grid; ++i) s += h[i];
    printf("body %4d KB  warps/SM %2d  cycles/instr/warp %.3f\n", KB2 * 2, warps_per_sm, s / grid / ((double)iters * KB2 * 128));
    cudaFree(out); cudaFree(cyc); delete[] h;
}
int main() {
    for (int w : {1, 4, 8}) {
        run<2>(w, 148); run<4>(w, 148); run<8>(w, 148); run<12>(w, 148); run<16>(w, 148); run<20>(w, 148); run<24>(w, 148);
        run<32>(w, 148); run<40>(w, 148); run<48>(w, 148); run<64>(w, 148); run<96>(w, 148);
    }
    return 0;
}
