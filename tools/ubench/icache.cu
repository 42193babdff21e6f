// Instruction-supply micro-benchmark (B200): a loop whose body is N independent-ish FFMAs (8 accumulator chains), run by W warps per block,
// one block per SM.  Reports cycles per warp-instruction vs. body size and warps per scheduler: locates the L0 / L1.5 instruction-cache
// capacities and the fetch-bound issue rate that the step kernel runs into (profiles/r02_notes.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu && ./icache
#include <cstdio>
#include <cuda_runtime.h>

template <int N> __device__ __forceinline__ void body(float (&a)[8], float b, float c) {
#pragma unroll
    for (int k = 0; k < N; k++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k & 7]) : "f"(b), "f"(c));
}

template <int N> __global__ void kern(int iters, float b, float c, float* out, long long* cyc, int stagger) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x + k;
    const int w = threadIdx.x >> 5;
    // optional stagger: warps start the loop at different times so that they do not share fetched lines
    if (stagger) { for (int s = 0; s < w * stagger; s++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[0]) : "f"(b), "f"(c)); }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) body<N>(a, b, c);
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + w] = t1 - t0;
}

template <int N> void run(int warps, int stagger) {
    const int blocks = 148, total = 1 << 16;
    int iters = total / N; if (iters < 4) iters = 4;
    float* out; long long* cyc;
    cudaMalloc(&out, blocks * warps * 32 * sizeof(float));
    cudaMalloc(&cyc, blocks * warps * sizeof(long long));
    for (int rep = 0; rep < 2; rep++) kern<N><<<blocks, warps * 32>>>(iters, 1.0001f, 0.5f, out, cyc, stagger);
    cudaDeviceSynchronize();
    long long* h = new long long[blocks * warps];
    cudaMemcpy(h, cyc, blocks * warps * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0, av = 0;
    for (int i = 0; i < blocks * warps; i++) { av += h[i]; if (h[i] > mx) mx = h[i]; }
    av /= blocks * warps;
    const double ninst = (double)iters * N;
    printf("N=%5d (%6.1f KB) warps/SM=%2d stagger=%4d  cyc/inst per warp: avg %6.2f max %6.2f   SM IPC %5.2f\n", N, N * 16 / 1024.0, warps, stagger,
           av / ninst, mx / ninst, warps * ninst / mx);
    delete[] h; cudaFree(out); cudaFree(cyc);
}

int main() {
    int ws[] = {1, 4, 8, 16};
    for (int wi = 0; wi < 4; wi++) {
        const int w = ws[wi];
        run<64>(w, 0); run<128>(w, 0); run<256>(w, 0); run<320>(w, 0); run<384>(w, 0); run<512>(w, 0); run<768>(w, 0); run<1024>(w, 0);
        run<1536>(w, 0); run<2048>(w, 0); run<3072>(w, 0); run<4096>(w, 0); run<8192>(w, 0);
    }
    printf("--- staggered warps (each warp starts 977 instructions later than the previous one)\n");
    for (int wi = 1; wi < 4; wi++) {
        const int w = ws[wi];
        run<256>(w, 977); run<512>(w, 977); run<1024>(w, 977); run<2048>(w, 977); run<4096>(w, 977);
    }
    return 0;
}
