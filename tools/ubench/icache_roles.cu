// Instruction-supply micro-benchmark 2: THREE different instruction streams per SM (the role kernel's situation).  12 warps per block, one block per
// SM; warp w runs role w % 3 (so that every scheduler hosts one role: warp w sits on scheduler w % 4 ... here role = (w / 4) keeps roles on
// DIFFERENT schedulers, role = w % 3 mixes them) and loops `iters` times over its role's body of N FFMAs.  Variants: all warps start together vs the
// four warps of a role staggered by D instructions (do warps that drift apart still share the fetched lines?).
#include <cstdio>
#include <cuda_runtime.h>

template <int N, int TAG> __device__ __forceinline__ void body(float (&a)[8], float b, float c) {
#pragma unroll
    for (int k = 0; k < N; k++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[(k + TAG) & 7]) : "f"(b), "f"(c + TAG));
}

template <int N> __global__ void __launch_bounds__(384, 1) kern(int iters, float b, float c, float* out, long long* cyc, int stagger, int role_by_mod) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x + k;
    const int w = threadIdx.x >> 5;
    const int role = role_by_mod ? (w % 3) : (w / 4);
    const int idx = role_by_mod ? (w / 3) : (w % 4);         // index of this warp among the four warps of its role
    for (int s = 0; s < idx * stagger; s++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[0]) : "f"(b), "f"(c));
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (role == 0) body<N, 0>(a, b, c);
        else if (role == 1) body<N, 1>(a, b, c);
        else body<N, 2>(a, b, c);
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 12 + w] = t1 - t0;
}

template <int N> void run(int stagger, int role_by_mod) {
    const int blocks = 148, iters = 8;
    float* out; long long* cyc;
    cudaMalloc(&out, blocks * 384 * sizeof(float));
    cudaMalloc(&cyc, blocks * 12 * sizeof(long long));
    for (int rep = 0; rep < 2; rep++) kern<N><<<blocks, 384>>>(iters, 1.0001f, 0.5f, out, cyc, stagger, role_by_mod);
    cudaDeviceSynchronize();
    long long h[148 * 12];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double av = 0, mx = 0;
    for (int i = 0; i < 148 * 12; i++) { av += h[i]; if (h[i] > mx) mx = h[i]; }
    av /= 148 * 12;
    const double ninst = (double)iters * N;
    printf("3 roles x %5d instr (%5.1f KB each) x 4 warps, stagger %5d, roles %s: cyc/inst per warp avg %5.2f max %5.2f  SM IPC %4.2f\n", N, N * 16 / 1024.0,
           stagger, role_by_mod ? "mixed on schedulers" : "one per scheduler ", av / ninst, mx / ninst, 12 * ninst / mx);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int mod = 0; mod < 2; mod++) {
        run<512>(0, mod); run<512>(300, mod);
        run<1024>(0, mod); run<1024>(300, mod); run<1024>(2000, mod);
        run<1792>(0, mod); run<1792>(100, mod); run<1792>(300, mod); run<1792>(1000, mod); run<1792>(5000, mod);
        run<3584>(0, mod); run<3584>(300, mod); run<3584>(5000, mod);
    }
    return 0;
}
