// Dependent-chain latencies on sm_100a (perf triage for K3's frame step): cycles per op, one warp.
// Every chain step is an asm volatile touching the chain value, and the clock reads take the chain value as an
// input, so neither can be moved across the other.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
#define CLK(t, a) asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "d"(a) : "memory")
__global__ void k(double* out, long long* cyc, double a0, double inc, float f0) {
    double a = a0, b = a0 * 0.5;
    long long t0, t1;
    CLK(t0, a);
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("add.f64 %0, %0, %1;" : "+d"(a) : "d"(inc));
    CLK(t1, a); cyc[0] = t1 - t0;
    CLK(t0, a);
#pragma unroll 16
    for (int i = 0; i < N; ++i)
        asm volatile("{.reg .pred p; setp.gt.f64 p, %0, %1; selp.f64 %0, %1, %0, p; neg.f64 %1, %1;}" : "+d"(a), "+d"(b));
    CLK(t1, a); cyc[1] = t1 - t0;
    CLK(t0, a);
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        int lo = __double2loint(a), hi = __double2hiint(a);
        asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(lo));
        asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(hi));
        a = __hiloint2double(hi, lo);
    }
    CLK(t1, a); cyc[2] = t1 - t0;
    float f = f0;
    CLK(t0, a);
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("add.f32 %0, %0, %1;" : "+f"(f) : "f"(f0));
    a += f;
    CLK(t1, a); cyc[3] = t1 - t0;
    CLK(t0, a);
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        asm volatile("add.f64 %0, %0, %1;" : "+d"(a) : "d"(inc));
        int lo = __double2loint(a), hi = __double2hiint(a);
        asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(lo));
        asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(hi));
        a = __hiloint2double(hi, lo);
    }
    CLK(t1, a); cyc[4] = t1 - t0;
    CLK(t0, a);
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("{.reg .pred p; setp.gt.f64 p, %0, %1; selp.f64 %0, %0, %1, p;}" : "+d"(a) : "d"(b));
    CLK(t1, a); cyc[5] = t1 - t0;
    CLK(t0, a);
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("{.reg .f32 t; cvt.rn.f32.f64 t, %0; cvt.f64.f32 %0, t;}" : "+d"(a));
    CLK(t1, a); cyc[6] = t1 - t0;
    // shared-memory round trip: st.volatile + ld.volatile of the chain value
    __shared__ double slot[32];
    unsigned sa = (unsigned)__cvta_generic_to_shared(&slot[threadIdx.x]);
    CLK(t0, a);
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        asm volatile("st.volatile.shared.f64 [%1], %0;" :: "d"(a), "r"(sa));
        asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(a) : "r"(sa));
    }
    CLK(t1, a); cyc[7] = t1 - t0;
    out[threadIdx.x] = a + b + f;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 32 * 8); cudaMalloc(&c, 8 * 8);
    for (int r = 0; r < 3; ++r) k<<<1, 32>>>(d, c, 1.5, 1e-3, 1.0f);
    long long h[8]; cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
    const char* nm[8] = {"DADD", "DSETP+SELP (+DNEG)", "SHFL.UP x2 (64-bit)", "FADD", "DADD -> SHFL64", "DSETP+SELP",
                         "F2F f64->f32->f64", "STS.64 -> LDS.64 (volatile)"};
    for (int i = 0; i < 8; ++i) printf("%-28s %.1f cycles/iter\n", nm[i], (double)h[i] / N);
    return 0;
}
