// Micro-benchmark (debug tool, not part of libvtc): per-SM throughput of the instructions in the attention softmax.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipes tools/ubench/pipes.cu && /tmp/pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float* out, int iters, unsigned long long* cyc) {
    float a[8];
    uint64_t p[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i * 0.1f - 3.0f; p[i] = ((uint64_t)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.0f); }
    const uint64_t c2 = ((uint64_t)__float_as_uint(0.999f) << 32) | __float_as_uint(0.999f);
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
            if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
            if (OP == 3) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); a[i] = __uint_as_float(r << 16); }
            if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(0.999f));
            if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 3) & 7]));
            if (OP == 6) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 7) asm volatile("fma.rn.f32 %0, %0, %1, 0f3E800000;" : "+f"(a[i]) : "f"(a[(i + 3) & 7]));          // immediate addend (Horner step)
            if (OP == 8) asm volatile("{.reg .b64 c; mov.b64 c, {0f3E800000, 0f3E800000}; fma.rn.f32x2 %0, %0, %1, c;}" : "+l"(p[i]) : "l"(p[(i + 3) & 7]));
            if (OP == 9) asm volatile("add.rn.f32 %0, %0, 0f4B400000;" : "+f"(a[i]));
            if (OP == 10) { uint32_t r = __float_as_uint(a[i]); asm volatile("mad.lo.u32 %0, %1, 8388608, %0;" : "+r"(r) : "r"(__float_as_uint(a[(i + 3) & 7]))); a[i] = __uint_as_float(r); }
        }
    }
    unsigned long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((uint32_t)p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name) {
    float* out; unsigned long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int warps : {4, 8, 16, 32}) {
        k<OP><<<148, warps * 32>>>(out, iters, cyc);
        k<OP><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        unsigned long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double ops = (double)iters * 8 * warps * 32;
        printf("%-22s warps/SM %2d: %.2f lane-ops/clk/SM  (%.2f clk per warp-instr per SMSP)\n", name, warps, ops / h, (double)h / (iters * 8.0 * warps / 4));
    }
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("MUFU.EX2");
    run<6>("MUFU.RCP");
    run<1>("FFMA2 (f32x2)");
    run<2>("FADD2 (f32x2)");
    run<3>("F2FP bf16x2 pack");
    run<4>("FFMA scalar");
    run<5>("FMNMX");
    run<7>("FFMA scalar, imm addend");
    run<8>("FFMA2, imm addend");
    run<9>("FADD scalar, imm");
    run<10>("IMAD imm");
    return 0;
}
