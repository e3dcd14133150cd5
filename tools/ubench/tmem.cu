// Micro-benchmark (debug tool, not part of libvtc): TMEM load/store throughput per SM for the shapes the attention
// softmax uses, alone and mixed with MUFU.EX2.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void ldwait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void stwait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE 0: LDTM.x32 + wait each; 1: two LDTM.x32 in flight; 2: LDTM.x32 + 32 MUFU on the previous data (prefetch);
// 3: 32 MUFU only; 4: STTM.x16 only; 5: mode 2 + STTM.x16 + 16 FMNMX3-ish + adds (softmax-like)
template <int MODE>
__global__ void k(float* out, int iters, unsigned long long* cyc) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t a[32], b[32];
    for (int i = 0; i < 32; ++i) { a[i] = __float_as_uint(-1.0f - i * 0.01f); b[i] = a[i]; }
    float acc = 0.f;
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t col = (it * 32) & 255;
        if (MODE == 0) { ld32(base + col, a); ldwait(); acc += __uint_as_float(a[it & 31]); }
        if (MODE == 1) { ld32(base + col, a); ld32(base + ((col + 32) & 255), b); ldwait(); acc += __uint_as_float(a[it & 31]) + __uint_as_float(b[it & 31]); }
        if (MODE == 2 || MODE == 5) {
            ld32(base + col, (it & 1) ? a : b);                 // prefetch next
            const uint32_t (&c)[32] = (it & 1) ? b : a;
            float e[32];
            float mx = -1e30f;
            if (MODE == 5) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(c[i]), __uint_as_float(c[i + 1])));
                if (__any_sync(0xffffffffu, mx > 1e20f)) acc += 1.f;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) { float x = __uint_as_float(c[i]) * 0.5f - (MODE == 5 ? mx * 1e-30f : 0.f); asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[i]) : "f"(x)); }
            if (MODE == 5) {
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 32; i += 2) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i >> 1]) : "f"(e[i + 1]), "f"(e[i])); acc += e[i] + e[i + 1]; }
                st16(base + 256 + ((it * 16) & 127), pk);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) acc += e[i];
            }
            ldwait();
        }
        if (MODE == 3) {
            float e[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) { float x = __uint_as_float(a[i]) * 0.5f; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[i]) : "f"(x)); }
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += e[i];
            a[it & 31] = __float_as_uint(acc * 1e-30f - 1.0f);
        }
        if (MODE == 4) { st16(base + col, a); stwait(); }
    }
    unsigned long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(a[3]) + __uint_as_float(b[5]);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512u) : "memory");
}

template <int MODE>
void run(const char* name, int loads_per_iter, int bytes_per_load) {
    float* out; unsigned long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    for (int warps : {4, 8, 16}) {
        k<MODE><<<148, warps * 32>>>(out, iters, cyc);
        k<MODE><<<148, warps * 32>>>(out, iters, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-44s warps/SM %2d: %7.1f clk per iter per warp, %6.1f clk per iter per SMSP, TMEM %6.1f B/clk/SM  (%s)\n", name, warps, (double)h / iters,
               (double)h / iters / (warps / 4.0), (double)loads_per_iter * bytes_per_load * warps * iters / h, cudaGetErrorString(e));
    }
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("LDTM.x32 + wait", 1, 4096);
    run<1>("2 x LDTM.x32 in flight + wait", 2, 4096);
    run<3>("32 x MUFU.EX2 (+32 FADD)", 0, 0);
    run<2>("LDTM.x32 prefetch + 32 MUFU + 32 FADD", 1, 4096);
    run<4>("STTM.x16 + wait", 1, 2048);
    run<5>("softmax-like chunk (LDTM, max, 32 MUFU, pack, STTM)", 1, 4096);
    return 0;
}
