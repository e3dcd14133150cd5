// Micro-benchmark (debug tool, not part of libvtc): round-trip time of the two tensor-core steps of one attention work
// item -- S = Q K^T (4 x M128 N208 K16, operands in shared memory) and O = P V (13 x M128 N64 K16) with P as a TMEM or
// shared-memory A operand and V MN-major or K-major -- alone and while other warps keep the TMEM read port busy with
// tcgen05.ld.x32, the way the other group's softmax does in attention_cs.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../vision_transformer_cam_b200/csrc -o mma_bench mma.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace vtc;

constexpr int HAMMER_MAX = 16;
constexpr int THREADS = (4 + HAMMER_MAX) * 32;
constexpr int SMEM = 16384 + 4 * 16384 + 4 * 16384 + 1024;      // Q | K or P blocks | V blocks | barrier

// MODE 0: QK SS (4 MMAs N=208)   1: PV TS, V MN-major   2: PV SS (P in smem), V MN-major   3: PV TS, V K-major
//      4: PV TS MN-major issued as 2 x N=32 halves? (not used)   5: QK then PV(TS) back to back, one commit
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) k(int iters, int batch, int hammer, unsigned long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    __shared__ int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < SMEM / 4; i += THREADS) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
    if (warp == 0) tmem_alloc(&tptr, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tptr;
    const uint32_t q_addr = smem_u32(smem), k_addr = q_addr + 16384, v_addr = k_addr + 4 * 16384;
    if (warp == 0) {
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc_bf16(128, 208, 0, 0);
            const uint32_t idesc_o_mn = make_idesc_bf16(128, 64, 0, 1), idesc_o_k = make_idesc_bf16(128, 64, 0, 0);
            uint32_t ph = 0;
            unsigned long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                for (int bt = 0; bt < batch; ++bt) {
                    if (MODE == 0 || MODE == 5) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16(tb, make_smem_desc_sw128(q_addr + kk * 32, 1024, 16), make_smem_desc_sw128(k_addr + kk * 32, 1024, 16), idesc_s, kk != 0);
                    }
                    if (MODE == 1 || MODE == 5) {
                        for (int ks = 0; ks < 13; ++ks)
                            umma_bf16_ts(tb + 192, tb + 16 + 8 * ks, make_smem_desc_sw128(v_addr + ks * 2048, 1024, 1024), idesc_o_mn, ks != 0);
                    }
                    if (MODE == 2) {
                        for (int ks = 0; ks < 13; ++ks)
                            umma_bf16(tb + 192, make_smem_desc_sw128(k_addr + (ks >> 2) * 16384 + (ks & 3) * 32, 1024, 16),
                                      make_smem_desc_sw128(v_addr + ks * 2048, 1024, 1024), idesc_o_mn, ks != 0);
                    }
                    if (MODE == 3) {
                        for (int ks = 0; ks < 13; ++ks)
                            umma_bf16_ts(tb + 192, tb + 16 + 8 * ks, make_smem_desc_sw128(v_addr + (ks >> 2) * 8192 + (ks & 3) * 32, 1024, 16), idesc_o_k, ks != 0);
                    }
                }
                umma_commit(&bar);
                mbar_wait_fast(&bar, ph);
                ph ^= 1;
                tc_fence_after();
            }
            unsigned long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = (t1 - t0) / iters;
            *reinterpret_cast<volatile int*>(&stop) = 1;
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 4 + hammer) {
        // keep the TMEM read port of this warp's sub-partition busy (columns 256.. are not touched by the MMAs)
        const uint32_t base = tb + 256 + ((uint32_t)((warp & 3) * 32) << 16);
        float acc = 0.f;
        unsigned long long n = 0;
        while (*reinterpret_cast<volatile int*>(&stop) == 0) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(base + ((n & 3) * 32), r);
            tmem_ld_wait();
            acc += __uint_as_float(r[n & 31]);
            ++n;
        }
        if (acc == 123.456f) out[1] = n;
        if (blockIdx.x == 0 && lane == 0 && warp == 4) out[2] = n;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <int MODE>
void run(const char* name, unsigned long long* d) {
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    for (int batch : {1, 8}) {
        for (int hammer : {0, 4, 8, 16}) {
            unsigned long long h[3] = {0, 0, 0};
            cudaMemset(d, 0, 24);
            k<MODE><<<148, THREADS, SMEM>>>(200, batch, hammer, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
            cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
            printf("%-28s batch %d hammer warps %2d: %7llu clk per round trip (%6.0f per batch)   hammer loads %llu\n", name, batch, hammer, h[0],
                   (double)h[0] / batch, h[2]);
        }
    }
}

int main() {
    unsigned long long* d;
    cudaMalloc(&d, 64);
    run<0>("QK SS N=208 x4", d);
    run<1>("PV TS V MN-major x13", d);
    run<2>("PV SS V MN-major x13", d);
    run<3>("PV TS V K-major x13", d);
    run<5>("QK + PV TS", d);
    return 0;
}
