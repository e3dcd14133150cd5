// Host helper: encode a bf16 tiled TMA descriptor with 128-byte swizzle (zero fill out of bounds).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace vtc {
// dims / box are innermost-first (dims[0] = contiguous dimension, in elements); strides_bytes[i] is the byte
// stride of dimension i+1 (rank-1 entries).
int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);
// same for fp32 elements (the epilogue's TMA store / reduce-add of the residual stream)
int make_tmap_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);
}  // namespace vtc
