// helpers.hpp — host-side tile-size helpers, same names and argument lists as the reference's
// (reference: helpers.hpp:8-36: calculateSizeBlockQ / calculateSizeBlockKV sketch register- and L2-driven formulas
// and then return the constant 64; getNumCta asserts divisibility).
// Here they answer from the MEASURED sm_100a tile table the launcher dispatches from (fa_tile_table / fa_choose_tile in the C
// ABI: per head dim, causal flag and key-length bucket the kernel variant that measured fastest on B200), and getNumCta
// rounds ragged lengths up instead of asserting, because TMA zero-fill handles partial tiles.
#pragma once

#include <cuda_runtime.h>

#include "../include/fa_b200.h"

// Query rows handled by one CTA.  The fp32-I/O kernel is picked when the device cannot run the tcgen05 path's
// 16-bit tiles, i.e. never on B200; callers choose the dtype, this overload keeps the reference's signature
// and reports the bf16/fp16 tile (2 x 128 rows: two query tiles ping-ponged through the tensor pipe).
inline int calculateSizeBlockQ(cudaDeviceProp& prop, int d_head) {
    (void)prop;
    return fa_block_q(d_head, FA_DTYPE_BF16);
}

// Key/value rows per pipeline stage (one TMA box pair, 128 rows).  `device` kept for signature compatibility.
inline int calculateSizeBlockKV(cudaDeviceProp& prop, int d_head, int device) {
    (void)prop;
    (void)device;
    return fa_block_kv(d_head, FA_DTYPE_BF16);
}

// What the launcher would run for a given problem: the row of the measured tile table (fa_tile_table) that applies.
inline fa_tile_choice_t chooseTile(int d_head, int dtype, bool causal, int seqLenQ, int seqLenK) {
    fa_tile_choice_t t{};
    fa_choose_tile(d_head, dtype, causal ? 1 : 0, seqLenQ, seqLenK, &t);
    return t;
}

// dtype-aware variants (not in the reference)
inline int calculateSizeBlockQ(int d_head, int dtype) { return fa_block_q(d_head, dtype); }
inline int calculateSizeBlockKV(int d_head, int dtype) { return fa_block_kv(d_head, dtype); }

// CTAs along the query axis of one (batch, head): ceil(q_dim / q_block_size); 0 for empty input.
inline int getNumCta(int q_dim, int q_block_size) { return fa_num_cta(q_dim, q_block_size); }
