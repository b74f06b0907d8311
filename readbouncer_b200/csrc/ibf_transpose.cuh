// ibf_transpose.cuh -- 32x32 bit-matrix transpose in registers (host/device).
//
// Used by the column build (ibf_insert.cu): bins are built as bit columns (bit r of a column = row r)
// and then turned into the reference's row-interleaved layout, 32 bins x 32
// rows at a time.  Bit numbering is little-endian: after the call, bit i of x[j] is what bit j of
// x[i] was.  Five butterfly stages of 16 masked swaps each (80 swaps of 3 logic ops).
#pragma once

#include "ibf_common.cuh"

namespace rb {

RB_HD void transpose32(uint32_t (&x)[32])
{
    uint32_t m = 0x0000FFFFu;
#pragma unroll
    for (int j = 16; j != 0; j >>= 1, m ^= m << j) {
#pragma unroll
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            // swap the high-j bits of x[k] with the low-j bits of x[k + j] (block-wise off-diagonal exchange)
            const uint32_t t = ((x[k] >> j) ^ x[k + j]) & m;
            x[k] ^= t << j;
            x[k + j] ^= t;
        }
    }
}

}  // namespace rb
