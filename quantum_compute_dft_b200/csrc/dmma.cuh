// dmma.cuh -- FP64 tensor-core primitive and fragment conventions.
//
// On sm_100a the only FP64 tensor path is the warp-synchronous `mma.sync ... f64`, which ptxas
// lowers to SASS `DMMA.8x8x4` with register accumulators (there is no tcgen05 kind for f64 and
// TMEM is not involved).  Fragment ownership for m8n8k4, lane = 0..31:
//   A (8x4, row-major)  : lane holds A[lane/4][lane%4]
//   B (4x8, col-major)  : lane holds B[lane%4][lane/4]
//   C/D (8x8)           : lane holds C[lane/4][2*(lane%4) + {0,1}]
#pragma once

namespace dmma {

__device__ __forceinline__ void mma8x8x4(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

}  // namespace dmma
