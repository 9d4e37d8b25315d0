// pg_debug.cuh -- index assertions for a debug build (-DPG_DEBUG_BOUNDS): compute-sanitizer is closed on the GPU pool,
// so the kernels that compute their own addresses (the tcgen05 epilogue, the DMMA compression, the fixed-lambda
// contraction, the solver scratch) can be built with explicit bounds checks that trap with a message instead.
//   PG_NVCC_EXTRA=-DPG_DEBUG_BOUNDS PG_LIB_OUT=pygemma_b200/libpygemma_b200_dbg.so python -m pygemma_b200.build --force
//   PYGEMMA_B200_LIB=pygemma_b200/libpygemma_b200_dbg.so python -m pytest tests -m gpu
#pragma once

#include <cstdio>

#if defined(PG_DEBUG_BOUNDS) && defined(__CUDA_ARCH__)
#define PG_BOUNDS(cond, what)                                                                                   \
    do {                                                                                                        \
        if (!(cond)) {                                                                                          \
            printf("PG_BOUNDS violated at %s:%d: %s (block %d thread %d)\n", __FILE__, __LINE__, what, (int)blockIdx.x, \
                   (int)threadIdx.x);                                                                           \
            __trap();                                                                                           \
        }                                                                                                       \
    } while (0)
#else
#define PG_BOUNDS(cond, what) ((void)0)
#endif
