// tsb_launch.hpp — launch entry points of the per-size chain kernels (one translation unit per
// frame size, tsb_chain_inst.cu compiled with -DTSB_LOG2N=k).
#pragma once
#include <cuda_runtime.h>

#include "tsb_chain.cuh"

namespace tsb {
// tables: ChainTables<float> or ChainTables<double> matching `precision`
typedef cudaError_t (*chain_launch_fn)(int precision, int program, const void *tables, const ChainArgs &a,
                                       cudaStream_t stream);
#define TSB_DECL_LAUNCH(L) \
    cudaError_t launch_chain_##L(int precision, int program, const void *tables, const ChainArgs &a, cudaStream_t stream);
TSB_DECL_LAUNCH(6)
TSB_DECL_LAUNCH(7)
TSB_DECL_LAUNCH(8)
TSB_DECL_LAUNCH(9)
TSB_DECL_LAUNCH(10)
TSB_DECL_LAUNCH(11)
TSB_DECL_LAUNCH(12)
TSB_DECL_LAUNCH(13)
#undef TSB_DECL_LAUNCH
}  // namespace tsb
