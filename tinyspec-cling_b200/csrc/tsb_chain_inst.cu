// tsb_chain_inst.cu — the fused chain kernel for one frame size N = 2^TSB_LOG2N (sm_100a).
// One CTA per lane; the CTA walks the lane's frames in order (see tsb_chain.cuh for the phases).
#include "tsb_launch.hpp"

#ifndef TSB_LOG2N
#error "compile with -DTSB_LOG2N=<6..13>"
#endif

namespace tsb {

template <class CH>
struct DevExec {
    typename CH::Regs rg;
    template <class F>
    __device__ __forceinline__ void each(F f) {
        f((int)threadIdx.x, rg);
        __syncthreads();
    }
    template <class F>
    __device__ __forceinline__ void each_nosync(F f) {
        f((int)threadIdx.x, rg);
    }
};

template <typename real, int LOG2N, int PROG>
__global__ void __launch_bounds__(Chain<real, LOG2N, PROG>::T)
chain_kernel(const ChainTables<real> tb, const ChainArgs a) {
    using CH = Chain<real, LOG2N, PROG>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename CH::Smem &sm = *reinterpret_cast<typename CH::Smem *>(smem_raw);
    DevExec<CH> ex;
    for (int lane = blockIdx.x; lane < a.lanes; lane += gridDim.x) run_lane<CH>(ex, lane, sm, tb, a);
}

template <typename real, int LOG2N, int PROG>
static cudaError_t launch(const void *tables, const ChainArgs &a, cudaStream_t stream) {
    using CH = Chain<real, LOG2N, PROG>;
    auto kern = chain_kernel<real, LOG2N, PROG>;
    size_t smem = sizeof(typename CH::Smem);
    static bool configured = false;  // per instantiation; attribute is per device function
    if (!configured || true) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (a.lanes <= 0 || a.n_frames <= 0) return cudaSuccess;
    kern<<<a.lanes, CH::T, smem, stream>>>(*reinterpret_cast<const ChainTables<real> *>(tables), a);
    return cudaGetLastError();
}

#define TSB_CAT_(a, b) a##b
#define TSB_CAT(a, b) TSB_CAT_(a, b)
cudaError_t TSB_CAT(launch_chain_, TSB_LOG2N)(int precision, int program, const void *tables, const ChainArgs &a,
                                              cudaStream_t stream) {
    if (precision == 0) {
        return program == PROG_PV_SHIFT ? launch<float, TSB_LOG2N, PROG_PV_SHIFT>(tables, a, stream)
                                        : launch<float, TSB_LOG2N, PROG_GAIN>(tables, a, stream);
    }
    return program == PROG_PV_SHIFT ? launch<double, TSB_LOG2N, PROG_PV_SHIFT>(tables, a, stream)
                                    : launch<double, TSB_LOG2N, PROG_GAIN>(tables, a, stream);
}

}  // namespace tsb
