// tsb_chain_inst.cu — the fused chain kernel for one frame size N = 2^TSB_LOG2N (sm_100a).
// One CTA per lane; the CTA walks the lane's frames in order (see tsb_chain.cuh for the phases).
#include "tsb_launch.hpp"

#include <cstdlib>

#ifndef TSB_LOG2N
#error "compile with -DTSB_LOG2N=<6..13>"
#endif
#ifndef TSB_REGS_PV32
#define TSB_REGS_PV32 64
#endif
#ifndef TSB_REGS_PV64
#define TSB_REGS_PV64 96
#endif
#ifndef TSB_REGS_PV32_RT
// fp32 with the overlap tail in registers (N=1024): 96 registers, five lanes per SM (80 registers / six lanes spill: measured
// 227 M frame-channels/s against 231 M; without the tail 64 registers / eight lanes: 223 M)
#define TSB_REGS_PV32_RT 96
#endif
#ifndef TSB_REGS_PV64_RT
// the kernel that keeps its overlap tail in registers (Chain::RT_OK, N=1024): four lanes per SM, 128 registers
#define TSB_REGS_PV64_RT 128
#endif
#ifndef TSB_REGS_PV64_INPLACE
// pipelined fp64 kernels with the in-place inverse transform (Chain::INV_INPLACE) have shared memory for one more lane
// per SM than 96 registers allow; at 80 registers they spill a little.  Measured: N=2048 +6 % with the third lane,
// N=512 +18 %; at N=1024 six lanes at 80 registers equal five at 96 (137 M frame-channels/s either way: the kernel
// is bound by the L1/shared-memory data pipe, not by latency — even four lanes per SM reach 137 M), so that size keeps
// 96 registers and gives the shared memory it no longer needs to the L1 cache instead (see launch(): +4 %).
#define TSB_REGS_PV64_INPLACE 80
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
    // f, then a barrier over this thread's half of the CTA only (threads [0, half) and [half, 2*half): named barriers 1, 2)
    template <class F>
    __device__ __forceinline__ void each_half(F f, int half) {
        f((int)threadIdx.x, rg);
        if ((int)threadIdx.x >= half) asm volatile("bar.sync 2, %0;" ::"r"(half) : "memory");
        else asm volatile("bar.sync 1, %0;" ::"r"(half) : "memory");
    }
    // same, and tells every thread whether f returned true for any of them
    template <class F>
    __device__ __forceinline__ bool each_or(F f) {
        return __syncthreads_or(f((int)threadIdx.x, rg) ? 1 : 0) != 0;
    }
};

// Resident CTAs per SM the kernel is compiled for: what shared memory allows (227 KB, 1 KB reserved per
// CTA), capped by the thread limit and by a register budget that still keeps the FFT butterflies and the
// PV state in registers.
template <class CH>
constexpr int min_blocks() {
    constexpr int by_smem = (227 * 1024) / ((int)sizeof(typename CH::Smem) + 1024);
    constexpr int by_threads = 2048 / CH::T;
    constexpr int by_regs = 65536 / (CH::T * (sizeof(typename CH::C) == 8 ? (CH::IS_PV ? (CH::RT_OK ? TSB_REGS_PV32_RT : TSB_REGS_PV32) : 80) : (CH::IS_PV ? (CH::RT_OK ? TSB_REGS_PV64_RT : ((CH::INV_INPLACE && CH::N != 1024) ? TSB_REGS_PV64_INPLACE : TSB_REGS_PV64)) : (CH::RT_OK ? TSB_REGS_PV64_RT : 96))));
    int b = by_smem < by_threads ? by_smem : by_threads;
    b = b < by_regs ? b : by_regs;
    return b < 1 ? 1 : (b > 32 ? 32 : b);
}


// PREPASS: the time-parallel pre-pass (ChainArgs::mode == 1: analysis and per-frame increments only)
// The fp32 pre-pass carries the exact-wrap phase (Chain::phase_wrap_fix) in its plain frame loop; one CTA less per
// SM gives it the registers to do so without spilling (the pre-pass is latency-bound on its barriers either way).
template <class CH, bool PREPASS>
constexpr int kernel_blocks() {
    return (PREPASS && CH::WRAP_FIX && min_blocks<CH>() > 2) ? min_blocks<CH>() - 1 : min_blocks<CH>();
}
template <typename real, int LOG2N, int PROG, bool OUT32, bool PREPASS>
__global__ void __launch_bounds__(Chain<real, LOG2N, PROG, OUT32>::T, kernel_blocks<Chain<real, LOG2N, PROG, OUT32>, PREPASS>())
chain_kernel(const ChainTables<real> tb, const ChainArgs a) {
    using CH = Chain<real, LOG2N, PROG, OUT32>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename CH::Smem &sm = *reinterpret_cast<typename CH::Smem *>(smem_raw);
    DevExec<CH> ex;
    const int nseg = a.n_seg > 1 ? a.n_seg : 1;
    for (long long b = blockIdx.x; b < (long long)a.lanes * nseg; b += gridDim.x)
        run_lane<CH, PREPASS>(ex, (int)(b % a.lanes), (int)(b / a.lanes), sm, tb, a);
}

template <typename real, int LOG2N, int PROG, bool OUT32 = false>
static cudaError_t launch(const void *tables, const ChainArgs &a, cudaStream_t stream) {
    using CH = Chain<real, LOG2N, PROG, OUT32>;
    // the pre-pass writes no samples: one instantiation (OUT32 = false) serves both output types
    auto kern = (a.mode == 1 && PROG == PROG_PV_SHIFT) ? chain_kernel<real, LOG2N, PROG, OUT32, true>
                                                         : chain_kernel<real, LOG2N, PROG, OUT32, false>;
    size_t smem = sizeof(typename CH::Smem);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // Shared memory and L1 share 228 KB per SM, and the constant tables (twiddles, windows, per-bin phase constants,
    // shift map: ~20 KB fp32 / ~34 KB fp64 at N=1024) are read through L1 every frame (TSB_CARVEOUT_KB overrides the
    // carve-out, for tuning).  The fp64 kernels with the in-place inverse transform ask for exactly what their resident lanes use (N=1024: 5 x
    // 37 KB -> 43 KB of L1 instead of ~3: 139 -> 143 M frame-channels/s); the others keep the maximum (fp32 at N=1024
    // measured 1.5-3 % slower with any smaller carve-out).
    int carve_kb = 228;
    if (CH::INV_INPLACE) {
        constexpr int blocks = kernel_blocks<CH, false>();
        carve_kb = (int)((blocks * (sizeof(typename CH::Smem) + 1024) + 1023) / 1024);
    }
    if (const char *ev = getenv("TSB_CARVEOUT_KB")) carve_kb = atoi(ev);
    int pct = carve_kb >= 228 ? 100 : (carve_kb * 100 + 227) / 228;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e != cudaSuccess) return e;
    if (a.lanes <= 0 || a.n_frames <= 0) return cudaSuccess;
    const long long ctas = (long long)a.lanes * (a.n_seg > 1 ? a.n_seg : 1);
    kern<<<(unsigned)ctas, CH::T, smem, stream>>>(*reinterpret_cast<const ChainTables<real> *>(tables), a);
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
    if (precision == 2) {  // fp64 arithmetic, float32 output stream (TSB_F64_OUT32)
        return program == PROG_PV_SHIFT ? launch<double, TSB_LOG2N, PROG_PV_SHIFT, true>(tables, a, stream)
                                        : launch<double, TSB_LOG2N, PROG_GAIN, true>(tables, a, stream);
    }
    return program == PROG_PV_SHIFT ? launch<double, TSB_LOG2N, PROG_PV_SHIFT>(tables, a, stream)
                                    : launch<double, TSB_LOG2N, PROG_GAIN>(tables, a, stream);
}

}  // namespace tsb
