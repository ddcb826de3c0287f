// oracle/ref/fftw_shim.cpp — TEST INFRASTRUCTURE (not product code).
//
// Defines the five FFTW entry points the unmodified reference links against
// (rtinc/buffer.hpp:19-20, rtlib/fft.hpp:13-18).  FFTW itself (an external, version-unpinned
// dependency of the reference: `-lfftw3`, compile:13) is not installed in this image, so the
// contract implemented here is FFTW's documented definition of the forward transform,
//     Y[k] = sum_j X[j] * exp(-2*pi*i*j*k/n),   unnormalised,
// in double precision.  Two interchangeable providers:
//   default        a from-scratch double FFT: power-of-two n -> iterative radix-2 with twiddles
//                  generated in long double; any other n -> direct O(n^2) DFT accumulated in
//                  long double (slow, exact to double rounding)
//   -DTSREF_USE_MKL  Intel MKL DFTI (double) as exported by torch's libtorch_cpu.so — used for the
//                  CPU baseline timing and to measure the oracle's own FFT-backend noise floor.
#include <fftw3.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

struct tsref_fftw_plan_s {
    int n, howmany;
    std::complex<double> *in, *out;
    int idist, odist;
    int sign;
    void *mkl;  // DFTI descriptor when built with TSREF_USE_MKL
};

extern "C" void *fftw_malloc(size_t n) {
    void *p = nullptr;
    if (posix_memalign(&p, 64, n ? n : 1)) return nullptr;
    return p;
}
extern "C" void fftw_free(void *p) { free(p); }

#ifdef TSREF_USE_MKL
// Minimal DFTI surface (values from mkl_dfti.h: the public, stable enum of Intel MKL).
extern "C" {
typedef struct DFTI_DESCRIPTOR *DFTI_DESCRIPTOR_HANDLE;
long DftiCreateDescriptor_d_1d(DFTI_DESCRIPTOR_HANDLE *, int domain, long n);
long DftiSetValue(DFTI_DESCRIPTOR_HANDLE, int param, ...);
long DftiCommitDescriptor(DFTI_DESCRIPTOR_HANDLE);
long DftiComputeForward(DFTI_DESCRIPTOR_HANDLE, void *, ...);
long DftiComputeBackward(DFTI_DESCRIPTOR_HANDLE, void *, ...);
long DftiFreeDescriptor(DFTI_DESCRIPTOR_HANDLE *);
}
enum { DFTI_NUMBER_OF_TRANSFORMS = 7, DFTI_PLACEMENT = 11, DFTI_INPUT_DISTANCE = 14,
       DFTI_OUTPUT_DISTANCE = 15, DFTI_COMPLEX = 32, DFTI_INPLACE = 43, DFTI_NOT_INPLACE = 44 };
#else
namespace {
struct Tw {
    std::vector<std::complex<double>> w;  // exp(-2*pi*i*t/n), t < n
};
std::map<int, Tw> g_tw;
std::mutex g_tw_mtx;
const Tw &twiddles(int n) {
    std::lock_guard<std::mutex> lk(g_tw_mtx);
    auto it = g_tw.find(n);
    if (it != g_tw.end()) return it->second;
    Tw t;
    t.w.resize(n);
    const long double tau = 6.283185307179586476925286766559005768L;
    for (int k = 0; k < n; k++) {
        long double a = -tau * (long double)k / (long double)n;
        t.w[k] = std::complex<double>((double)cosl(a), (double)sinl(a));
    }
    return g_tw.emplace(n, std::move(t)).first->second;
}
void fft_pow2(const std::complex<double> *in, std::complex<double> *out, int n, int sign) {
    const Tw &tw = twiddles(n);
    int lg = 0;
    while ((1 << lg) < n) lg++;
    std::vector<std::complex<double>> a(n);
    for (int i = 0; i < n; i++) {
        unsigned r = 0;
        for (int b = 0; b < lg; b++) r |= ((i >> b) & 1u) << (lg - 1 - b);
        a[r] = in[i];
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int s = 0; s < n; s += len)
            for (int j = 0; j < half; j++) {
                std::complex<double> w = tw.w[j * step];
                if (sign > 0) w = std::conj(w);
                std::complex<double> u = a[s + j], v = a[s + j + half] * w;
                a[s + j] = u + v;
                a[s + j + half] = u - v;
            }
    }
    memcpy(out, a.data(), sizeof(std::complex<double>) * n);
}
void dft_any(const std::complex<double> *in, std::complex<double> *out, int n, int sign) {
    const long double tau = 6.283185307179586476925286766559005768L;
    std::vector<long double> c(n), s(n);
    for (int k = 0; k < n; k++) {
        long double a = tau * (long double)k / (long double)n;
        c[k] = cosl(a);
        s[k] = (sign < 0 ? -1.0L : 1.0L) * sinl(a);
    }
    std::vector<std::complex<double>> res(n);
    for (int k = 0; k < n; k++) {
        long double re = 0, im = 0;
        for (int j = 0; j < n; j++) {
            int t = (int)(((long long)j * k) % n);
            long double xr = in[j].real(), xi = in[j].imag();
            re += xr * c[t] - xi * s[t];
            im += xr * s[t] + xi * c[t];
        }
        res[k] = std::complex<double>((double)re, (double)im);
    }
    memcpy(out, res.data(), sizeof(std::complex<double>) * n);
}
}  // namespace
#endif

extern "C" fftw_plan fftw_plan_many_dft(int rank, const int *n, int howmany, fftw_complex *in, const int *,
                                        int istride, int idist, fftw_complex *out, const int *, int ostride,
                                        int odist, int sign, unsigned) {
    if (rank != 1 || istride != 1 || ostride != 1) {
        fprintf(stderr, "fftw_shim: only rank-1 unit-stride plans are supported\n");
        abort();
    }
    fftw_plan p = new tsref_fftw_plan_s{n[0], howmany, (std::complex<double> *)in, (std::complex<double> *)out,
                                        idist, odist, sign, nullptr};
#ifdef TSREF_USE_MKL
    DFTI_DESCRIPTOR_HANDLE h = nullptr;
    long st = DftiCreateDescriptor_d_1d(&h, DFTI_COMPLEX, (long)n[0]);
    st |= DftiSetValue(h, DFTI_NUMBER_OF_TRANSFORMS, (long)howmany);
    st |= DftiSetValue(h, DFTI_INPUT_DISTANCE, (long)idist);
    st |= DftiSetValue(h, DFTI_OUTPUT_DISTANCE, (long)odist);
    st |= DftiSetValue(h, DFTI_PLACEMENT, in == out ? DFTI_INPLACE : DFTI_NOT_INPLACE);
    st |= DftiCommitDescriptor(h);
    if (st) { fprintf(stderr, "fftw_shim: MKL DFTI setup failed (%ld)\n", st); abort(); }
    p->mkl = h;
#endif
    return p;
}

extern "C" void fftw_execute(const fftw_plan p) {
#ifdef TSREF_USE_MKL
    DFTI_DESCRIPTOR_HANDLE h = (DFTI_DESCRIPTOR_HANDLE)p->mkl;
    long st;
    if (p->sign < 0) st = p->in == p->out ? DftiComputeForward(h, p->in) : DftiComputeForward(h, p->in, p->out);
    else st = p->in == p->out ? DftiComputeBackward(h, p->in) : DftiComputeBackward(h, p->in, p->out);
    if (st) { fprintf(stderr, "fftw_shim: MKL DFTI compute failed (%ld)\n", st); abort(); }
#else
    const int n = p->n;
    const bool pow2 = n > 0 && (n & (n - 1)) == 0;
    for (int h = 0; h < p->howmany; h++) {
        const std::complex<double> *src = p->in + (size_t)h * p->idist;
        std::complex<double> *dst = p->out + (size_t)h * p->odist;
        if (pow2) fft_pow2(src, dst, n, p->sign);
        else dft_any(src, dst, n, p->sign);
    }
#endif
}

extern "C" void fftw_destroy_plan(fftw_plan p) {
#ifdef TSREF_USE_MKL
    DFTI_DESCRIPTOR_HANDLE h = (DFTI_DESCRIPTOR_HANDLE)p->mkl;
    if (h) DftiFreeDescriptor(&h);
#endif
    delete p;
}
