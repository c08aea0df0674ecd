// specialfft.cu -- the CKKS encoder's "special" FFT on the device (last item):
//   ckks.SpecialFFTDouble / SpecialIFFTDouble       schemes/ckks/ckks_vector_ops.go:18-77
// called by Encoder.FFT / IFFT (schemes/ckks/encoder.go:764-816) on the slot vector (n = 2^LogSlots complex128 values, m = 2N roots).
// Floating point, but reproducible: the library is built with --fmad=false, every butterfly evaluates the reference's expressions in the
// reference's order (Go's complex multiply is (ac - bd) + (ad + bc)i without fusion on amd64, its complex division by complex(n, 0) divides
// both parts by n), and the caller passes the reference's own `roots` and `rotGroup` tables -- so results equal the Go code bit for bit
// (tests/test_gpu_specialfft.py compares with a scalar restatement, tolerance 0).
//
// Stages whose butterflies stay inside a 2048-value chunk (loglen <= 11) run in one kernel on a shared-memory tile (32 KB); the few wider
// stages of a large slot count run one launch each on the L2-resident vector (512 KB at 2^15 slots).
#include <cstring>
#include "capi_common.h"

namespace lgpu {

struct cplx { double re, im; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    cplx r;
    r.re = __dadd_rn(__dmul_rn(a.re, b.re), -__dmul_rn(a.im, b.im));     // ac - bd
    r.im = __dadd_rn(__dmul_rn(a.re, b.im), __dmul_rn(a.im, b.re));      // ad + bc
    return r;
}
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cplx{__dadd_rn(a.re, b.re), __dadd_rn(a.im, b.im)}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cplx{__dadd_rn(a.re, -b.re), __dadd_rn(a.im, -b.im)}; }

// one butterfly of stage `loglen` on the pair (k, k + lenh), j = k mod lenh
template <bool INV>
__device__ __forceinline__ void sfft_bfly(cplx& u, cplx& v, int j, int loglen, int logM, const long long* rot, const cplx* roots) {
    const int lenq = 4 << loglen, logGap = logM - 2 - loglen, mask = lenq - 1;
    const int r = (int)(rot[j] & mask);
    if (!INV) {
        v = cmul(v, roots[(size_t)r << logGap]);
        const cplx a = u;
        u = cadd(a, v); v = csub(a, v);
    } else {
        const cplx a = u, d = csub(u, v);
        u = cadd(a, v);
        v = cmul(d, roots[(size_t)(lenq - r) << logGap]);
    }
}

constexpr int kSfftChunkLog = 11;

// stages [lo, hi] (inclusive, all <= min(kSfftChunkLog, logN)) on one chunk of 2^min(logN, 11) values per CTA
template <bool INV>
__global__ void __launch_bounds__(256) sfft_chunk_kernel(cplx* values, size_t bs, int logN, int logM, int lo, int hi, const long long* rot, const cplx* roots) {
    extern __shared__ double2 sfft_sm[];
    cplx* sm = reinterpret_cast<cplx*>(sfft_sm);
    const int clog = logN < kSfftChunkLog ? logN : kSfftChunkLog;
    const int csize = 1 << clog;
    cplx* v = values + (size_t)blockIdx.y * bs + ((size_t)blockIdx.x << clog);
    for (int i = threadIdx.x; i < csize; i += blockDim.x) sm[i] = v[i];
    __syncthreads();
    for (int s = 0; s <= hi - lo; s++) {
        const int loglen = INV ? hi - s : lo + s;
        const int lenh = 1 << (loglen - 1);
        for (int t = threadIdx.x; t < csize / 2; t += blockDim.x) {
            const int j = t & (lenh - 1);
            const int k = ((t >> (loglen - 1)) << loglen) + j;
            sfft_bfly<INV>(sm[k], sm[k + lenh], j, loglen, logM, rot, roots);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < csize; i += blockDim.x) v[i] = sm[i];
}

// one wide stage (loglen > kSfftChunkLog) straight on global memory
template <bool INV>
__global__ void __launch_bounds__(256) sfft_stage_kernel(cplx* values, size_t bs, int n, int loglen, int logM, const long long* rot, const cplx* roots) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n / 2) return;
    cplx* v = values + (size_t)blockIdx.y * bs;
    const int lenh = 1 << (loglen - 1);
    const int j = t & (lenh - 1);
    const int k = ((t >> (loglen - 1)) << loglen) + j;
    cplx a = v[k], b = v[k + lenh];
    sfft_bfly<INV>(a, b, j, loglen, logM, rot, roots);
    v[k] = a; v[k + lenh] = b;
}

// utils.BitReverseInPlaceSlice; scale != 0: values[i] /= complex(n, 0) first (the inverse transform divides, THEN permutes)
__global__ void __launch_bounds__(256) sfft_bitrev_kernel(cplx* values, size_t bs, int n, int logN, double scale) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (unsigned)n) return;
    const unsigned j = logN ? (__brev(i) >> (32 - logN)) : 0;
    if (j < i) return;
    cplx* v = values + (size_t)blockIdx.y * bs;
    cplx a = v[i], b = v[j];
    if (scale != 0.0) { a.re = __ddiv_rn(a.re, scale); a.im = __ddiv_rn(a.im, scale); b.re = __ddiv_rn(b.re, scale); b.im = __ddiv_rn(b.im, scale); }
    v[i] = b; v[j] = a;
}

}  // namespace lgpu

using namespace lgpu;

extern "C" int lgpu_ckks_special_fft(lgpu_ctx* ctx, double* values, int n, int m, const int64_t* rot_group, const double* roots, int inverse, int batch,
                                     size_t batch_stride, void* stream) {
    REQUIRE_DEVICE(ctx);
    REQUIRE(values && rot_group && roots, "null argument");
    REQUIRE(n >= 1 && (n & (n - 1)) == 0 && m >= n && (m & (m - 1)) == 0 && m >= 4 * n, "invalid call of SpecialFFTDouble: n, m must be powers of two with m >= 4 n");
    REQUIRE(batch >= 1 && batch <= 65535, "batch out of range");
    REQUIRE((reinterpret_cast<uintptr_t>(values) & 15u) == 0 && (reinterpret_cast<uintptr_t>(roots) & 15u) == 0, "complex128 buffers must be 16-byte aligned");
    int logN = 0, logM = 0;
    while ((1 << logN) < n) logN++;
    while ((1 << logM) < m) logM++;
    if (logN == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    cplx* v = reinterpret_cast<cplx*>(values);
    const cplx* r = reinterpret_cast<const cplx*>(roots);
    const long long* rot = reinterpret_cast<const long long*>(rot_group);
    const int clog = logN < kSfftChunkLog ? logN : kSfftChunkLog;
    const dim3 cgrid((unsigned)(n >> clog), batch), sgrid((unsigned)((n / 2 + 255) / 256), batch), bgrid((unsigned)((n + 255) / 256), batch);
    const size_t smem = (size_t)sizeof(cplx) << clog;
    if (!inverse) {
        count_launch(2 + (logN - clog));
        sfft_bitrev_kernel<<<bgrid, 256, 0, st>>>(v, batch_stride, n, logN, 0.0);
        sfft_chunk_kernel<false><<<cgrid, 256, smem, st>>>(v, batch_stride, logN, logM, 1, clog, rot, r);
        for (int loglen = clog + 1; loglen <= logN; loglen++) sfft_stage_kernel<false><<<sgrid, 256, 0, st>>>(v, batch_stride, n, loglen, logM, rot, r);
    } else {
        count_launch(2 + (logN - clog));
        for (int loglen = logN; loglen > clog; loglen--) sfft_stage_kernel<true><<<sgrid, 256, 0, st>>>(v, batch_stride, n, loglen, logM, rot, r);
        sfft_chunk_kernel<true><<<cgrid, 256, smem, st>>>(v, batch_stride, logN, logM, 1, clog, rot, r);
        sfft_bitrev_kernel<<<bgrid, 256, 0, st>>>(v, batch_stride, n, logN, (double)n);
    }
    LGPU_CUDA_OK(cudaGetLastError());
    return 0;
}
