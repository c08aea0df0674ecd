// ringmisc.cu -- the remaining coefficient-moving Ring methods of ring/operations.go and ring/ringqp/operations.go:
// Shift, MultByMonomial, MapSmallDimensionToLargerDimensionNTT, ExtendBasisSmallNormAndCenter.
// Pure data movement (+ a negation): one thread per output coefficient, coalesced writes.
#include <cstring>
#include "capi_common.h"
#include "modarith.cuh"

using namespace lgpu;


namespace lgpu {

struct MoveParams {
    const LimbConst* limbs;
    RowMap rm;
    const u64* in;
    u64* out;
    size_t bs_in, bs_out;
    int N;
    int k;      // Shift: left rotation (0 <= k < N); MultByMonomial: shift in [0, 2N)
};

// Ring.Shift, ring/operations.go:278-282 (utils.RotateSliceAllocFree: out[i] = in[(i + k) mod N])
__global__ void __launch_bounds__(256) shift_kernel(MoveParams p) {
    const int row = p.rm.drow[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    const u64* in = p.in + (size_t)blockIdx.z * p.bs_in + (size_t)row * p.N;
    u64* out = p.out + (size_t)blockIdx.z * p.bs_out + (size_t)row * p.N;
    int s = i + p.k;
    if (s >= p.N) s -= p.N;
    out[i] = in[s];
}

// Ring.MultByMonomial, ring/operations.go:306-363: p2 = p1 * X^k, with the reference's literal negation (q - x, so
// a zero coefficient that wraps becomes q, exactly as the reference produces it).
__global__ void __launch_bounds__(256) monomial_kernel(MoveParams p) {
    const LimbConst L = p.limbs[p.rm.limb[blockIdx.y]];
    const int row = p.rm.drow[blockIdx.y];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.N) return;
    const u64* in = p.in + (size_t)blockIdx.z * p.bs_in + (size_t)row * p.N;
    u64* out = p.out + (size_t)blockIdx.z * p.bs_out + (size_t)row * p.N;
    const u64 q = L.q;
    if (p.k == 0) { out[j] = in[j]; return; }
    const bool neg = p.k >= p.N;
    const int shift = neg ? p.k - p.N : p.k;
    u64 v;
    if (j < shift) {
        v = in[p.N - shift + j];
        if (neg) v = q - v;
        v = q - v;
    } else {
        v = in[j - shift];
        if (neg) v = q - v;
    }
    out[j] = v;
}

// MapSmallDimensionToLargerDimensionNTT, ring/operations.go:380-392
__global__ void __launch_bounds__(256) map_small_to_large_kernel(const u64* small_, u64* large, int n_small, int lg_gap, int rows) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_large = (size_t)n_small << lg_gap;
    if (i >= n_large * rows) return;
    const size_t row = i / n_large, c = i % n_large;
    large[i] = small_[row * n_small + (c >> lg_gap)];
}

// ringqp.Ring.ExtendBasisSmallNormAndCenter, ring/ringqp/operations.go:325-351
__global__ void __launch_bounds__(256) extend_small_norm_kernel(const u64* inq0, u64* outp, const LimbConst* limbs, int nQ, int np, int N,
                                                                 size_t bs_in, size_t bs_out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const u64 Q = limbs[0].q, half = Q >> 1;
    u64 coeff = inq0[(size_t)blockIdx.z * bs_in + j];
    const bool negative = coeff > half;
    if (negative) coeff = Q - coeff;
    for (int i = 0; i < np; i++) {
        const u64 pi = limbs[nQ + i].q;
        outp[(size_t)blockIdx.z * bs_out + (size_t)i * N + j] = negative ? pi - coeff : coeff;
    }
}
}  // namespace lgpu


// in-place calls go through a stream-ordered temporary copy of the input (the reference allows p1 == p2 for both)
static int move_common(lgpu_ctx* ctx, int ring, int level, const uint64_t* in, uint64_t* out, int k, int batch, size_t batch_stride, bool monomial,
                       cudaStream_t st) {
    REQUIRE(in && out, "null polynomial");
    REQUIRE(batch >= 1, "batch must be >= 1");
    const Ctx& c = ctx->c;
    MoveParams p;
    if (make_rowmap(c, ring, level, p.rm)) return -1;
    p.limbs = c.d_limbs; p.N = c.N; p.k = k; p.out = (u64*)out; p.bs_out = batch_stride;
    u64* tmp = nullptr;
    const size_t rows_words = (size_t)(level + 1) * c.N;
    if ((const void*)in == (const void*)out) {
        LGPU_CUDA_OK(cudaMallocAsync((void**)&tmp, (size_t)batch * rows_words * sizeof(u64), st));
        LGPU_CUDA_OK(cudaMemcpy2DAsync(tmp, rows_words * sizeof(u64), in, (batch > 1 ? batch_stride : rows_words) * sizeof(u64), rows_words * sizeof(u64),
                                       batch, cudaMemcpyDeviceToDevice, st));
        p.in = tmp; p.bs_in = rows_words;
    } else {
        p.in = (const u64*)in; p.bs_in = batch_stride;
    }
    const dim3 grid((unsigned)((c.N + 255) / 256), p.rm.nrows, batch);
    if (monomial) monomial_kernel<<<grid, 256, 0, st>>>(p);
    else shift_kernel<<<grid, 256, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (tmp) cudaFreeAsync(tmp, st);
    if (e != cudaSuccess) { set_error(std::string("move kernel: ") + cudaGetErrorString(e)); return -1; }
    return 0;
}

// bootstrapping.Evaluator.ModUp's coefficient loops (circuits/ckks/bootstrapping/evaluator.go:652-699, :729-741): row 0 (mod q_0) is lifted,
// centred around q_0, to the Q rows [first_q, level_q] and the P rows [0, level_p] of the output, literally:
//   coeff >= q/2 (strict = 0) or coeff > q/2 (strict = 1)  ->  coeff = q - coeff, negative;   out = BRedAdd(coeff) or Q_i - BRedAdd(coeff)
// (a negative multiple of Q_i therefore comes out as Q_i, not 0, like in the reference).
__global__ void __launch_bounds__(256) modup_centered_kernel(const u64* in, u64* out_q, u64* out_p, const lgpu::LimbConst* limbs, int nQ, int first_q,
                                                             int nq, int np, int n, int strict, size_t in_bs, size_t q_bs, size_t p_bs) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int b = blockIdx.z;
    const u64 q = limbs[0].q;
    u64 coeff = in[(size_t)b * in_bs + j];
    const bool neg = strict ? (coeff > (q >> 1)) : (coeff >= (q >> 1));
    if (neg) coeff = q - coeff;
    for (int i = first_q; i < nq; i++) {
        const lgpu::LimbConst& L = limbs[i];
        const u64 t = lgpu::bred_add(coeff, L.q, L.bred_hi);
        out_q[(size_t)b * q_bs + (size_t)i * n + j] = neg ? L.q - t : t;
    }
    for (int i = 0; i < np; i++) {
        const lgpu::LimbConst& L = limbs[nQ + i];
        const u64 t = lgpu::bred_add(coeff, L.q, L.bred_hi);
        out_p[(size_t)b * p_bs + (size_t)i * n + j] = neg ? L.q - t : t;
    }
}


extern "C" {

int lgpu_shift(lgpu_ctx* ctx, int ring, int level, const uint64_t* in, int k, uint64_t* out, int batch, size_t batch_stride, void* stream) {
    REQUIRE_DEVICE(ctx);
    const int N = ctx->c.N;
    int kk = k % N;
    if (kk < 0) kk += N;
    return move_common(ctx, ring, level, in, out, kk, batch, batch_stride, false, (cudaStream_t)stream);
}

int lgpu_mult_by_monomial(lgpu_ctx* ctx, int ring, int level, const uint64_t* in, int k, uint64_t* out, int batch, size_t batch_stride, void* stream) {
    REQUIRE_DEVICE(ctx);
    const long long twoN = 2ll * ctx->c.N;
    REQUIRE((long long)k > -twoN, "MultByMonomial: k must be > -2N (the reference computes (k + 2N) % 2N)");
    const int shift = (int)(((long long)k + twoN) % twoN);
    return move_common(ctx, ring, level, in, out, shift, batch, batch_stride, true, (cudaStream_t)stream);
}

int lgpu_map_small_dimension_to_larger_dimension_ntt(lgpu_ctx* ctx, const uint64_t* pol_small, int n_small, uint64_t* pol_large, int n_large, int rows,
                                                     void* stream) {
    REQUIRE_DEVICE(ctx);
    REQUIRE(pol_small && pol_large, "null polynomial");
    REQUIRE(n_small > 0 && n_large >= n_small && (n_large % n_small) == 0 && ((n_large / n_small) & (n_large / n_small - 1)) == 0 && rows > 0,
            "MapSmallDimensionToLargerDimensionNTT: degrees must be powers of two with n_small | n_large");
    int lg = 0;
    while ((n_small << lg) < n_large) lg++;
    const size_t total = (size_t)n_large * rows;
    map_small_to_large_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const u64*)pol_small, (u64*)pol_large, n_small, lg, rows);
    LGPU_CUDA_OK(cudaGetLastError());
    return 0;
}

int lgpu_extend_basis_small_norm_and_center(lgpu_ctx* ctx, const uint64_t* poly_in_q, int level_p, uint64_t* poly_out_p, int batch, size_t stride_q,
                                            size_t stride_p, void* stream) {
    REQUIRE_DEVICE(ctx);
    REQUIRE(poly_in_q && poly_out_p, "null polynomial");
    REQUIRE(level_p >= 0 && level_p < ctx->c.nP, "levelP out of range");
    REQUIRE(batch >= 1, "batch must be >= 1");
    const Ctx& c = ctx->c;
    extend_small_norm_kernel<<<dim3((unsigned)((c.N + 255) / 256), 1, batch), 256, 0, (cudaStream_t)stream>>>(
        (const u64*)poly_in_q, (u64*)poly_out_p, c.d_limbs, c.nQ, level_p + 1, c.N, stride_q, stride_p);
    LGPU_CUDA_OK(cudaGetLastError());
    return 0;
}

int lgpu_modup_centered(lgpu_ctx* ctx, const uint64_t* row0, int first_q, int level_q, int level_p, int strict, uint64_t* out_q, uint64_t* out_p,
                        int batch, size_t stride_in, size_t stride_q, size_t stride_p, void* stream) {
    REQUIRE_DEVICE(ctx);
    REQUIRE(row0 && (out_q || level_q < first_q) && (out_p || level_p < 0), "null polynomial");
    const Ctx& c = ctx->c;
    REQUIRE(first_q >= 0 && level_q < c.nQ && level_p < c.nP, "level out of range");
    REQUIRE(batch >= 1 && batch <= 65535, "batch out of range");
    modup_centered_kernel<<<dim3((unsigned)((c.N + 255) / 256), 1, batch), 256, 0, (cudaStream_t)stream>>>(
        (const u64*)row0, (u64*)out_q, (u64*)out_p, c.d_limbs, c.nQ, first_q, level_q + 1, level_p + 1, c.N, strict, stride_in, stride_q, stride_p);
    LGPU_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
