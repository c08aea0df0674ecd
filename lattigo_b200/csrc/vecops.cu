// vecops.cu -- the coefficient-wise kernel family of ring/vec_ops.go:7-890, driven like
// ring/operations.go:11-392 (loop over limbs) and ring/subring_ops.go:6-273 (single rows).
// One templated streaming kernel, opcode-specialised at compile time; 128-bit vector loads/stores
// over the (limb, coeff) row-major layout. Lazy variants return the reference's exact representative.
#include "engine.h"
#include "modarith.cuh"
#include "../../include/lattigo_b200.h"

namespace lgpu {

struct VecParams {
    const LimbConst* limbs;
    RowMap rm;
    const u64* p1; const u64* p2; u64* p3;
    size_t rs1, bs1, rs2, bs2, rs3, bs3;
    int n;
    u64 s0[kMaxRows];
    u64 s1[kMaxRows];
};

template <int OP>
__device__ __forceinline__ u64 apply(u64 x, u64 y, u64 z, const LimbConst& L, u64 s0, u64 s1) {
    const u64 q = L.q, qinv = L.qinv, bh = L.bred_hi, bl = L.bred_lo;
    const u64 twoq = q << 1;
    switch (OP) {
        case LGPU_OP_ADD: return cred(x + y, q);
        case LGPU_OP_ADDLAZY: return x + y;
        case LGPU_OP_SUB: return cred((x + q) - y, q);
        case LGPU_OP_SUBLAZY: return x + q - y;
        case LGPU_OP_NEG: return q - x;
        case LGPU_OP_REDUCE: return bred_add(x, q, bh);
        case LGPU_OP_REDUCELAZY: return bred_add_lazy(x, q, bh);
        case LGPU_OP_MULCOEFFSLAZY: return x * y;
        case LGPU_OP_MULCOEFFSLAZYTHENADDLAZY: return z + x * y;
        case LGPU_OP_MULCOEFFSBARRETT: return bred(x, y, q, bh, bl);
        case LGPU_OP_MULCOEFFSBARRETTLAZY: return bred_lazy(x, y, q, bh, bl);
        case LGPU_OP_MULCOEFFSBARRETTTHENADD: return cred(z + bred(x, y, q, bh, bl), q);
        case LGPU_OP_MULCOEFFSBARRETTTHENADDLAZY: return z + bred(x, y, q, bh, bl);
        case LGPU_OP_MULCOEFFSMONTGOMERY: return mred(x, y, q, qinv);
        case LGPU_OP_MULCOEFFSMONTGOMERYLAZY: return mred_lazy(x, y, q, qinv);
        case LGPU_OP_MULCOEFFSMONTGOMERYTHENADD: return cred(z + mred(x, y, q, qinv), q);
        case LGPU_OP_MULCOEFFSMONTGOMERYTHENADDLAZY: return z + mred(x, y, q, qinv);
        case LGPU_OP_MULCOEFFSMONTGOMERYLAZYTHENADDLAZY: return z + mred_lazy(x, y, q, qinv);
        case LGPU_OP_MULCOEFFSMONTGOMERYTHENSUB: return cred(z + (q - mred(x, y, q, qinv)), q);
        case LGPU_OP_MULCOEFFSMONTGOMERYTHENSUBLAZY: return z + (q - mred(x, y, q, qinv));
        case LGPU_OP_MULCOEFFSMONTGOMERYLAZYTHENSUBLAZY: return z + twoq - mred_lazy(x, y, q, qinv);
        case LGPU_OP_MULCOEFFSMONTGOMERYLAZYTHENNEG: return twoq - mred_lazy(x, y, q, qinv);
        case LGPU_OP_ADDLAZYTHENMULSCALARMONTGOMERY: return mred(x + y, s0, q, qinv);
        case LGPU_OP_ADDSCALARLAZYTHENMULSCALARMONTGOMERY: return mred(x + s0, s1, q, qinv);
        case LGPU_OP_ADDSCALAR: return cred(x + s0, q);
        case LGPU_OP_ADDSCALARLAZY: return x + s0;
        case LGPU_OP_ADDSCALARLAZYTHENNEGTWOMODULUSLAZY: return s0 + twoq - x;
        case LGPU_OP_SUBSCALAR: return cred(x + q - s0, q);
        case LGPU_OP_MULSCALARMONTGOMERY: return mred(x, s0, q, qinv);
        case LGPU_OP_MULSCALARMONTGOMERYLAZY: return mred_lazy(x, s0, q, qinv);
        case LGPU_OP_MULSCALARMONTGOMERYTHENADD: return cred(z + mred(x, s0, q, qinv), q);
        case LGPU_OP_MULSCALARMONTGOMERYTHENADDSCALAR: return cred(mred(x, s1, q, qinv) + s0, q);
        case LGPU_OP_SUBTHENMULSCALARMONTGOMERYTWOMODULUS: return mred(twoq - y + x, s0, q, qinv);
        case LGPU_OP_MFORM: return mform(x, q, bh, bl);
        case LGPU_OP_MFORMLAZY: return mform_lazy(x, q, bh, bl);
        case LGPU_OP_IMFORM: return imform(x, q, qinv);
        case LGPU_OP_ZERO: return 0;
        case LGPU_OP_MASK: return (x >> s0) & s1;
    }
    return 0;
}

__host__ __device__ constexpr bool op_uses_p1(int op) { return op != LGPU_OP_ZERO; }
__host__ __device__ constexpr bool op_uses_p2(int op) {
    return op == LGPU_OP_ADD || op == LGPU_OP_ADDLAZY || op == LGPU_OP_SUB || op == LGPU_OP_SUBLAZY ||
           (op >= LGPU_OP_MULCOEFFSLAZY && op <= LGPU_OP_ADDLAZYTHENMULSCALARMONTGOMERY) ||
           op == LGPU_OP_SUBTHENMULSCALARMONTGOMERYTWOMODULUS;
}
__host__ __device__ constexpr bool op_reads_p3(int op) {
    return op == LGPU_OP_MULCOEFFSLAZYTHENADDLAZY || op == LGPU_OP_MULCOEFFSBARRETTTHENADD ||
           op == LGPU_OP_MULCOEFFSBARRETTTHENADDLAZY ||
           (op >= LGPU_OP_MULCOEFFSMONTGOMERYTHENADD && op <= LGPU_OP_MULCOEFFSMONTGOMERYLAZYTHENSUBLAZY) ||
           op == LGPU_OP_MULSCALARMONTGOMERYTHENADD;
}

template <int OP, int VEC>
__global__ void __launch_bounds__(256) vecop_kernel(VecParams p) {
    const int row = blockIdx.y, b = blockIdx.z;
    const LimbConst L = p.limbs[p.rm.limb[row]];
    const u64 s0 = p.s0[row], s1 = p.s1[row];
    const size_t drow = p.rm.drow[row];        // data row of this launch row (identity for whole polynomials, sparse for row subsets)
    const u64* p1 = op_uses_p1(OP) ? p.p1 + (size_t)b * p.bs1 + drow * p.rs1 : nullptr;
    const u64* p2 = op_uses_p2(OP) ? p.p2 + (size_t)b * p.bs2 + drow * p.rs2 : nullptr;
    u64* p3 = p.p3 + (size_t)b * p.bs3 + drow * p.rs3;
    const int nv = p.n / VEC;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
        if (VEC == 2) {
            ulonglong2 x = make_ulonglong2(0, 0), y = x, z = x;
            if (op_uses_p1(OP)) x = reinterpret_cast<const ulonglong2*>(p1)[i];
            if (op_uses_p2(OP)) y = reinterpret_cast<const ulonglong2*>(p2)[i];
            if (op_reads_p3(OP)) z = reinterpret_cast<const ulonglong2*>(p3)[i];
            ulonglong2 r;
            r.x = apply<OP>(x.x, y.x, z.x, L, s0, s1);
            r.y = apply<OP>(x.y, y.y, z.y, L, s0, s1);
            reinterpret_cast<ulonglong2*>(p3)[i] = r;
        } else {
            u64 x = 0, y = 0, z = 0;
            if (op_uses_p1(OP)) x = p1[i];
            if (op_uses_p2(OP)) y = p2[i];
            if (op_reads_p3(OP)) z = p3[i];
            p3[i] = apply<OP>(x, y, z, L, s0, s1);
        }
    }
}

template <int OP>
static int launch_one(const VecParams& p, int rows, int batch, bool vec2, cudaStream_t st) {
    const int nv = vec2 ? p.n / 2 : p.n;
    int blocks = (nv + 255) / 256;
    if (blocks > 64) blocks = (blocks + 3) / 4;  // 4 vectors per thread on large rows
    if (blocks < 1) blocks = 1;
    dim3 grid(blocks, rows, batch);
    if (vec2) vecop_kernel<OP, 2><<<grid, 256, 0, st>>>(p);
    else      vecop_kernel<OP, 1><<<grid, 256, 0, st>>>(p);
    LGPU_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_vecop(const Ctx* c, const RowMap& rm, int op, CSpan p1, CSpan p2, Span p3, int batch,
                 const u64* h_s0, const u64* h_s1, u64 s0, u64 s1, int n, cudaStream_t st) {
    if (op < 0 || op >= LGPU_OP_COUNT) { set_error("unknown vec opcode"); return -1; }
    if (rm.nrows <= 0 || rm.nrows > kMaxRows || batch <= 0 || batch > 65535) { set_error("bad rows/batch"); return -1; }
    if ((op_uses_p1(op) && !p1.p) || (op_uses_p2(op) && !p2.p) || !p3.p) { set_error("null operand"); return -1; }
    VecParams p;
    p.limbs = c->d_limbs; p.rm = rm;
    p.p1 = p1.p; p.p2 = p2.p; p.p3 = p3.p;
    p.rs1 = p1.row_stride; p.bs1 = p1.batch_stride; p.rs2 = p2.row_stride; p.bs2 = p2.batch_stride;
    p.rs3 = p3.row_stride; p.bs3 = p3.batch_stride;
    p.n = n;
    for (int i = 0; i < rm.nrows; i++) { p.s0[i] = h_s0 ? h_s0[i] : s0; p.s1[i] = h_s1 ? h_s1[i] : s1; }
    auto al = [](const void* q, size_t rs, size_t bs) { return ((uintptr_t)q % 16 == 0) && (rs % 2 == 0) && (bs % 2 == 0); };
    const bool vec2 = (n % 2 == 0) && al(p3.p, p3.row_stride, p3.batch_stride) &&
                      (!op_uses_p1(op) || al(p1.p, p1.row_stride, p1.batch_stride)) &&
                      (!op_uses_p2(op) || al(p2.p, p2.row_stride, p2.batch_stride));
    const int nops = 1 + (op_uses_p1(op) ? 1 : 0) + (op_uses_p2(op) ? 1 : 0) + (op_reads_p3(op) ? 1 : 0);
    ProfScope ps(LGPU_KCLASS_VECOP, st, 8.0 * n * rm.nrows * batch * nops, 1);
    switch (op) {
#define CASE(OPC) case OPC: return launch_one<OPC>(p, rm.nrows, batch, vec2, st);
        CASE(LGPU_OP_ADD) CASE(LGPU_OP_ADDLAZY) CASE(LGPU_OP_SUB) CASE(LGPU_OP_SUBLAZY) CASE(LGPU_OP_NEG)
        CASE(LGPU_OP_REDUCE) CASE(LGPU_OP_REDUCELAZY) CASE(LGPU_OP_MULCOEFFSLAZY) CASE(LGPU_OP_MULCOEFFSLAZYTHENADDLAZY)
        CASE(LGPU_OP_MULCOEFFSBARRETT) CASE(LGPU_OP_MULCOEFFSBARRETTLAZY) CASE(LGPU_OP_MULCOEFFSBARRETTTHENADD)
        CASE(LGPU_OP_MULCOEFFSBARRETTTHENADDLAZY) CASE(LGPU_OP_MULCOEFFSMONTGOMERY) CASE(LGPU_OP_MULCOEFFSMONTGOMERYLAZY)
        CASE(LGPU_OP_MULCOEFFSMONTGOMERYTHENADD) CASE(LGPU_OP_MULCOEFFSMONTGOMERYTHENADDLAZY)
        CASE(LGPU_OP_MULCOEFFSMONTGOMERYLAZYTHENADDLAZY) CASE(LGPU_OP_MULCOEFFSMONTGOMERYTHENSUB)
        CASE(LGPU_OP_MULCOEFFSMONTGOMERYTHENSUBLAZY) CASE(LGPU_OP_MULCOEFFSMONTGOMERYLAZYTHENSUBLAZY)
        CASE(LGPU_OP_MULCOEFFSMONTGOMERYLAZYTHENNEG) CASE(LGPU_OP_ADDLAZYTHENMULSCALARMONTGOMERY)
        CASE(LGPU_OP_ADDSCALARLAZYTHENMULSCALARMONTGOMERY) CASE(LGPU_OP_ADDSCALAR) CASE(LGPU_OP_ADDSCALARLAZY)
        CASE(LGPU_OP_ADDSCALARLAZYTHENNEGTWOMODULUSLAZY) CASE(LGPU_OP_SUBSCALAR) CASE(LGPU_OP_MULSCALARMONTGOMERY)
        CASE(LGPU_OP_MULSCALARMONTGOMERYLAZY) CASE(LGPU_OP_MULSCALARMONTGOMERYTHENADD)
        CASE(LGPU_OP_MULSCALARMONTGOMERYTHENADDSCALAR) CASE(LGPU_OP_SUBTHENMULSCALARMONTGOMERYTWOMODULUS)
        CASE(LGPU_OP_MFORM) CASE(LGPU_OP_MFORMLAZY) CASE(LGPU_OP_IMFORM) CASE(LGPU_OP_ZERO) CASE(LGPU_OP_MASK)
#undef CASE
    }
    set_error("unknown vec opcode");
    return -1;
}

}  // namespace lgpu
