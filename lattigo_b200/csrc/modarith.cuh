// modarith.cuh -- 64-bit modular arithmetic device primitives (sm_100a).
//
// The GPU has no 64-bit multiplier: every 64x64 product is built from 32-bit IMAD/IMAD.WIDE
// on the FMA pipe, so these primitives are the instruction-count floor of the whole engine.
// Semantics follow the reference's scalar primitives bit for bit where a lazy (non-canonical)
// representative can escape to the caller (ring/modular_reduction.go:11-205); the cheaper
// "free-form" variants (Shoup products, conditional-subtract canonicalisation) are only used
// where the value that leaves the kernel is the canonical residue in [0, q).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

namespace lgpu {

__device__ __forceinline__ u64 mulhi64(u64 a, u64 b) { return __umul64hi(a, b); }

__device__ __forceinline__ void mul128(u64 a, u64 b, u64& hi, u64& lo) {
    // one 128-bit product: the compiler shares the four 32x32 partial products between the two halves
    // (separate `a * b` and `__umul64hi(a, b)` recompute the low ones: ~15 more SASS instructions per 4-term MAC)
    const unsigned __int128 p = (unsigned __int128)a * b;
    lo = (u64)p;
    hi = (u64)(p >> 64);
}

// MRedLazy(x, y): x*y*2^-64 mod q in [0, 2q)        ring/modular_reduction.go:90-95
__device__ __forceinline__ u64 mred_lazy(u64 x, u64 y, u64 q, u64 qinv) {
    u64 hi, lo;
    mul128(x, y, hi, lo);
    u64 h = mulhi64(lo * qinv, q);
    return hi - h + q;
}
// MRed(x, y): canonical                               ring/modular_reduction.go:78-86
__device__ __forceinline__ u64 mred(u64 x, u64 y, u64 q, u64 qinv) {
    u64 r = mred_lazy(x, y, q, qinv);
    return r >= q ? r - q : r;
}
// Montgomery reduction of a 128-bit value (hi, lo) -> [0, 2q) given hi < q... (caller guarantees range)
__device__ __forceinline__ u64 mred128_lazy(u64 hi, u64 lo, u64 q, u64 qinv) {
    u64 h = mulhi64(lo * qinv, q);
    return hi - h + q;
}
// MFormLazy / MForm: a*2^64 mod q                      ring/modular_reduction.go:11-45
__device__ __forceinline__ u64 mform_lazy(u64 a, u64 q, u64 bhi, u64 blo) {
    u64 mhi = mulhi64(a, blo);
    return (0ull - (a * bhi + mhi)) * q;
}
__device__ __forceinline__ u64 mform(u64 a, u64 q, u64 bhi, u64 blo) {
    u64 r = mform_lazy(a, q, bhi, blo);
    return r >= q ? r - q : r;
}
// IMForm: a*2^-64 mod q                                ring/modular_reduction.go:49-56
__device__ __forceinline__ u64 imform(u64 a, u64 q, u64 qinv) {
    u64 r = q - mulhi64(a * qinv, q);
    return r >= q ? r - q : r;
}
// BRedAdd[Lazy]: a mod q for any a < 2^64              ring/modular_reduction.go:110-124
__device__ __forceinline__ u64 bred_add_lazy(u64 a, u64 q, u64 bhi) { return a - mulhi64(a, bhi) * q; }
__device__ __forceinline__ u64 bred_add(u64 a, u64 q, u64 bhi) {
    u64 r = bred_add_lazy(a, q, bhi);
    return r >= q ? r - q : r;
}
// BRed[Lazy]: x*y mod q                                ring/modular_reduction.go:127-196
__device__ __forceinline__ u64 bred_lazy(u64 x, u64 y, u64 q, u64 bhi, u64 blo) {
    u64 mhi, mlo;
    mul128(x, y, mhi, mlo);
    u64 r = mhi * bhi;
    u64 hhi, hlo;
    mul128(mlo, bhi, hhi, hlo);
    r += hhi;
    u64 lhi = mulhi64(mlo, blo);
    u64 s0 = hlo + lhi;
    r += (s0 < hlo);
    mul128(mhi, blo, hhi, hlo);
    r += hhi;
    u64 s1 = hlo + s0;
    r += (s1 < hlo);
    return mlo - r * q;
}
__device__ __forceinline__ u64 bred(u64 x, u64 y, u64 q, u64 bhi, u64 blo) {
    u64 r = bred_lazy(x, y, q, bhi, blo);
    return r >= q ? r - q : r;
}
__device__ __forceinline__ u64 cred(u64 a, u64 q) { return a >= q ? a - q : a; }

// canonicalise a value known to be < 8q (q < 2^61) with conditional subtractions only (ALU pipe, no IMAD)
__device__ __forceinline__ u64 csub_lt8q(u64 a, u64 q) {
    u64 q4 = q << 2, q2 = q << 1;
    a = a >= q4 ? a - q4 : a;
    a = a >= q2 ? a - q2 : a;
    return a >= q ? a - q : a;
}

}  // namespace lgpu
