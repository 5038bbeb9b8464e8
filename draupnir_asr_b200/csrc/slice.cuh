// Fixed-point byte slices of panel rows (shared by oz_slice_kernel in ozaki.cu and the fused slicers in factor.cu, which
// must stay bit-identical).  A row with maximum magnitude m = f 2^e (f in [0.5, 1)) is scaled by 2^(8S-1-e), rounded to
// nearest-even integers q (|q| <= 2^(8S-1), clamped to 2^(8S-1) - 1) and cut into S bytes, most significant first.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// q[j] = clamp(rint(x[j] 2^(8S-1-e))).  Fast path (S <= 6, i.e. |q| < 2^48, and a row scale inside the exponent range): the
// scaling is one multiplication by an exactly representable power of two fused with the 1.5 * 2^52 shift whose sum holds
// rint(.) in its low mantissa bits - two fp64-pipe operations instead of scalbn + F2I.S64.F64 (the latter runs on the XU
// pipe), same integers.  Everything else (zero / huge / non-finite rows, S = 7) takes the library route.
__device__ __forceinline__ void drp_quant4(const double (&x)[4], double m, int e, int S, long long (&q)[4])
{
    const int sh = 8 * S - 1 - e;
    const long long hi = (1LL << (8 * S - 1)) - 1;
    if (S <= 6 && m > 0.0 && m < 1e300 && e > -900 && e < 900) {
        const double scale = __hiloint2double((1023 + sh) << 20, 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double v = fma(x[j], scale, 6755399441055744.0);
            const long long t = __double_as_longlong(v) - 0x4338000000000000LL;
            q[j] = t > hi ? hi : t;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long t = __double2ll_rn(scalbn(x[j], sh));
            q[j] = t > hi ? hi : (t < -hi - 1 ? -hi - 1 : t);
        }
    }
}

// byte `bi` (0 = least significant) of each of the four integers, packed: entry j -> byte j of the result
__device__ __forceinline__ uint32_t drp_pick4(const long long (&q)[4], int bi)
{
    uint32_t p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        p[j] = __byte_perm((uint32_t)q[j], (uint32_t)((unsigned long long)q[j] >> 32), (uint32_t)bi);   // wanted byte -> byte 0
    return __byte_perm(__byte_perm(p[0], p[1], 0x0040), __byte_perm(p[2], p[3], 0x0040), 0x5410);
}
