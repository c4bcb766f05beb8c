// Device building blocks shared by the predict kernel and the persistent simulation kernel:
// Philox4x32-10, the AS241 inverse normal, and the lock-step tree walker over packed 8-byte slots
// (layout: fmc_pack.hpp).  Compiled with -fmad=false: float64 state arithmetic must round exactly
// like CPython evaluates the reference's expressions (one IEEE operation at a time).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "fmc_pack.hpp"   // kIlp, kRootWords, slot format constants

namespace fmc {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011).  Counter = (game lo, game hi, matchup,
// iteration << 2 | block), key = seed.  One call yields the four 32-bit words of one block of the
// 16-slot draw record.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ double u01(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }

// Wichura (1988), algorithm AS241 PPND16: inverse standard normal CDF, |error| < 1e-16 relative.
__device__ __noinline__ double ppnd16(double p) {
    const double q = p - 0.5;
    double r, num, den;
    if (fabs(q) <= 0.425) {
        r = 0.180625 - q * q;
        num = 2.5090809287301226727e+3;
        num = num * r + 3.3430575583588128105e+4;
        num = num * r + 6.7265770927008700853e+4;
        num = num * r + 4.5921953931549871457e+4;
        num = num * r + 1.3731693765509461125e+4;
        num = num * r + 1.9715909503065514427e+3;
        num = num * r + 1.3314166789178437745e+2;
        num = num * r + 3.3871328727963666080e0;
        den = 5.2264952788528545610e+3;
        den = den * r + 2.8729085735721942674e+4;
        den = den * r + 3.9307895800092710610e+4;
        den = den * r + 2.1213794301586595867e+4;
        den = den * r + 5.3941960214247511077e+3;
        den = den * r + 6.8718700749205790830e+2;
        den = den * r + 4.2313330701600911252e+1;
        den = den * r + 1.0;
        return q * num / den;
    }
    r = q < 0.0 ? p : 1.0 - p;
    r = sqrt(-log(r));
    if (r <= 5.0) {
        r -= 1.6;
        num = 7.74545014278341407640e-4;
        num = num * r + 2.27238449892691845833e-2;
        num = num * r + 2.41780725177450611770e-1;
        num = num * r + 1.27045825245236838258e0;
        num = num * r + 3.64784832476320460504e0;
        num = num * r + 5.76949722146069140550e0;
        num = num * r + 4.63033784615654529590e0;
        num = num * r + 1.42343711074968357734e0;
        den = 1.05075007164441684324e-9;
        den = den * r + 5.47593808499534494600e-4;
        den = den * r + 1.51986665636164571966e-2;
        den = den * r + 1.48103976427480074590e-1;
        den = den * r + 6.89767334985100004550e-1;
        den = den * r + 1.67638483018380384940e0;
        den = den * r + 2.05319162663775882187e0;
        den = den * r + 1.0;
    } else {
        r -= 5.0;
        num = 2.01033439929228813265e-7;
        num = num * r + 2.71155556874348757815e-5;
        num = num * r + 1.24266094738807843860e-3;
        num = num * r + 2.65321895265761230930e-2;
        num = num * r + 2.96560571828504891230e-1;
        num = num * r + 1.78482653991729133580e0;
        num = num * r + 5.46378491116411436990e0;
        num = num * r + 6.65790464350110377720e0;
        den = 2.04426310338993978564e-15;
        den = den * r + 1.42151175831644588870e-7;
        den = den * r + 1.84631831751005468180e-5;
        den = den * r + 7.86869131145613259100e-4;
        den = den * r + 1.48753612908506148525e-2;
        den = den * r + 1.36929880922735805310e-1;
        den = den * r + 5.99832206555887937690e-1;
        den = den * r + 1.0;
    }
    const double v = num / den;
    return q < 0.0 ? -v : v;
}

// float exp with a single rounding: double exp (<= 1 ulp of double) rounded once to float.
__device__ __forceinline__ float expf_cr(float x) { return (float)exp((double)x); }

// ---------------------------------------------------------------------------------------------
// Tree walker.  One lane = one row; all 32 lanes of a warp walk the SAME tree at the same time
// (different paths), so the slots a warp touches per level sit inside one small tree (at most
// 2^level distinct 8-byte words) -- broadcast-friendly in L1.  Three trees are in flight per lane
// for instruction-level parallelism.
//
// Slot format (fmc_pack.hpp): internal iff (int)hi >= 0x50000000; hi = tag | (4*row) << 20 | child;
// lo = float threshold.  Leaf: sklearn = the float64 itself, xgboost = float32 in lo.
//
// The feature row lives in shared memory and is addressed with a 32-bit shared-window address
// (`frow`, from __cvta_generic_to_shared) so that the per-level address math is one integer add;
// the table is addressed as base + 32-bit byte offset.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kTag = 0x50000000u;

__device__ __forceinline__ uint2 ldg_slot(const char *tbl, uint32_t slot) {
    uint2 v;
    asm("{\n\t.reg .u64 ad;\n\t"
        "mad.wide.u32 ad, %2, 8, %3;\n\t"
        "ld.global.nc.v2.u32 {%0, %1}, [ad];\n\t}"
        : "=r"(v.x), "=r"(v.y) : "r"(slot), "l"(tbl));
    return v;
}
__device__ __forceinline__ bool slot_internal(uint32_t hi) { return (int)hi >= (int)kTag; }

// One level of one tree for this lane; does nothing once the lane holds a leaf.
//   frow_biased = shared-window address of the lane's feature row minus 0x500 (tag bits of hi >> 20)
// Target per level: ISETP, LEA.HI (row address), LDS, FSET (0 / 0xFFFFFFFF), LOP3, IADD, IMAD.WIDE, LDG.
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
template <bool SKL>
__device__ __forceinline__ void walk_step(uint2 &n, const char *__restrict__ tbl, uint32_t frow_biased) {
    if (slot_internal(n.y)) {
        const float fv = lds_f32(frow_biased + (n.y >> 20));
        uint32_t m;   // 0xFFFFFFFF when the lane goes right
        if (SKL) asm("set.gtu.u32.f32 %0, %1, %2;" : "=r"(m) : "f"(fv), "f"(__uint_as_float(n.x)));   // !(fv <= thr)
        else asm("set.geu.u32.f32 %0, %1, %2;" : "=r"(m) : "f"(fv), "f"(__uint_as_float(n.x)));        // !(fv < thr)
        n = ldg_slot(tbl, (n.y & 0xFFFFFu) - m);
    }
}

// Sum one output of a packed forest over `rounds_padded` trees (a multiple of kIlp), in tree order.
// SKL: float64 accumulate of pre-scaled leaves; XGB: float32 accumulate (returned widened).
// `frow` is the 32-bit shared address of this lane's feature row.  `roots` is the inline copy of the
// root slots, two trees per uint4, in tree order: every lane of the warp reads the same words (one
// broadcast load per two trees) and the next group is fetched while the current one is walked.
template <bool SKL, int MAX_DEPTH>
__device__ __forceinline__ double walk_output(const uint2 *__restrict__ slots, const uint4 *__restrict__ roots,
                                              int rounds_padded, uint32_t frow, double base) {
    constexpr int I = kIlp;
    static_assert(I % 2 == 0, "kIlp must be even (two root slots per uint4)");
    const char *tbl = reinterpret_cast<const char *>(slots);
    uint32_t fb = frow - 0x500u;
    asm volatile("" : "+r"(fb));          // keep the row address in a register (no re-materialisation)
    double acc64 = base;
    float acc32 = (float)base;
    const int groups = rounds_padded / I;
    uint2 n[I];
#pragma unroll
    for (int q = 0; q < I / 2; ++q) {
        const uint4 v = __ldg(roots + q);
        n[2 * q] = make_uint2(v.x, v.y);
        n[2 * q + 1] = make_uint2(v.z, v.w);
    }
#pragma unroll 1
    for (int t = 0; t < groups; ++t) {
        uint2 nx[I];
        const uint4 *next = roots + (size_t)(t + 1 < groups ? t + 1 : t) * (I / 2);
#pragma unroll
        for (int q = 0; q < I / 2; ++q) {
            const uint4 v = __ldg(next + q);
            nx[2 * q] = make_uint2(v.x, v.y);
            nx[2 * q + 1] = make_uint2(v.z, v.w);
        }
        if (MAX_DEPTH <= 4) {
#pragma unroll
            for (int d = 0; d < MAX_DEPTH; ++d) {
#pragma unroll
                for (int i = 0; i < I; ++i) walk_step<SKL>(n[i], tbl, fb);
            }
        } else {
            for (;;) {
                bool any;
                if (SKL) {
                    any = false;
#pragma unroll
                    for (int i = 0; i < I; ++i) any |= slot_internal(n[i].y);
                } else {          // xgboost leaves carry hi == 0, so OR-ing the words keeps the test exact
                    uint32_t o = 0;
#pragma unroll
                    for (int i = 0; i < I; ++i) o |= n[i].y;
                    any = slot_internal(o);
                }
                if (!__any_sync(0xFFFFFFFFu, any)) break;
#pragma unroll
                for (int i = 0; i < I; ++i) walk_step<SKL>(n[i], tbl, fb);
            }
        }
#pragma unroll
        for (int i = 0; i < I; ++i) {
            if (SKL) acc64 = __dadd_rn(acc64, __hiloint2double((int)n[i].y, (int)n[i].x));
            else acc32 = __fadd_rn(acc32, __uint_as_float(n[i].x));
            n[i] = nx[i];
        }
    }
    return SKL ? acc64 : (double)acc32;
}

}  // namespace fmc
