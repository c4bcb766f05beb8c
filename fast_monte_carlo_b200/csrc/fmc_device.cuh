// Device building blocks shared by the predict kernel and the persistent simulation kernel:
// Philox4x32-10, the AS241 inverse normal, and the lock-step tree walker over packed 8-byte slots
// (layout: fmc_pack.hpp).  Compiled with -fmad=false: float64 state arithmetic must round exactly
// like CPython evaluates the reference's expressions (one IEEE operation at a time).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "fmc_pack.hpp"   // kIlp, slot format constants

namespace fmc {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011).  Counter = (game lo, game hi, matchup,
// iteration << 2 | block), key = seed.  One call yields the four 32-bit words of one block of the
// 16-slot draw record.
// ---------------------------------------------------------------------------------------------
// (not inlined: the draw sites are many and divergent, one shared copy keeps the state machine's
// instruction footprint -- its real bottleneck, `stall_no_inst` 33 % -- small; +2.5 %)
__device__ __noinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
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
// Tree walker.  One lane = one request; all 32 lanes of a warp walk the SAME trees at the same
// time (different paths), kIlp trees in flight per lane.  Layout of the node table, the root
// stream and the constants side stream: fmc_pack.hpp.
//
// The walk is branch-free per lane: a group of kIlp trees is walked for exactly D levels (D is in
// the group's metadata, warp-uniform), leaves keep a lane in place (xgboost: self-pointing leaf;
// sklearn: pass-through chain), so one level of one tree is
//      LEA.HI   feature address  = chunk column of the lane + (hi >> 20)
//      LDS      feature value                       (feature-major rows: never a bank conflict)
//      LOP3     left-child address = (hi & mask) | window base
//      FSETP + predicated add of 8 (right child)
//      MOV      high half of the 64-bit address      (the window's, constant)
//      LDG.64   next slot                            (L1/L2-resident table, read-only path)
// The feature rows of a 32-request chunk live in shared memory as [feature][lane]; `fcol` is the
// 32-bit shared-window address of the lane's column (chunk base + 4 * lane).
// ---------------------------------------------------------------------------------------------
struct ForestView {          // everything here is warp-uniform
    uint32_t win_lo, win_hi; // address of the table's 1 MiB window
    const uint4 *stream;     // root stream of ONE output: 2 x uint4 (kIlp root slots) per group
    const uint2 *consts;     // side stream of that output
    uint32_t n_groups;
    bool multi_window;       // the node table spans more than one window
};

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 ldg_slot(uint32_t lo, uint32_t hi) {
    uint2 v;
    asm("{\n\t.reg .u64 ad;\n\tmov.b64 ad, {%2, %3};\n\tld.global.nc.v2.u32 {%0, %1}, [ad];\n\t}"
        : "=r"(v.x), "=r"(v.y) : "r"(lo), "r"(hi));
    return v;
}

#ifdef FMC_DEBUG_CHECKS
// Diagnostics build (compute-sanitizer is not available everywhere): every gather address is checked against
// the node arena of the launch and every feature offset against the widest row layout; violations are counted
// (fmc_debug_errors) and the offending load is skipped.
__device__ unsigned long long g_dbg_lo = 0, g_dbg_hi = 0, g_dbg_errors = 0;
#endif

template <bool SKL>
__device__ __forceinline__ void walk_step(uint2 &n, uint32_t fcol, uint32_t win_lo, uint32_t win_hi) {
#ifdef FMC_DEBUG_CHECKS
    if ((n.y >> kFeatShift) >= (uint32_t)(kPredRows * kFeatBytes) || ((n.y >> kFeatShift) & 127u)) { atomicAdd(&g_dbg_errors, 1ULL); return; }
#endif
    const float fv = lds_f32(fcol + (n.y >> kFeatShift));
    uint32_t a = (n.y & kChildMask) | win_lo;
    if (SKL)   // right iff !(fv <= thr)
        asm("{\n\t.reg .pred p;\n\tsetp.gtu.f32 p, %1, %2;\n\t@p add.u32 %0, %0, 8;\n\t}" : "+r"(a) : "f"(fv), "f"(__uint_as_float(n.x)));
    else       // right iff !(fv < thr)
        asm("{\n\t.reg .pred p;\n\tsetp.geu.f32 p, %1, %2;\n\t@p add.u32 %0, %0, 8;\n\t}" : "+r"(a) : "f"(fv), "f"(__uint_as_float(n.x)));
#ifdef FMC_DEBUG_CHECKS
    {
        const unsigned long long ad = ((unsigned long long)win_hi << 32) | a;
        if (ad < g_dbg_lo || ad + 8 > g_dbg_hi || (ad & 7)) { atomicAdd(&g_dbg_errors, 1ULL); return; }
    }
#endif
    n = ldg_slot(a, win_hi);
}


// Adds the leaves the lanes hold after a group's walk (plus the group's constants) in tree order.
template <bool SKL>
__device__ __forceinline__ void accumulate_group(const uint2 (&n)[kIlp], bool has_consts, const uint2 *&cp,
                                                 double &acc64, float &acc32) {
    if (!has_consts) {
#pragma unroll
        for (int i = 0; i < kIlp; ++i) {
            if (SKL) acc64 = __dadd_rn(acc64, __hiloint2double((int)n[i].y, (int)n[i].x));
            else acc32 = __fadd_rn(acc32, __uint_as_float(n[i].x));
        }
        return;
    }
    // constants (trees folded to one leaf) take their place in the tree order
    const uint32_t counts = __ldg(cp).x;
    ++cp;
#pragma unroll
    for (int i = 0; i < kIlp; ++i) {
        const uint32_t c = (counts >> (8 * i)) & 0xFFu;
#pragma unroll 1
        for (uint32_t j = 0; j < c; ++j) {
            const uint2 v = __ldg(cp);
            ++cp;
            if (SKL) acc64 = __dadd_rn(acc64, __hiloint2double((int)v.y, (int)v.x));
            else acc32 = __fadd_rn(acc32, __uint_as_float(v.x));
        }
        if (SKL) acc64 = __dadd_rn(acc64, __hiloint2double((int)n[i].y, (int)n[i].x));
        else acc32 = __fadd_rn(acc32, __uint_as_float(n[i].x));
    }
}

__device__ __forceinline__ void load_roots(const uint4 *sp, uint2 (&n)[kIlp]) {
    const uint4 a = __ldg(sp), b = __ldg(sp + 1);
    n[0] = make_uint2(a.x, a.y); n[1] = make_uint2(a.z, a.w);
    n[2] = make_uint2(b.x, b.y); n[3] = make_uint2(b.z, b.w);
}
__device__ __forceinline__ uint32_t group_depth(const uint2 (&n)[kIlp]) { return (n[0].y & 7u) | ((n[1].y & 1u) << 3); }
// 64-bit address of the 1 MiB window a group's nodes sit in.  MW = false: the table fits one window (every
// table of the simulation so far); MW = true: tables above 1 MiB span several windows.
template <bool MW>
__device__ __forceinline__ void group_window(const uint2 (&n)[kIlp], const ForestView &F, uint32_t &lo, uint32_t &hi) {
    if (!MW) { lo = F.win_lo; hi = F.win_hi; return; }
    const uint32_t w = (n[2].y & 7u) | ((n[3].y & 7u) << 3);
    const unsigned long long a = (((unsigned long long)F.win_hi << 32) | F.win_lo) + ((unsigned long long)w << 20);
    lo = (uint32_t)a;
    hi = (uint32_t)(a >> 32);
}
__device__ __forceinline__ bool group_has_consts(const uint2 (&n)[kIlp]) { return (n[1].y & 2u) != 0; }

// Sum one output of a packed forest in tree order.  SKL: float64 accumulate of pre-scaled leaves;
// XGB: float32 accumulate (returned widened).
// `levels` returns the number of tree levels the walk issued per tree slot (sum of the group depths).
template <bool SKL, bool MW>
__device__ __forceinline__ double walk_output(const ForestView &F, uint32_t fcol, int lane, double base, uint32_t &levels) {
    constexpr int I = kIlp;
    double acc64 = base;
    float acc32 = (float)base;
    levels = 0;
    if (F.n_groups == 0) return SKL ? acc64 : (double)acc32;
    asm volatile("" : "+r"(fcol));      // opaque: keep the column address in a register instead of re-deriving it per loop
    const uint4 *sp = F.stream;
    const uint2 *cp = F.consts;
    // Two groups (2 x kIlp trees) in flight per lane: both are walked together for the depth they share
    // (8 independent gather chains, no branch inside), then the deeper one finishes alone -- a sklearn
    // lane must not step past its float64 leaf.  The roots are fetched at the top of each pass.
    uint32_t g = 0;
#pragma unroll 1
    for (; g + 1 < F.n_groups; g += 2) {
        uint2 n0[I], n1[I];
        load_roots(sp + 2 * g, n0);
        load_roots(sp + 2 * g + 2, n1);
        const uint32_t d0 = group_depth(n0), d1 = group_depth(n1);
        const bool c0 = group_has_consts(n0), c1 = group_has_consts(n1);   // the metadata leaves with the root slots
        uint32_t w0l, w0h, w1l, w1h;
        group_window<MW>(n0, F, w0l, w0h);
        group_window<MW>(n1, F, w1l, w1h);
        const uint32_t dmin = d0 < d1 ? d0 : d1, dmax = d0 < d1 ? d1 : d0;
        levels += d0 + d1;
#pragma unroll 1
        for (uint32_t d = 0; d < dmin; ++d) {
#pragma unroll
            for (int i = 0; i < I; ++i) walk_step<SKL>(n0[i], fcol, w0l, w0h);
#pragma unroll
            for (int i = 0; i < I; ++i) walk_step<SKL>(n1[i], fcol, w1l, w1h);
        }
        if (d0 > d1) {
#pragma unroll 1
            for (uint32_t d = dmin; d < dmax; ++d)
#pragma unroll
                for (int i = 0; i < I; ++i) walk_step<SKL>(n0[i], fcol, w0l, w0h);
        } else {
#pragma unroll 1
            for (uint32_t d = dmin; d < dmax; ++d)
#pragma unroll
                for (int i = 0; i < I; ++i) walk_step<SKL>(n1[i], fcol, w1l, w1h);
        }
        accumulate_group<SKL>(n0, c0, cp, acc64, acc32);
        accumulate_group<SKL>(n1, c1, cp, acc64, acc32);
    }
    if (g < F.n_groups) {
        uint2 n0[I];
        load_roots(sp + 2 * g, n0);
        const uint32_t d0 = group_depth(n0);
        const bool c0 = group_has_consts(n0);
        uint32_t w0l, w0h;
        group_window<MW>(n0, F, w0l, w0h);
        levels += d0;
#pragma unroll 1
        for (uint32_t d = 0; d < d0; ++d)
#pragma unroll
            for (int i = 0; i < I; ++i) walk_step<SKL>(n0[i], fcol, w0l, w0h);
        accumulate_group<SKL>(n0, c0, cp, acc64, acc32);
    }
    (void)lane;
    return SKL ? acc64 : (double)acc32;
}

}  // namespace fmc
