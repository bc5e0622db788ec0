// Optical depth accumulation (replaces part_int.cpp:20-51 + absorption.cpp:212-279 +
// singleabs.h:63-175 + the Re w(z) consumer of Faddeeva.cpp).
//
// Decomposition.  Persistent CTAs of kTauWarps warps pull work items (sightline, run of its
// candidate list) from a global counter.  Per batch of kBatch particles one lane per particle
// gathers the particle and derives every per-particle constant into a shared-memory slab
// (field-major, so the later warp-uniform reads are broadcasts).  Then, particle by particle, the
// 32 lanes of the warp march outward over its pixels (absorption.cpp:250-278): while both
// directions are live each gets a half warp, afterwards all 32 lanes serve the remaining one; the
// reference's "add, then stop once taulast < tautail" rule is a ballot + find-first-set.
//
// NL lines of one ion (Lya + Lyb ...) are fused in one pass: they share every node position, the
// Gaussian and the Dawson-type table value; only the damping parameter y (hence the coefficient
// sets Pe, A, BQ) and the amplitude differ.
//
// Per pixel the 7-node kernel x Voigt quadrature (singleabs.h:143-167) takes one of four
// warp-uniform routes:
//   NEAR   all nodes of all lanes inside the table (|x| < 24): branch-free, the 7 nodes interleaved for ILP;
//          G(x) from the shared-memory table, Gaussians by a two-level recurrence (across the nodes
//          of a pixel, and from one march step to the next: no exp() in steady state)
//   FAR    all nodes at |x| >= 12: damping-wing series in 1/x^2, shared across the fused lines
//   MIXED  the rare warp step that straddles |x| = 16, and pixels wider than btherm/2
//          (sub-sampling rule of singleabs.h:110-125): generic per-node evaluation
//   EXACT  FSB_VOIGT_EXACT or y outside (1e-30, 0.03]: restatement of the reference's Faddeeva::w
#include <stdlib.h>
#include <string.h>

#include "fsb_items.cuh"
#include "fsb_voigt.cuh"

namespace fsb {

namespace {

#ifndef FSB_TAU_MIN_BLOCKS
#define FSB_TAU_MIN_BLOCKS 1  // resident CTAs per SM the register allocation must allow (x FSB_TAU_WARPS = 16 warps per SM)
#endif
#ifndef FSB_TAU_BATCH
#define FSB_TAU_BATCH 16
#endif

#ifndef FSB_TAU_WARPS
#define FSB_TAU_WARPS 16  // one persistent CTA per SM: one copy of the G(x) table per SM; measured 4 % faster than 4 x 4 warps
#endif
constexpr int kTauWarps = FSB_TAU_WARPS;
constexpr int kTauThreads = 32 * kTauWarps;
constexpr int kBatch = FSB_TAU_BATCH;  // particles per slab refill (one lane each), <= 32
// A single line in FP64 has room for 32 records per warp: the per-particle setup then runs on full warps, which
// halves its share of the instruction stream (weak lines spend a fifth of their instructions there).
template <int NL, bool F32> struct BatchOf { static constexpr int value = (NL == 1 && !F32 && kBatch == 16) ? 32 : kBatch; };
constexpr int kMaxTauLines = 2;

// ---- slab layout: one record of doubles per particle, kBatch records per warp ------------------------
// The march is bound by shared-memory wavefronts, so fields that are read together sit in 16-byte pairs
// (even index first) and are fetched with one 128-bit broadcast load.
enum SharedField {
    S_STEP = 0,  // node spacing in units of btherm: (2 vhigh/8)/btherm
    S_XB0,       // xb of the centre of pixel zmax: -vhigh/b - ((zmax + 1/2) bintov - vel)/b
    S_PIX,       // pixel width in units of btherm
    S_ZMAX,      // int2 {zmax = floor(vel/bintov), zmax mod nbins}
    S_THR_N,     // int2 {up, down}: outward pixels o < N have all nodes inside the table (|x| < 24)
    S_THR_F,     // int2: outward pixels o >= F have all nodes on the wing series (|x| >= 12)
    S_THR_G,     // int2: outward pixels o >= G are out of reach of the Gaussian
    S_RECOK,     // int2 {1 when the march-step recurrence is safe (all factors within e^+-500), degree class: 1 = the
                 // cubic terms of A(s) and Pe(s) matter for this particle, 0 = they are below 5e-13 of the profile}
    S_Q,         // exp(-2 step^2): second-order ratio of the Gaussian recurrence across nodes
    S_K16,       // exp(-2 D^2), D = 16 pixels in units of btherm: march-step recurrence of the Gaussian
    S_LU16,      // exp(+2 D step): ratio update of the inter-node factor, upward march
    S_LD16,      // exp(-2 D step): downward march
    S_KW0,       // 7 kernel weights x deltav                           singleabs.h:152-163
    S_MODE = S_KW0 + 7,  // 0 skip, 1 fast, 2 exact, 3 sub-sampled pixels (per-pixel fallback), 4 sub-sampled pixels (fast sums)
    S_VEL,       // velfac*pos + pvel                                  absorption.cpp:234
    S_INVB,      // 1/btherm
    S_HALFB,     // btherm/2: sub-sampling threshold                    singleabs.h:110
    S_XOFF,      // -vhigh/btherm
    S_XU2,       // x^2 beyond which exp(-x^2) is negligible against the damping wing
    S_PAD,       // mode 4: number of inner points per pixel (singleabs.h:116)
    S_COUNT
};
// Per fused line.  The profile is H = U Pe(s) + G A(s) + B(s) (fsb_voigt.cuh); every coefficient set below is
// already multiplied by the line's amplitude CD = amp dens / velfac, so node sums come out as optical depths.
enum LineField {
    L_AC0 = 0,   // CD x A(s): 4 coefficients
    L_PC0 = 4,   // CD x Pe(s): 4 coefficients
    L_BQ0 = 8,   // CD x sum_i kw_i B(s_i) as a quartic in xb: 5 coefficients
    L_FAR = 13,  // CD y / sqrt(pi): amplitude of the damping-wing series
    L_Y2,        // y^2
    L_CD,        // amp*dens/velfac
    L_BC0,       // CD x B(s): 3 raw coefficients (generic route)
    L_Y = L_BC0 + 3,  // aa = voigt_fac/btherm
    L_ERFCX,     // exact mode: erfcx(aa)
    L_COUNT = L_ERFCX + 2
};
static_assert(L_FAR == L_BQ0 + 5 && L_BQ0 % 2 == 0 && L_Y2 % 2 == 0 && L_BC0 % 2 == 0 && S_COUNT % 2 == 0 && L_COUNT % 2 == 0 &&
              S_KW0 % 2 == 0 && S_Q % 2 == 0 && S_LU16 % 2 == 0, "16-byte field pairs");
template <int NL, bool F32 = false> struct SlabSize {
    static constexpr int kFields = S_COUNT + NL * L_COUNT;
    // record stride in doubles: an odd number of 16-byte units, so the setup lanes spread over the banks
    static constexpr int kStride = (kFields / 2) % 2 ? kFields : kFields + 2;
    static constexpr int kDoubles = kStride * BatchOf<NL, F32>::value;
};

// FP32 fast path: float copies of the node constants, one record of floats per particle after the doubles
enum FShared { F_STEP = 0, F_KW0, F_COUNT = F_KW0 + 7 };
enum FLineField { FL_A0 = 0, FL_PE0 = 3, FL_BQ0 = 6, FL_Y2 = 11, FL_YISP, FL_COUNT };
template <int NL> struct FSlabSize {
    static constexpr int kStride = F_COUNT + NL * FL_COUNT;
    static constexpr int kFloats = kStride * kBatch;  // (FP32 instantiations keep kBatch records)
};
#define FS(f) fl[(f)]
#define FLF(l, f) fl[F_COUNT + (l) * FL_COUNT + (f)]

#define SF(f) sl[(f)]
#define LF(l, f) sl[S_COUNT + (l) * L_COUNT + (f)]
#define SF2(f) (*reinterpret_cast<const double2 *>(sl + (f)))
#define LF2(l, f) (*reinterpret_cast<const double2 *>(sl + S_COUNT + (l) * L_COUNT + (f)))

// ---- node sums ------------------------------------------------------------------------------------
// All return tau_l = CD_l sum_i kw_i H(x_i, y_l) for x_i = xb + (i+1) step, per fused line l.
//
// The sums are taken in MOMENT form: with H = U Pe_l(s) + G A_l(s) + B_l(s), s = x^2, and Pe_l, A_l polynomials
// in s whose coefficients depend on the line only,
//     sum_i kw_i H(x_i, y_l) = sum_k pe_lk MU_k + sum_k a_lk MG_k + BQ_l(xb),
//     MU_k = sum_i kw_i U(x_i) s_i^k,   MG_k = sum_i kw_i G(x_i) s_i^k,
// so the per-node work (table lookup, Gaussian recurrence, 2 (DEG + 1) accumulations) does not depend on the
// number of fused lines, and a line costs 2 (DEG + 1) + 5 multiply-adds per PIXEL.  DEG = 3 keeps the cubic terms
// of A and Pe; DEG = 2 drops them for particles whose damping parameter makes them < 5e-13 of the profile
// (setup_particle decides; nearly all H I gas above 2000 K).

// NEAR: every node inside the table (|x| < 24).  U0 = exp(-x_1^2), R = exp(-(2 x_1 + step) step), q = exp(-2 step^2)
// (GAUSS = false when the Gaussian is negligible for the whole warp step).  Lanes with nodes beyond the table
// compute finite garbage that the caller discards.
template <int NL, int DEG, bool GAUSS>
__device__ __forceinline__ void node_sum_near_m(double xb, double step, const double *__restrict__ sl, const double2 *__restrict__ tabA,
                                                double U0, double R, double q, unsigned lmask, double (&tot)[NL])
{
    double MG[DEG + 1], MU[DEG + 1];
    #pragma unroll
    for (int k = 0; k <= DEG; ++k) MG[k] = 0, MU[k] = 0;
    double u = U0, r = R;
    #pragma unroll
    for (int i = 0; i < 7; ++i) {
        const double x = fma((double) (i + 1), step, xb);
        int k;
        double t;
        g_index(fabs(x), k, t);
        k = (int) min((unsigned) k, (unsigned) (FSB_GTAB_NINT - 1));
        const double2 *e = tabA + 3 * k;
        const double g = g_poly(e[0], e[1], e[2], t);
        const double s = x * x;
        double w[DEG + 1];
        w[0] = SF(S_KW0 + i);
        #pragma unroll
        for (int k2 = 1; k2 <= DEG; ++k2) w[k2] = w[k2 - 1] * s;
        #pragma unroll
        for (int k2 = 0; k2 <= DEG; ++k2) MG[k2] = fma(g, w[k2], MG[k2]);
        if (GAUSS) {
            #pragma unroll
            for (int k2 = 0; k2 <= DEG; ++k2) MU[k2] = fma(u, w[k2], MU[k2]);
            u *= r;
            r *= q;
        }
    }
    #pragma unroll
    for (int l = 0; l < NL; ++l) {
        if (NL > 1 && !((lmask >> l) & 1u)) {
            tot[l] = 0;
            continue;
        }
        // sum_i kw_i B(s_i): a quartic in xb
        const double2 b01 = LF2(l, L_BQ0), b23 = LF2(l, L_BQ0 + 2);
        double acc = fma(fma(fma(fma(LF(l, L_BQ0 + 4), xb, b23.y), xb, b23.x), xb, b01.y), xb, b01.x);
        const double2 a01 = LF2(l, L_AC0), a23 = LF2(l, L_AC0 + 2);
        acc = fma(a01.x, MG[0], acc);
        acc = fma(a01.y, MG[1], acc);
        acc = fma(a23.x, MG[2], acc);
        if (DEG > 2) acc = fma(a23.y, MG[3], acc);
        if (GAUSS) {
            const double2 p01 = LF2(l, L_PC0), p23 = LF2(l, L_PC0 + 2);
            double acu = p01.x * MU[0];
            acu = fma(p01.y, MU[1], acu);
            acu = fma(p23.x, MU[2], acu);
            if (DEG > 2) acu = fma(p23.y, MU[3], acu);
            acc += acu;
        }
        tot[l] = acc;
    }
}

template <int NL>
__device__ __forceinline__ void node_sum_near(double xb, double step, const double *__restrict__ sl,
                                              const double2 *__restrict__ tabA, double U0, double R, double q, bool gauss,
                                              bool cubic, unsigned lmask, double (&tot)[NL])
{
    if (cubic) {
        if (gauss) node_sum_near_m<NL, 3, true>(xb, step, sl, tabA, U0, R, q, lmask, tot);
        else node_sum_near_m<NL, 3, false>(xb, step, sl, tabA, U0, R, q, lmask, tot);
    } else {
        if (gauss) node_sum_near_m<NL, 2, true>(xb, step, sl, tabA, U0, R, q, lmask, tot);
        else node_sum_near_m<NL, 2, false>(xb, step, sl, tabA, U0, R, q, lmask, tot);
    }
}

// FAR: every node at |x| >= 12 (the Gaussian is < e^-144).  Lanes with nodes inside compute garbage
// (possibly inf/NaN) that the caller discards.  Also in moment form:
//   H = (y/sqrt(pi)) u [P1(u) - v (P3(u) - v P5(u))], u = 1/x^2, v = y^2 u <= 3.6e-6 (far_polys, fsb_voigt.cuh), hence
//   sum_i kw_i H = (y/sqrt(pi)) [F1 - y^2 F3 + y^4 F5],  F1 = sum kw u P1, F3 = sum kw u^2 P3, F5 = sum kw u^3 P5.
// With u <= 1/144: the u^3.. tail of P1 (<= 5e-6 of P1, ten terms in all) runs in FP32, P3 stops at u^4 and P5 = 1
// (what is dropped is below 3e-12 of H at y = 0.03, |x| = 12, and falls with y^2 and 1/x^2).
template <int NL>
__device__ __forceinline__ void node_sum_far(double xb, double step, const double *__restrict__ sl, unsigned lmask,
                                             double (&tot)[NL])
{
    double F1 = 0, F3 = 0, F5 = 0;
    #pragma unroll
    for (int i = 0; i < 7; ++i) {
        const double x = fma((double) (i + 1), step, xb);
        const double u = fast_rcp(x * x);
        const float uf = (float) u;
        const float tail = fmaf(fmaf(fmaf(fmaf(fmaf(fmaf(1278767.75f, uf, 134607.125f), uf, 15836.1328125f), uf, 2111.484375f), uf, 324.84375f), uf, 59.0625f), uf, 13.125f);
        const double p1 = fma(fma(fma((double) tail, u, 3.75), u, 1.5), u, 1.0);
        const double p3 = fma(fma(fma(fma(1082.8125, u, 157.5), u, 26.25), u, 5.0), u, 1.0);
        const double w1 = SF(S_KW0 + i) * u, w2 = w1 * u;
        F1 = fma(w1, p1, F1);
        F3 = fma(w2, p3, F3);
        F5 = fma(w2, u, F5);
    }
    #pragma unroll
    for (int l = 0; l < NL; ++l) {
        if (NL > 1 && !((lmask >> l) & 1u)) {
            tot[l] = 0;
            continue;
        }
        const double y2 = LF(l, L_Y2);
        tot[l] = LF(l, L_FAR) * fma(-y2, fma(-y2, F5, F3), F1);
    }
}

// MIXED: the warp step straddles the table / series overlap (12 <= |x| < 24), or a pixel's own nodes do (kernels
// much wider than the thermal width: metal lines).  Node by node, with a warp-uniform choice per node: the lanes
// of the downward run take the nodes in mirror order (node 8 - n where the upward run takes node n), so that at
// every loop index all 32 lanes sit at nearly the same |x| (the two runs are mirror images about the particle up
// to a pixel): either every lane is inside the table (|x| < 24) or every lane is on the series (|x| >= 12), unless
// 16 pixels are wider than the overlap, in which case both are evaluated with the weights masked.
// Table nodes feed MG, MU (Gaussian by direct exp where some lane is within its reach) and MW_k = sum kw s^k
// (for B(s), which the all-table route takes from the quartic in xb); series nodes feed F1, F3, F5.
// `active`: lanes whose result is used (the others compute garbage that must not steer the uniform choices).
template <int NL>
__device__ __noinline__ void node_sum_mixed(double xb, double step, const double *__restrict__ sl, const double2 *__restrict__ tab,
                                            bool down, bool active, unsigned lmask, double (&tot)[NL])
{
    double MG[4] = {0, 0, 0, 0}, MU[4] = {0, 0, 0, 0}, MW[3] = {0, 0, 0}, F1 = 0, F3 = 0, F5 = 0;
    const double xu2 = SF(S_XU2);
    #pragma unroll 1
    for (int n = 0; n < 7; ++n) {
        const int ni = down ? 6 - n : n;
        const double x = fma((double) (ni + 1), step, xb), ax = fabs(x), s = x * x, kw = SF(S_KW0 + ni);
        const bool lane_far = ax >= kFarXMin, lane_tab = ax < FSB_GTAB_XMAX - 0.005;
        const bool all_far = __all_sync(kFull, lane_far || !active), all_tab = __all_sync(kFull, lane_tab || !active);
        if (!all_far) {  // table (all lanes, or masked to the lanes below the series' start when not every lane is inside)
            const double wt = (all_tab || !lane_far) ? kw : 0.0;
            const double g = g_table(fmin(ax, FSB_GTAB_XMAX - 0.005), tab);
            const double w1 = wt * s, w2 = w1 * s, w3 = w2 * s;
            MG[0] = fma(g, wt, MG[0]), MG[1] = fma(g, w1, MG[1]), MG[2] = fma(g, w2, MG[2]), MG[3] = fma(g, w3, MG[3]);
            MW[0] += wt, MW[1] += w1, MW[2] += w2;
            if (__any_sync(kFull, active && s < xu2)) {
                const double u = s < xu2 ? fast_exp(-s) : 0.0;
                MU[0] = fma(u, wt, MU[0]), MU[1] = fma(u, w1, MU[1]), MU[2] = fma(u, w2, MU[2]), MU[3] = fma(u, w3, MU[3]);
            }
        }
        if (all_far || !all_tab) {  // series (all lanes, or masked to the lanes the table did not take)
            const double wf = (all_far || lane_far) ? kw : 0.0;
            const double u = fast_rcp(fmax(s, kFarXMin * kFarXMin));
            double p1, p3, p5;
            far_polys(u, p1, p3, p5);
            const double w1 = wf * u, w2 = w1 * u;
            F1 = fma(w1, p1, F1), F3 = fma(w2, p3, F3), F5 = fma(w2 * u, p5, F5);
        }
    }
    #pragma unroll
    for (int l = 0; l < NL; ++l) {
        if (NL > 1 && !((lmask >> l) & 1u)) {
            tot[l] = 0;
            continue;
        }
        const double2 a01 = LF2(l, L_AC0), a23 = LF2(l, L_AC0 + 2), p01 = LF2(l, L_PC0), p23 = LF2(l, L_PC0 + 2);
        const double y2 = LF(l, L_Y2);
        double acc = LF(l, L_FAR) * fma(-y2, fma(-y2, F5, F3), F1);
        acc = fma(LF(l, L_BC0), MW[0], acc), acc = fma(LF(l, L_BC0 + 1), MW[1], acc), acc = fma(LF(l, L_BC0 + 2), MW[2], acc);
        acc = fma(a01.x, MG[0], acc), acc = fma(a01.y, MG[1], acc), acc = fma(a23.x, MG[2], acc), acc = fma(a23.y, MG[3], acc);
        acc = fma(p01.x, MU[0], acc), acc = fma(p01.y, MU[1], acc), acc = fma(p23.x, MU[2], acc), acc = fma(p23.y, MU[3], acc);
        tot[l] = acc;
    }
}

// ---- FP32 fast path (FSB_PRECISION_FP32: flux within 1e-5 of the reference) ------------------------
// Same decomposition in single precision: the pixel coordinate xb is formed in FP64 (velocities of
// thousands of km/s against 1e-5 accuracy), everything per node runs in FP32: degree-3 table pieces
// (one 16-byte load per node), quadratics for A and Pe, one __expf per node for the Gaussian.
__device__ __forceinline__ float exp2_ftz(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <int NL>
__device__ __forceinline__ void node_sum_near32(float xb, float step, const float *__restrict__ fl,
                                                const float4 *__restrict__ tab32, bool gauss, unsigned lmask,
                                                float (&tot)[NL])
{
    // node by node, accumulating into the lines at once: nothing but the line coefficients and the accumulators
    // lives across nodes (per-node arrays were spilled to local memory in this instantiation; measured 1.4 % faster)
    float a0[NL], a1[NL], a2[NL], p0[NL], p1[NL], p2[NL], acc[NL];
    #pragma unroll
    for (int l = 0; l < NL; ++l) {
        a0[l] = FLF(l, FL_A0), a1[l] = FLF(l, FL_A0 + 1), a2[l] = FLF(l, FL_A0 + 2);
        p0[l] = FLF(l, FL_PE0), p1[l] = FLF(l, FL_PE0 + 1), p2[l] = FLF(l, FL_PE0 + 2);
        acc[l] = fmaf(fmaf(fmaf(fmaf(FLF(l, FL_BQ0 + 4), xb, FLF(l, FL_BQ0 + 3)), xb, FLF(l, FL_BQ0 + 2)), xb, FLF(l, FL_BQ0 + 1)),
                      xb, FLF(l, FL_BQ0));
    }
    #pragma unroll
    for (int i = 0; i < 7; ++i) {
        const float x = fmaf((float) (i + 1), step, xb), ax = fabsf(x);
        const float m = fmaf(ax, (float) FSB_GTAB_INV_DELTA, 12582912.0f);  // 1.5 * 2^23: low bits = rint(8|x|)
        const int k = min(__float_as_int(m) & 0x3fffff, FSB_GTAB_NINT - 1);
        const float t = fmaf(m - 12582912.0f, -1.0f / (float) FSB_GTAB_INV_DELTA, ax);
        const float4 c = tab32[k];
        const float kw = FS(F_KW0 + i);
        const float g = fmaf(fmaf(fmaf(c.w, t, c.z), t, c.y), t, c.x) * kw;
        const float s = x * x;
        if (gauss) {  // one special-function-unit exponential per node: no recurrence to overflow in FP32
            const float U = exp2_ftz(s * -1.4426950408889634f) * kw;  // exp(-s); below 2^-126 flushes to 0
            #pragma unroll
            for (int l = 0; l < NL; ++l) {
                const float A = fmaf(fmaf(a2[l], s, a1[l]), s, a0[l]);
                const float Pe = fmaf(fmaf(p2[l], s, p1[l]), s, p0[l]);
                acc[l] = fmaf(U, Pe, fmaf(g, A, acc[l]));
            }
        } else {
            #pragma unroll
            for (int l = 0; l < NL; ++l) acc[l] = fmaf(g, fmaf(fmaf(a2[l], s, a1[l]), s, a0[l]), acc[l]);
        }
    }
    #pragma unroll
    for (int l = 0; l < NL; ++l) tot[l] = (NL > 1 && !((lmask >> l) & 1u)) ? 0.f : acc[l];
}

template <int NL>
__device__ __forceinline__ void node_sum_far32(float xb, float step, const float *__restrict__ fl, unsigned lmask,
                                               float (&tot)[NL])
{
    float u[7], p1[7], p3[7];
    #pragma unroll
    for (int i = 0; i < 7; ++i) {
        const float x = fmaf((float) (i + 1), step, xb);
        u[i] = __frcp_rn(x * x);
        p1[i] = fmaf(fmaf(fmaf(13.125f, u[i], 3.75f), u[i], 1.5f), u[i], 1.0f);
        p3[i] = fmaf(5.0f, u[i], 1.0f);
    }
    #pragma unroll
    for (int l = 0; l < NL; ++l) {
        if (NL > 1 && !((lmask >> l) & 1u)) {
            tot[l] = 0;
            continue;
        }
        const float y2 = FLF(l, FL_Y2);
        float acc = 0;
        #pragma unroll
        for (int i = 0; i < 7; ++i) acc = fmaf(u[i] * fmaf(-y2 * u[i], p3[i], p1[i]), FS(F_KW0 + i), acc);
        tot[l] = FLF(l, FL_YISP) * acc;
    }
}

// Generic per-node evaluation at velocity offset vouter for ONE line (mixed near/far warp steps and
// sub-sampled pixels).  Returns the optical depth (amplitude included).
__device__ __noinline__ double node_sum_generic(double vouter, const double *__restrict__ sl, int l,
                                                const double2 *__restrict__ tabA)
{
    const double2 a01 = LF2(l, L_AC0), a23 = LF2(l, L_AC0 + 2), p01 = LF2(l, L_PC0), p23 = LF2(l, L_PC0 + 2);
    const double b0 = LF(l, L_BC0), b1 = LF(l, L_BC0 + 1), b2 = LF(l, L_BC0 + 2);
    const double xu2 = SF(S_XU2), y = LF(l, L_Y), cd = LF(l, L_CD);
    const double xb = fma(-vouter, SF(S_INVB), SF(S_XOFF)), step = SF(S_STEP);
    double total = 0;
    #pragma unroll 1
    for (int i = 0; i < 7; ++i) {
        const double x = fma((double) (i + 1), step, xb), ax = fabs(x), s = x * x;
        double hval;
        if (ax >= kFarXMin) {
            hval = cd * voigt_far(s, y);
        } else {
            const double U = s < xu2 ? fast_exp(-s) : 0.0;
            const double G = g_table(ax, tabA);
            const double Pe = fma(fma(fma(p23.y, s, p23.x), s, p01.y), s, p01.x);
            const double A = fma(fma(fma(a23.y, s, a23.x), s, a01.y), s, a01.x);
            const double B = fma(fma(b2, s, b1), s, b0);
            hval = fma(U, Pe, fma(G, A, B));
        }
        total = fma(hval, SF(S_KW0 + i), total);
    }
    return total;
}

__device__ __noinline__ double node_sum_exact(double vouter, const double *__restrict__ sl, int l)
{
    const double xb = fma(-vouter, SF(S_INVB), SF(S_XOFF)), step = SF(S_STEP);
    const double y = LF(l, L_Y), erfcx_y = LF(l, L_ERFCX);
    double total = 0;
    #pragma unroll 1
    for (int i = 0; i < 7; ++i) total += voigt_exact(fma((double) (i + 1), step, xb), y, erfcx_y) * SF(S_KW0 + i);
    return LF(l, L_CD) * total;
}

// Pixel average tau_kern_outer (singleabs.h:104-126) for one line through the generic / exact
// evaluators; returns the optical depth of the pixel and the number of inner sums.
template <bool EXACT>
__device__ __noinline__ double pixel_sum_slow(double vlow, double vhigh_px, const double *__restrict__ sl, int l,
                                              const double2 *__restrict__ tabA, int &ninner)
{
    const double width = vhigh_px - vlow;
    if (width < SF(S_HALFB)) {
        ninner = 1;
        const double vmid = (vhigh_px + vlow) / 2.;
        return EXACT ? node_sum_exact(vmid, sl, l) : node_sum_generic(vmid, sl, l, tabA);
    }
    const int npoints = (int) (2 * ceil(width / SF(S_HALFB) / 2) + 1.);
    const double dv = width / (npoints - 1);
    double total = 0;
    for (int i = 0; i < npoints; ++i) {
        const double v = (i == 0) ? vlow : ((i == npoints - 1) ? vhigh_px : i * dv + vlow);
        const double wgt = (i == 0 || i == npoints - 1) ? 0.5 : 1.0;
        total += wgt * (EXACT ? node_sum_exact(v, sl, l) : node_sum_generic(v, sl, l, tabA));
    }
    ninner = npoints;
    return total / (npoints - 1);
}

struct Tally {
    unsigned pix = 0, inner = 0, iter = 0;
    unsigned route[5] = {0, 0, 0, 0, 0};  // near+U, near, far, mixed, slow
};

// ---- outward pixel march of one particle (absorption.cpp:250-278) for NL fused lines --------------
// Lane layout, fixed for the whole march: lanes 0-15 serve the upward run (z = zmax + o), lanes 16-31
// the downward run (z = zmax - 1 - o), o = base + (lane & 15); both runs advance 16 pixels per step.
// The profile is symmetric about the particle, so the two runs end within a step of each other and a
// finished run idles its half warp for at most that step.
// live[l]: bit 0 = the upward run of line l is still going, bit 1 = the downward run.
struct MarchGeom {
    int dir, sub, half, zmax, j0;
    unsigned grp_lt, up_lanes, dn_lanes;
};

__device__ __forceinline__ MarchGeom march_geom(int lane, int nbins, int2 zj)
{
    MarchGeom g;
    g.dir = lane >> 4;
    g.sub = lane & 15;
    g.half = nbins / 2;
    g.zmax = zj.x;
    g.j0 = zj.y;
    g.up_lanes = 0x0000ffffu;
    g.dn_lanes = 0xffff0000u;
    g.grp_lt = ((1u << lane) - 1u) & (g.dir ? g.dn_lanes : g.up_lanes);
    return g;
}

// add, then stop each run at its first pixel below tautail (absorption.cpp:260-263,274-277)
template <int NL, bool COUNT>
__device__ __forceinline__ void march_commit(const MarchGeom &g, const double (&t)[NL], const double (&cur)[NL], unsigned (&live)[NL],
                                             bool mine, int base, int j, int ninner, double *__restrict__ row0, int64_t line_stride,
                                             double tautail, Tally &tally)
{
    const bool done = base + 16 >= g.half;
    #pragma unroll
    for (int l = 0; l < NL; ++l) {
        const bool on = mine && ((live[l] >> g.dir) & 1u);
        const unsigned stop = __ballot_sync(kFull, on && (t[l] < tautail));
        if (on && !(stop & g.grp_lt)) {  // no lane of my run below me has stopped
            row0[l * line_stride + j] = cur[l] + t[l];
            if (COUNT) {
                ++tally.pix;
                tally.inner += ninner;
            }
        }
        if (done) live[l] = 0;
        else live[l] &= ~(((stop & g.up_lanes) ? 1u : 0u) | ((stop & g.dn_lanes) ? 2u : 0u));
    }
    if (COUNT) ++tally.iter;
}

// Fast routes.  Which route a step takes follows from integer pixel thresholds computed once per
// particle (setup_particle): x is affine in the outward pixel index, so "all nodes inside the table",
// "all nodes on the wing series" and "out of reach of the Gaussian" are index ranges per direction.
//
// Control state is kept in a handful of warp-uniform integers so a march step costs few instructions
// beyond the quadrature: `live` has bit (2 l + d) set while direction d (0 up, 1 down) of line l is going;
// near_lim / far_beg / gauss_end are the route limits of the directions still live and are refreshed only
// when a direction ends.
// Predicated 8-byte store: no divergent region (BSSY / BSYNC) around one instruction.
__device__ __forceinline__ void store_if(double *p, double v, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.f64 [%0], %1;\n\t}" ::"l"(p), "d"(v), "r"((unsigned) pred) : "memory");
}

template <int NL, bool COUNT, bool F32>
__device__ __forceinline__ void march_fast(const double *__restrict__ sl, const float *__restrict__ fl, const double2 *__restrict__ tab,
                                           const float4 *__restrict__ tab32, double *__restrict__ row0, int64_t line_stride, int nbins,
                                           double tautail, int lane, Tally &tally)
{
    constexpr unsigned kUpBits = NL == 2 ? 0x5u : 0x1u, kDnBits = NL == 2 ? 0xau : 0x2u;
    const int dir = lane >> 4, sub = lane & 15, half = nbins >> 1;
    const unsigned grp_lt = ((1u << lane) - 1u) & (dir ? 0xffff0000u : 0x0000ffffu);
    const double2 sx = SF2(S_STEP), pz = SF2(S_PIX), nf = SF2(S_THR_N), gr = SF2(S_THR_G);
    const double step = sx.x, xb0 = sx.y, pix = pz.x;
    const int2 zj = make_int2(__double2loint(pz.y), __double2hiint(pz.y));
    const int2 thrN = make_int2(__double2loint(nf.x), __double2hiint(nf.x));
    const int2 thrF = make_int2(__double2loint(nf.y), __double2hiint(nf.y));
    const int2 thrG = make_int2(__double2loint(gr.x), __double2hiint(gr.x));
    const bool rec_ok = __double2loint(gr.y) != 0, cubic = __double2hiint(gr.y) != 0;
    unsigned live = half > 0 ? (kUpBits | kDnBits) : 0u;
    int near_lim = min(thrN.x, thrN.y), far_beg = max(thrF.x, thrF.y), gauss_end = max(thrG.x, thrG.y);
    double U0 = 0, R = 0, rho = 0;  // Gaussian recurrence state of this lane
    bool rec_valid = false;
    for (int base = 0; live; base += 16) {
        const int o = base + sub;           // outward pixel index
        const int dz = dir ? ~o : o;        // z - zmax  (~o = -1 - o)
        int j = zj.y + dz;                  // z mod nbins: |z - zmax| <= nbins/2
        j += j < 0 ? nbins : (j >= nbins ? -nbins : 0);
        const unsigned mybits = o < half ? (live >> dir) & kUpBits : 0u;  // bit 2 l: my run of line l is going
        // start the read of the output pixels now; they are consumed after the quadrature (reading them just
        // before the add instead was measured 2 % slower even with the position-ordered traversal)
        double *const pj = row0 + j;
        double cur[NL];
        #pragma unroll
        for (int l = 0; l < NL; ++l) cur[l] = (mybits >> (2 * l)) & 1u ? __ldcg(pj + l * line_stride) : 0.0;
        unsigned lmask = 0;  // lines with a live run
        #pragma unroll
        for (int l = 0; l < NL; ++l) lmask |= (live >> (2 * l)) & 3u ? (1u << l) : 0u;
        // node 0 of pixel z in units of btherm: affine in the pixel offset from zmax (the reference's
        // (vlow + vhigh)/2 differs from this by rounding only, < 1e-13 in x)
        const double xb = fma((double) dz, -pix, xb0);
        const bool gauss = base < gauss_end;
        double tot[NL];
        #pragma unroll
        for (int l = 0; l < NL; ++l) tot[l] = 0;
        if (base >= far_beg) {
            if (F32) {
                float tf[NL];
                node_sum_far32<NL>((float) xb, FS(F_STEP), fl, lmask, tf);
                #pragma unroll
                for (int l = 0; l < NL; ++l) tot[l] = LF(l, L_CD) * (double) tf[l];
            } else {
                node_sum_far<NL>(xb, step, sl, lmask, tot);
            }
            rec_valid = false;
            if (COUNT) ++tally.route[2];
        } else if (min(base + 16, half) <= near_lim) {
            if (F32) {
                float tf[NL];
                node_sum_near32<NL>((float) xb, FS(F_STEP), fl, tab32, gauss, lmask, tf);
                #pragma unroll
                for (int l = 0; l < NL; ++l) tot[l] = LF(l, L_CD) * (double) tf[l];
            } else {
                const double2 qk = SF2(S_Q);  // {q, K16}
                if (gauss) {
                    if (rec_valid) {  // one march step outward: x -> x -+ 16 pixels
                        const double2 lud = SF2(S_LU16);
                        U0 *= rho;
                        rho *= qk.y;
                        R *= dir ? lud.y : lud.x;
                    } else {
                        const double x1 = xb + step;
                        U0 = fast_exp(-x1 * x1);
                        R = fast_exp(-fma(2.0, x1, step) * step);
                        // the march-step recurrence only pays when another step within reach of the Gaussian follows
                        // (weak metal lines end inside their first step: one exponential less per particle)
                        rec_valid = rec_ok && base + 16 < gauss_end;
                        if (rec_valid) {
                            const double delta = (dir ? 16.0 : -16.0) * pix;
                            rho = fast_exp(-fma(2.0, x1, delta) * delta);
                        }
                    }
                } else {
                    rec_valid = false;
                }
                node_sum_near<NL>(xb, step, sl, tab, U0, R, qk.x, gauss, cubic, lmask, tot);
            }
            if (COUNT) ++tally.route[gauss ? 0 : 1];
        } else {
            rec_valid = false;
            if (!F32) {
                // transition step, node by node with a warp-uniform route per node (node_sum_mixed)
                node_sum_mixed<NL>(xb, step, sl, tab, dir != 0, mybits != 0, lmask, tot);
            } else {
                // transition step: lanes differ.  Lane class: 0 table, 1 series, 2 neither (own nodes on both
                // sides of the overlap: node by node), 3 idle.
                const int myN = dir ? thrN.y : thrN.x, myF = dir ? thrF.y : thrF.x;
                const int lc = !mybits ? 3 : (o >= myF ? 1 : (o < myN ? 0 : 2));
                const unsigned cls = __reduce_or_sync(kFull, 1u << lc);
                rec_valid = false;
                if (cls & 1u) {
                    if (F32) {
                        float tf[NL];
                        node_sum_near32<NL>((float) xb, FS(F_STEP), fl, tab32, gauss, lmask, tf);
                        #pragma unroll
                        for (int l = 0; l < NL; ++l) tot[l] = LF(l, L_CD) * (double) tf[l];
                    } else {
                        if (gauss) {
                            const double x1 = xb + step;
                            U0 = fast_exp(-x1 * x1);
                            R = fast_exp(-fma(2.0, x1, step) * step);
                        }
                        node_sum_near<NL>(xb, step, sl, tab, U0, R, SF(S_Q), gauss, cubic, lmask, tot);
                    }
                }
                if (cls & 2u) {
                    double tfar[NL];
                    if (F32) {
                        float tf[NL];
                        node_sum_far32<NL>((float) xb, FS(F_STEP), fl, lmask, tf);
                        #pragma unroll
                        for (int l = 0; l < NL; ++l) tfar[l] = LF(l, L_CD) * (double) tf[l];
                    } else {
                        node_sum_far<NL>(xb, step, sl, lmask, tfar);
                    }
                    #pragma unroll
                    for (int l = 0; l < NL; ++l) tot[l] = lc == 1 ? tfar[l] : tot[l];
                }
                if (cls & 4u) {
                    if (lc == 2) {
                        #pragma unroll
                        for (int l = 0; l < NL; ++l)
                            if ((lmask >> l) & 1u) tot[l] = node_sum_generic((SF(S_XOFF) - xb) / SF(S_INVB), sl, l, tab);
                    }
                }
            }
            if (COUNT) ++tally.route[3];
        }
        // add, then stop each run at its first pixel below tautail (absorption.cpp:260-263,274-277)
        unsigned ended = 0;
        #pragma unroll
        for (int l = 0; l < NL; ++l) {
            const double t = tot[l];  // already scaled by the line amplitude
            const bool on = (mybits >> (2 * l)) & 1u;
            const unsigned stop = __ballot_sync(kFull, on && (t < tautail));
            const bool wr = on && !(stop & grp_lt);  // no lane of my run below me has stopped
            store_if(pj + l * line_stride, cur[l] + t, wr);
            if (COUNT && wr) {
                ++tally.pix;
                ++tally.inner;
            }
            ended |= ((stop & 0xffffu ? 1u : 0u) | (stop >> 16 ? 2u : 0u)) << (2 * l);
        }
        if (COUNT) ++tally.iter;
        if (base + 16 >= half) ended = ~0u;
        if (ended & live) {
            live &= ~ended;
            const bool up = live & kUpBits, dn = live & kDnBits;
            near_lim = min(up ? thrN.x : 0x7fffffff, dn ? thrN.y : 0x7fffffff);
            far_beg = max(up ? thrF.x : 0, dn ? thrF.y : 0);
            gauss_end = max(up ? thrG.x : 0, dn ? thrG.y : 0);
        }
    }
}

// Pixels at least btherm/2 wide (coarse spectra): tau_kern_outer's trapezoid over npoints inner positions
// (singleabs.h:110-125), every inner position a 7-node sum by the fast routes.  Same lane layout and stop rule
// as march_fast; the Gaussians are started afresh at every inner position (no recurrence across positions).
// The route thresholds of these particles carry a half-pixel margin (setup_particle).
template <int NL, bool COUNT>
__device__ __noinline__ void march_sub(const double *__restrict__ sl, const double2 *__restrict__ tab, double *__restrict__ row0,
                                       int64_t line_stride, int nbins, double bintov, double tautail, int lane, Tally &tally)
{
    constexpr unsigned kUpBits = NL == 2 ? 0x5u : 0x1u, kDnBits = NL == 2 ? 0xau : 0x2u;
    const int dir = lane >> 4, sub = lane & 15, half = nbins >> 1;
    const unsigned grp_lt = ((1u << lane) - 1u) & (dir ? 0xffff0000u : 0x0000ffffu);
    const double2 sx = SF2(S_STEP), pz = SF2(S_PIX), nf = SF2(S_THR_N), gr = SF2(S_THR_G);
    const double step = sx.x;
    const int2 zj = make_int2(__double2loint(pz.y), __double2hiint(pz.y));
    const int2 thrN = make_int2(__double2loint(nf.x), __double2hiint(nf.x));
    const int2 thrF = make_int2(__double2loint(nf.y), __double2hiint(nf.y));
    const int2 thrG = make_int2(__double2loint(gr.x), __double2hiint(gr.x));
    const int npts = (int) SF(S_PAD);
    const double vel = SF(S_VEL), inv_b = SF(S_INVB), xoff = SF(S_XOFF), q = SF(S_Q);
    unsigned live = half > 0 ? (kUpBits | kDnBits) : 0u;
    int near_lim = min(thrN.x, thrN.y), far_beg = max(thrF.x, thrF.y), gauss_end = max(thrG.x, thrG.y);
    for (int base = 0; live; base += 16) {
        const int o = base + sub;
        const int dz = dir ? ~o : o;
        int j = zj.y + dz;
        j += j < 0 ? nbins : (j >= nbins ? -nbins : 0);
        const unsigned mybits = o < half ? (live >> dir) & kUpBits : 0u;
        double *const pj = row0 + j;
        double cur[NL];
        #pragma unroll
        for (int l = 0; l < NL; ++l) cur[l] = (mybits >> (2 * l)) & 1u ? __ldcg(pj + l * line_stride) : 0.0;
        unsigned lmask = 0;
        #pragma unroll
        for (int l = 0; l < NL; ++l) lmask |= (live >> (2 * l)) & 3u ? (1u << l) : 0u;
        // the pixel's velocity interval exactly as absorption.cpp:252-255,268-271
        const double vlow = __dsub_rn(__dmul_rn((double) (zj.x + dz), bintov), vel);
        const double vhigh_px = __dadd_rn(vlow, bintov);
        const double dv = (vhigh_px - vlow) / (npts - 1);
        const bool gauss = base < gauss_end;
        const bool all_far = base >= far_beg, all_near = min(base + 16, half) <= near_lim;
        double acc[NL];
        #pragma unroll
        for (int l = 0; l < NL; ++l) acc[l] = 0;
        for (int i = 0; i < npts; ++i) {
            const double v = i == 0 ? vlow : (i == npts - 1 ? vhigh_px : fma((double) i, dv, vlow));
            const double wgt = (i == 0 || i == npts - 1) ? 0.5 : 1.0;
            const double xb = fma(-v, inv_b, xoff);
            double tot[NL];
            #pragma unroll
            for (int l = 0; l < NL; ++l) tot[l] = 0;
            if (all_far) {
                node_sum_far<NL>(xb, step, sl, lmask, tot);
            } else if (all_near) {
                double U0 = 0, R = 0;
                if (gauss) {
                    const double x1 = xb + step;
                    U0 = fast_exp(-x1 * x1);
                    R = fast_exp(-fma(2.0, x1, step) * step);
                }
                node_sum_near<NL>(xb, step, sl, tab, U0, R, q, gauss, true, lmask, tot);
            } else {
                node_sum_mixed<NL>(xb, step, sl, tab, dir != 0, mybits != 0, lmask, tot);
            }
            #pragma unroll
            for (int l = 0; l < NL; ++l) acc[l] = fma(wgt, tot[l], acc[l]);
        }
        if (COUNT) ++tally.route[all_far ? 2 : (all_near ? (gauss ? 0 : 1) : 3)];
        unsigned ended = 0;
        #pragma unroll
        for (int l = 0; l < NL; ++l) {
            const double t = acc[l] / (npts - 1);
            const bool on = (mybits >> (2 * l)) & 1u;
            const unsigned stop = __ballot_sync(kFull, on && (t < tautail));
            const bool wr = on && !(stop & grp_lt);
            store_if(pj + l * line_stride, cur[l] + t, wr);
            if (COUNT && wr) {
                ++tally.pix;
                tally.inner += npts;
            }
            ended |= ((stop & 0xffffu ? 1u : 0u) | (stop >> 16 ? 2u : 0u)) << (2 * l);
        }
        if (COUNT) ++tally.iter;
        if (base + 16 >= half) ended = ~0u;
        if (ended & live) {
            live &= ~ended;
            const bool up = live & kUpBits, dn = live & kDnBits;
            near_lim = min(up ? thrN.x : 0x7fffffff, dn ? thrN.y : 0x7fffffff);
            far_beg = max(up ? thrF.x : 0, dn ? thrF.y : 0);
            gauss_end = max(up ? thrG.x : 0, dn ? thrG.y : 0);
        }
    }
}

// Slow routes: the exact Faddeeva restatement, and pixels wider than btherm/2 (sub-sampling rule of
// singleabs.h:110-125).  Same lane layout, per-pixel evaluation through pixel_sum_slow.
template <int NL, bool EXACT, bool COUNT>
__device__ __noinline__ void march_slow(const double *__restrict__ sl, const double2 *__restrict__ tab, double *__restrict__ row0,
                                        int64_t line_stride, int nbins, double bintov, double tautail, int lane, Tally &tally)
{
    const MarchGeom g = march_geom(lane, nbins, *reinterpret_cast<const int2 *>(&SF(S_ZMAX)));
    const double vel = SF(S_VEL);
    unsigned live[NL];
    #pragma unroll
    for (int l = 0; l < NL; ++l) live[l] = g.half > 0 ? 3u : 0u;
    for (int base = 0;; base += 16) {
        unsigned any = 0;
        #pragma unroll
        for (int l = 0; l < NL; ++l) any |= live[l];
        if (!any) break;
        const int o = base + g.sub;
        const bool mine = o < g.half;
        const int dz = g.dir ? -1 - o : o;
        const int z = g.zmax + dz;
        int j = g.j0 + dz;
        j += j < 0 ? nbins : (j >= nbins ? -nbins : 0);
        double cur[NL], t[NL];
        int ninner = 1;
        const double vlow = __dsub_rn(__dmul_rn((double) z, bintov), vel);
        const double vhigh_px = __dadd_rn(vlow, bintov);
        #pragma unroll
        for (int l = 0; l < NL; ++l) {
            const bool on = mine && ((live[l] >> g.dir) & 1u);
            cur[l] = on ? __ldcg(row0 + l * line_stride + j) : 0.0;
            t[l] = on ? pixel_sum_slow<EXACT>(vlow, vhigh_px, sl, l, tab, ninner) : 0.0;
        }
        if (COUNT) ++tally.route[4];
        march_commit<NL, COUNT>(g, t, cur, live, mine, base, j, ninner, row0, line_stride, tautail, tally);
    }
}

// Everything but the plain fast march behind ONE call site, so the register allocation of the kernel's main
// loop sees a single cold call (measured: separate call sites cost the main path 1.4 %).
template <int NL, bool COUNT>
__device__ __noinline__ void march_other(int mode, const double *__restrict__ sl, const double2 *__restrict__ tab,
                                         double *__restrict__ row0, int64_t line_stride, int nbins, double bintov,
                                         double tautail, int lane, Tally &tally)
{
    if (mode == 4) march_sub<NL, COUNT>(sl, tab, row0, line_stride, nbins, bintov, tautail, lane, tally);
    else if (mode == 2) march_slow<NL, true, COUNT>(sl, tab, row0, line_stride, nbins, bintov, tautail, lane, tally);
    else march_slow<NL, false, COUNT>(sl, tab, row0, line_stride, nbins, bintov, tautail, lane, tally);
}

__device__ __forceinline__ int clamp_index(double v) { return (int) fmin(fmax(v, 0.0), 1073741824.0); }

// Per-particle constants, one particle per lane (absorption.cpp:218-246, singleabs.h:81-90).
template <int KERNEL, int NL, bool F32>
__device__ __noinline__ void setup_particle(const InterpConsts &C, double *__restrict__ sl, float *__restrict__ fl, int64_t k, int ax,
                                               const int32_t *__restrict__ zorder, const int32_t *__restrict__ particle, const double *__restrict__ dr2s,
                                               const float *__restrict__ pos, const float *__restrict__ vel,
                                               const float *__restrict__ dens, const float *__restrict__ temp,
                                               const float *__restrict__ hsml, const float *__restrict__ cells)
{
    k = zorder[k];  // traversal order along the sightline (fsb_index.cu), not list order
    const int64_t ip = particle[k];
    const float ppos = pos[3 * ip + ax], pvel = vel[3 * ip + ax];
    const float pdens = dens[ip], ptemp = temp[ip];
    double dr2;
    float smooth;
    if (KERNEL == FSB_KERNEL_VORONOI) {
        dr2 = (double) cells[2 * k];
        smooth = cells[2 * k + 1];
    } else {
        dr2 = dr2s[k];
        smooth = hsml[ip];
    }
    double pos1 = (double) ppos;
    int mode = 1;
    if (KERNEL == FSB_KERNEL_VORONOI) {
        const double lim = 2 * C.vbox / C.velfac;
        if (dr2 > lim || (double) smooth > lim) mode = 0;
        pos1 = __dmul_rn(__dadd_rn(dr2, (double) smooth), 0.5);
    } else {
        if (__dsub_rn((double) __fmul_rn(smooth, smooth), dr2) <= 0) mode = 0;
    }
    const double btherm = C.bfac * sqrt((double) ptemp);
    const double velp = __dadd_rn(__dmul_rn(C.velfac, pos1), (double) pvel);
    double vdr2 = C.velfac * dr2;
    if (KERNEL != FSB_KERNEL_VORONOI) vdr2 *= C.velfac;
    const double vsmooth = C.velfac * (double) smooth;
    const double inv_b = 1.0 / btherm;
    double vhigh = (vsmooth * vsmooth > vdr2) ? sqrt(vsmooth * vsmooth - vdr2) : 0;
    if (KERNEL == FSB_KERNEL_VORONOI) vhigh = (vdr2 > 0 && vsmooth > 0) ? (vsmooth - vdr2) / 2. : 0;
    const double deltav = 2. * vhigh / kNGrid;
    const double step = deltav * inv_b;
    const bool force_exact = C.voigt == FSB_VOIGT_EXACT;
    SF(S_VEL) = velp;
    SF(S_INVB) = inv_b;
    SF(S_HALFB) = btherm / 2.;
    SF(S_STEP) = step;
    SF(S_XOFF) = -vhigh * inv_b;
    SF(S_Q) = fast_exp(-2.0 * step * step);
    // kernel weights and their moments about xb, in units of btherm: M_n = sum kw_i (i step)^n.  A rolled loop on
    // purpose: this function runs once per batch of particles and its length is paid in instruction fetches.
    double M[5] = {0, 0, 0, 0, 0};
    const double inv_vs = 1.0 / vsmooth;
    #pragma unroll 1
    for (int i = 1; i < kNGrid; ++i) {
        const double vv = i * deltav - vhigh;
        const double kwi = sph_kernel<KERNEL>(sqrt(vdr2 + vv * vv) * inv_vs) * deltav;
        SF(S_KW0 + i - 1) = kwi;
        if (F32) FS(F_KW0 + i - 1) = (float) kwi;
        const double d = i * step;
        double p = kwi;
        #pragma unroll
        for (int n = 0; n < 5; ++n) {
            M[n] += p;
            p *= d;
        }
    }
    if (F32) FS(F_STEP) = (float) step;
    const double zmaxd = floor(velp / C.bintov);
    {
        int2 zj;
        zj.x = (int) zmaxd;
        zj.y = wrap_bin(zj.x, C.nbins);
        *reinterpret_cast<int2 *>(&SF(S_ZMAX)) = zj;
    }
    const double xb0 = fma(-(fma(zmaxd + 0.5, C.bintov, -velp)), inv_b, -vhigh * inv_b);
    const double pix = C.bintov * inv_b;
    SF(S_XB0) = xb0;
    SF(S_PIX) = pix;
    // march-step recurrence factors: D = 16 pixels in units of btherm
    const double D16 = 16.0 * pix;
    const bool rec_usable = step <= 1.0 && D16 <= 10.0;
    if (rec_usable) {
        SF(S_K16) = fast_exp(-2.0 * D16 * D16);
        SF(S_LU16) = fast_exp(2.0 * D16 * step);
        SF(S_LD16) = fast_exp(-2.0 * D16 * step);
    }
    double ymin = 1e300, ymax = 0;
    #pragma unroll
    for (int l = 0; l < NL; ++l) {
        const double aa = C.line[l].voigt_fac * inv_b;
        const double amp = C.line[l].sigma_a / kSqrtPi * (kLight / 1e5 * inv_b);
        if (mode && (force_exact || !fast_domain(aa))) mode = 2;
        ymin = fmin(ymin, aa);
        ymax = fmax(ymax, aa);
        const double cd = amp * (double) pdens / C.velfac;
        LF(l, L_CD) = cd;
        LF(l, L_Y) = aa;
        LF(l, L_Y2) = aa * aa;
        LF(l, L_FAR) = cd * (0.56418958354775628694807945156 * aa);
        FastCoef fc;
        fast_coefs(aa, fc);
        #pragma unroll
        for (int i = 0; i < 4; ++i) LF(l, L_PC0 + i) = cd * fc.pe[i];
        #pragma unroll
        for (int i = 0; i < 4; ++i) LF(l, L_AC0 + i) = cd * fc.a[i];
        #pragma unroll
        for (int i = 0; i < 3; ++i) LF(l, L_BC0 + i) = cd * fc.b[i];
        // sum_i kw_i B((xb + d_i)^2), B(s) = b0 + b1 s + b2 s^2, as a quartic in xb
        const double bq0 = fma(fc.b[2], M[4], fma(fc.b[1], M[2], fc.b[0] * M[0]));
        const double bq1 = fma(4.0 * fc.b[2], M[3], 2.0 * fc.b[1] * M[1]);
        const double bq2 = fma(6.0 * fc.b[2], M[2], fc.b[1] * M[0]);
        const double bq3 = 4.0 * fc.b[2] * M[1];
        const double bq4 = fc.b[2] * M[0];
        LF(l, L_BQ0) = cd * bq0;
        LF(l, L_BQ0 + 1) = cd * bq1;
        LF(l, L_BQ0 + 2) = cd * bq2;
        LF(l, L_BQ0 + 3) = cd * bq3;
        LF(l, L_BQ0 + 4) = cd * bq4;
        LF(l, L_ERFCX) = 0;
        if (F32) {
            #pragma unroll
            for (int i = 0; i < 3; ++i) FLF(l, FL_A0 + i) = (float) fc.a[i];
            #pragma unroll
            for (int i = 0; i < 3; ++i) FLF(l, FL_PE0 + i) = (float) fc.pe[i];
            FLF(l, FL_BQ0) = (float) bq0;
            FLF(l, FL_BQ0 + 1) = (float) bq1;
            FLF(l, FL_BQ0 + 2) = (float) bq2;
            FLF(l, FL_BQ0 + 3) = (float) bq3;
            FLF(l, FL_BQ0 + 4) = (float) bq4;
            FLF(l, FL_Y2) = (float) (aa * aa);
            FLF(l, FL_YISP) = (float) (aa * 0.56418958354775628694807945156);
        }
    }
    if (mode == 2) {
        #pragma unroll
        for (int l = 0; l < NL; ++l) LF(l, L_ERFCX) = erfcx(C.line[l].voigt_fac * inv_b);
    }
    {
        // every factor of the march-step recurrence stays within e^+-500 while a lane is within reach of the Gaussian core
        int2 rc;
        rc.x = rec_usable ? 1 : 0;
        // degree class: the s^3 terms of A(s) and Pe(s) relative to the profile are at most (4/315) y^6 s^3 on the table
        // route, whose nodes stay below |x| = 12 + the reach of one warp step (and below the table's end)
        const double xs = fmin(FSB_GTAB_XMAX, kFarXMin + D16 + 6.0 * step + 0.5 * pix), y2m = ymax * ymax, s3 = (xs * xs) * y2m;
        rc.y = ((4.0 / 315.0) * s3 * s3 * s3 > 5e-13) ? 1 : 0;
        *reinterpret_cast<int2 *>(&SF(S_RECOK)) = rc;
    }
    const double xu2 = ymin > 0 ? 37.0 - log(ymin) : 1e300;
    SF(S_XU2) = xu2;
    // pixels at least btherm/2 wide are sub-sampled (singleabs.h:110-125; bintov is rounded differently per
    // pixel by at most an ulp): generic per-pixel route
    if (mode == 1 && !(C.bintov * (1 + 1e-12) < btherm / 2.)) mode = 3;
    // ... and served by the fast node sums (march_sub, mode 4) when the inner-point count is the same for every
    // pixel whatever the rounding of its width, and a pixel is narrower than the table/series overlap
    double hm = 0;  // half a pixel in units of btherm: the inner points of a pixel reach that far from its centre
    if (mode == 3) {
        const double r = C.bintov / (btherm / 2.) / 2., rc = ceil(r);
        const bool stable = fabs(r - rint(r)) > 1e-9 * r && C.bintov > (btherm / 2.) * (1 + 1e-9);
#ifndef FSB_NO_MARCH_SUB
        // (measured on the 1-10 km/s sweep: routing every such particle here beats routing only those with narrow
        // kernels; a sightline that alternates between this route and the per-pixel fallback is slower than either)
        if (stable && pix < 6.0 && rc < 1e6) {
            mode = 4;
            hm = 0.5 * pix * (1 + 1e-9) + 1e-9;
            SF(S_PAD) = 2.0 * rc + 1.0;  // npoints, singleabs.h:116
        }
#endif
    }
    SF(S_MODE) = (double) mode;
    // Route thresholds in outward pixels.  Upward run: nodes x_i(o) = X_i - o pix; downward run:
    // x_i(o) = X_i + (1 + o) pix; X_1 = xb0 + step <= X_7 = xb0 + 7 step.  The table covers |x| < 24 and the
    // wing series |x| >= 12; margins of 0.01 dwarf the rounding of these expressions.
    {
        const double X1 = xb0 + step, X7 = fma(7.0, step, xb0), ipix = 1.0 / pix;
        const double lim_n = FSB_GTAB_XMAX - 0.01 - hm, lim_f = kFarXMin + 0.01 + hm, xu = sqrt(fmin(xu2, 1e12)) + hm;
        const bool near_any = X7 < lim_n && X1 > -lim_n;
        int2 tn, tf, tg;
        tn.x = near_any ? clamp_index(floor((X1 + lim_n) * ipix)) : 0;
        tn.y = near_any ? clamp_index(floor((lim_n - X7) * ipix - 1.0)) : 0;
        tf.x = clamp_index(ceil((X7 + lim_f) * ipix));
        tf.y = clamp_index(ceil((lim_f - X1) * ipix - 1.0));
        tg.x = clamp_index(ceil((X7 + xu) * ipix));
        tg.y = clamp_index(ceil((xu - X1) * ipix - 1.0));
        *reinterpret_cast<int2 *>(&SF(S_THR_N)) = tn;
        *reinterpret_cast<int2 *>(&SF(S_THR_F)) = tf;
        *reinterpret_cast<int2 *>(&SF(S_THR_G)) = tg;
    }
}

// The FP32 tables and float slabs exist only in the FP32 instantiation (they would cost the FP64 kernel a
// resident CTA per SM).
constexpr int kTabDoubles = FSB_GTAB_SIZE;  // the staged G(x) table
template <int NL, bool F32> constexpr size_t tau_smem_bytes()
{
    return sizeof(double) * (size_t) (kTabDoubles + kTauWarps * SlabSize<NL, F32>::kDoubles) +
           (F32 ? sizeof(float) * (size_t) (4 * FSB_GTAB_NINT + kTauWarps * FSlabSize<NL>::kFloats) : 0);
}

template <int KERNEL, int NL, bool COUNT, bool F32, bool STREAM>
__global__ void __launch_bounds__(kTauThreads, FSB_TAU_MIN_BLOCKS)
k_tau(InterpConsts C, Items items, int n_items, int *__restrict__ next_item, const int64_t *__restrict__ offsets,
      const int32_t *__restrict__ zorder, const int32_t *__restrict__ particle, const double *__restrict__ dr2s,
      const int32_t *__restrict__ axis,
      const float *__restrict__ pos, const float *__restrict__ vel, const float *__restrict__ dens,
      const float *__restrict__ temp, const float *__restrict__ hsml, const float *__restrict__ cells,
      double *__restrict__ out, double *__restrict__ scratch, int64_t scratch_stride,
      unsigned long long *__restrict__ counters, int *__restrict__ chunk_done, int *host_flags, int chunk_lines, fsb_push push)
{
    extern __shared__ __align__(16) double smem[];
    double2 *tab = reinterpret_cast<double2 *>(smem);  // [kGtabPieces]: three 16-byte pieces per interval
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *slab = smem + kTabDoubles + warp * SlabSize<NL, F32>::kDoubles;
    // floats follow the doubles: the degree-3 table of the FP32 path (16-byte aligned), then the float slabs
    static_assert((kTabDoubles + kTauWarps * SlabSize<NL, F32>::kDoubles) % 2 == 0, "float4 table must stay 16-byte aligned");
    float *tab32f = reinterpret_cast<float *>(smem + kTabDoubles + kTauWarps * SlabSize<NL, F32>::kDoubles);
    const float4 *tab32 = reinterpret_cast<const float4 *>(tab32f);
    float *fslab = F32 ? tab32f + 4 * FSB_GTAB_NINT + warp * FSlabSize<NL>::kFloats : nullptr;
    for (int i = threadIdx.x; i < kGtabPieces; i += kTauThreads) tab[i] = d_gtable[i];
    if (F32)
        for (int i = threadIdx.x; i < 4 * FSB_GTAB_NINT; i += kTauThreads) tab32f[i] = d_gtable32[i];
    __syncthreads();

    const int nbins = C.nbins;
    const int64_t out_stride = (int64_t) C.nlos * nbins;
    Tally tally;
    unsigned long long n_pairs = 0;

    // one item per sightline and a host sink: the last row of a chunk of sightlines to finish raises the
    // chunk's flag in pinned host memory; the host copies that chunk out while the kernel carries on
    auto row_done = [&](int finished_line) {
        __threadfence();
        __syncwarp();
        if (push.npeers > 0) {
            // the finished rows of this sightline (one per fused line) go to every destination array: peers' memory
            // over NVLink, stores only; 8-byte pieces (rows of an odd number of pixels are not 16-byte aligned)
            for (int l = 0; l < NL; ++l) {
                const double *src = out + (int64_t) l * ((int64_t) C.nlos * C.nbins) + (int64_t) finished_line * C.nbins;
                const int64_t doff = (int64_t) l * push.line_stride + (int64_t) finished_line * C.nbins;
                for (int j = lane; j < C.nbins; j += 32) {
                    const double v = __ldcg(src + j);
                    for (int q = 0; q < push.npeers; ++q) push.dest[q][doff + j] = v;
                }
            }
            __threadfence_system();
            if (chunk_done == nullptr) return;
        }
        if (lane == 0) {
            const int c = finished_line / chunk_lines;
            const int in_chunk = min(chunk_lines, C.nlos - c * chunk_lines);
            if (atomicAdd(&chunk_done[c], 1) + 1 == in_chunk) {
                __threadfence_system();
                *reinterpret_cast<volatile int *>(host_flags + c) = 1;
            }
        }
    };
    if (STREAM && items.item_start == nullptr && items.ticket_pairs > 0) {
        // ticketed runs only list sightlines that have candidates: the empty ones are finished here, spread over the warps
        for (int l = C.line0 + blockIdx.x * kTauWarps + warp; l < C.line0 + C.nrange; l += gridDim.x * kTauWarps)
            if (offsets[l + 1] == offsets[l]) row_done(l);
    }
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(next_item, 1);
        item = __shfl_sync(kFull, item, 0);
        if (item >= n_items) break;
        int line;
        int64_t kbeg, kend;
        int run = 0;
        bool last_run = true;
        if (items.item_start == nullptr && items.ticket_pairs > 0) {
            // ticketed run of a line's list: phases by remaining runs (fsb_items.cuh)
            const int nph = items.max_runs;
            if (item >= items.phase_start[nph]) break;
            int lo = 0, hi = nph;  // last phase with phase_start <= item
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (items.phase_start[mid] <= item) lo = mid;
                else hi = mid;
            }
            const int in_phase = items.phase_start[lo + 1] - items.phase_start[lo];
            line = items.order[in_phase - 1 - (item - items.phase_start[lo])];  // lines that start in this phase first
            const int64_t b0 = offsets[line], b1 = offsets[line + 1];
            run = (int) ((b1 - b0 + items.ticket_pairs - 1) / items.ticket_pairs) - (nph - lo);
            kbeg = b0 + (int64_t) run * items.ticket_pairs;
            kend = min(kbeg + (int64_t) items.ticket_pairs, b1);
            last_run = kend >= b1;
            if (kbeg >= b1) {  // the line has fewer runs (or no candidates at all)
                if (STREAM && run == 0) row_done(line);
                continue;
            }
            if (run > 0) {  // the previous run of this line must have finished adding to the row
                if (lane == 0) {
                    const volatile int *flag = items.line_done + line;
                    while (*flag != run) __nanosleep(64);
                }
                __syncwarp();
                __threadfence();
            }
        } else if (items.item_start == nullptr) {  // one item per sightline of the range
            if (item >= C.nrange) continue;
            line = C.line0 + item;
            kbeg = offsets[line];
            kend = offsets[line + 1];
            if (kend <= kbeg) {
                if (STREAM) row_done(line);
                continue;
            }
        } else if (!locate_item(items, offsets, C.nlos, item, line, kbeg, kend)) {
            continue;
        }
        double *row0 = items.item_start ? scratch + (int64_t) item * nbins : out + (int64_t) line * nbins;
        const int64_t line_stride = items.item_start ? scratch_stride : out_stride;
        const int ax = axis[line] - 1;
        n_pairs += (unsigned long long) (kend - kbeg);

        constexpr int kB = BatchOf<NL, F32>::value;
        for (int64_t k0 = kbeg; k0 < kend; k0 += kB) {
            const int nb = (int) min((int64_t) kB, kend - k0);
            __syncwarp();
            if (lane < nb) setup_particle<KERNEL, NL, F32>(C, slab + lane * SlabSize<NL>::kStride, fslab + lane * FSlabSize<NL>::kStride, k0 + lane, ax, zorder, particle, dr2s, pos, vel, dens, temp, hsml, cells);
            __syncwarp();
            // plain particles first, in list order; the rare others (exact Voigt, coarse pixels) of the batch after
            // them, so that the main loop holds no call (its register allocation is what the step time hangs on).
            // The order is fixed by the data, hence still deterministic.
            unsigned other = 0;
            for (int b = 0; b < nb; ++b) {
                const double *sl = slab + b * SlabSize<NL>::kStride;
                const int mode = (int) SF(S_MODE);
                if (mode == 0) continue;
                if (mode != 1) {
                    other |= 1u << b;
                    continue;
                }
                march_fast<NL, COUNT, F32>(sl, fslab + b * FSlabSize<NL>::kStride, tab, tab32, row0, line_stride, nbins, C.tautail, lane, tally);
                __syncwarp();
            }
            while (other) {
                const int b = __ffs(other) - 1;
                other &= other - 1;
                const double *sl = slab + b * SlabSize<NL>::kStride;
                march_other<NL, COUNT>((int) SF(S_MODE), sl, tab, row0, line_stride, nbins, C.bintov, C.tautail, lane, tally);
                __syncwarp();
            }
        }
        if (items.item_start == nullptr && items.ticket_pairs > 0 && !last_run) {
            __threadfence();  // this run's additions are visible before the next run is released
            __syncwarp();
            if (lane == 0) atomicExch(items.line_done + line, run + 1);
        }
        if (STREAM && last_run) row_done(line);
    }
    if (COUNT) {
        unsigned long long pix = tally.pix, vg = 7ull * tally.inner;
        #pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            pix += __shfl_down_sync(kFull, pix, d);
            vg += __shfl_down_sync(kFull, vg, d);
        }
        if (lane == 0) {
            atomicAdd(&counters[0], n_pairs);
            atomicAdd(&counters[1], pix);
            atomicAdd(&counters[2], vg);
            atomicAdd(&counters[3], 32ull * tally.iter * NL);
            #pragma unroll
            for (int r = 0; r < 5; ++r) atomicAdd(&counters[4 + r], (unsigned long long) tally.route[r]);
        }
    }
}

#undef SF
#undef LF
#undef FS
#undef FLF

__global__ void k_voigt_profile(const double *__restrict__ x, const double *__restrict__ y, double *__restrict__ out,
                                int64_t n, int voigt)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double yy = y[i];
    if (voigt == FSB_VOIGT_FAST && fast_domain(yy)) {
        FastCoef fc;
        fast_coefs(yy, fc);
        out[i] = voigt_fast(x[i], fc, d_gtable);
    } else {
        out[i] = voigt_exact(x[i], yy, erfcx(yy));
    }
}

template <int KERNEL, int NL>
int launch_tau_k(const fsb_index *idx, const InterpConsts &c, const ItemPlan &plan, int *next_item, const float *pos,
                 const float *vel, const float *dens, const float *temp, const float *h, const float *cells, double *out,
                 unsigned long long *ctr, int precision, cudaStream_t stream, int *chunk_done, int *host_flags, int chunk_lines,
                 const fsb_push &push)
{
    int dev = 0, sms = 0, per_sm = 0;
    FSB_CUDA_TRY(cudaGetDevice(&dev));
    FSB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int n_items = (int) plan.n_items;
    auto go = [&](auto kern, size_t smem) -> int {
        FSB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        FSB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTauThreads, smem));
        const int grid = std::max(1, std::min((n_items + kTauWarps - 1) / kTauWarps, sms * std::max(per_sm, 1)));
        count_launch();
        kern<<<grid, kTauThreads, smem, stream>>>(c, plan.items, n_items, next_item, idx->offsets, idx->zorder, idx->particle, idx->dr2, idx->axis,
                                                  pos, vel, dens, temp, h, cells, out, plan.scratch_rows.as<double>(),
                                                  plan.n_items * (int64_t) c.nbins, ctr, chunk_done, host_flags, chunk_lines, push);
        FSB_CUDA_TRY(cudaGetLastError());
        return FSB_OK;
    };
    constexpr size_t smem64 = tau_smem_bytes<NL, false>(), smem32 = tau_smem_bytes<NL, true>();
    // the row-streaming variant exists without counters only (the host one-shot entry never asks for them)
    if ((chunk_done || push.npeers > 0) && !ctr)
        return precision == FSB_PRECISION_FP32 ? go(k_tau<KERNEL, NL, false, true, true>, smem32)
                                               : go(k_tau<KERNEL, NL, false, false, true>, smem64);
    if (chunk_done || push.npeers > 0) {
        set_error("launch_tau: row streaming and counters are exclusive");
        return FSB_EINVAL;
    }
    if (precision == FSB_PRECISION_FP32)
        return ctr ? go(k_tau<KERNEL, NL, true, true, false>, smem32) : go(k_tau<KERNEL, NL, false, true, false>, smem32);
    return ctr ? go(k_tau<KERNEL, NL, true, false, false>, smem64) : go(k_tau<KERNEL, NL, false, false, false>, smem64);
}

template <int NL>
int launch_tau_nl(const fsb_index *idx, const InterpConsts &c, const ItemPlan &plan, int *next_item, const float *pos,
                  const float *vel, const float *dens, const float *temp, const float *h, const float *cells, double *out,
                  unsigned long long *ctr, int precision, cudaStream_t stream, int *chunk_done, int *host_flags, int chunk_lines,
                  const fsb_push &push)
{
    switch (c.kernel) {
    case FSB_KERNEL_TOPHAT: return launch_tau_k<FSB_KERNEL_TOPHAT, NL>(idx, c, plan, next_item, pos, vel, dens, temp, h, cells, out, ctr, precision, stream, chunk_done, host_flags, chunk_lines, push);
    case FSB_KERNEL_CUBIC: return launch_tau_k<FSB_KERNEL_CUBIC, NL>(idx, c, plan, next_item, pos, vel, dens, temp, h, cells, out, ctr, precision, stream, chunk_done, host_flags, chunk_lines, push);
    case FSB_KERNEL_VORONOI: return launch_tau_k<FSB_KERNEL_VORONOI, NL>(idx, c, plan, next_item, pos, vel, dens, temp, h, cells, out, ctr, precision, stream, chunk_done, host_flags, chunk_lines, push);
    case FSB_KERNEL_QUINTIC: return launch_tau_k<FSB_KERNEL_QUINTIC, NL>(idx, c, plan, next_item, pos, vel, dens, temp, h, cells, out, ctr, precision, stream, chunk_done, host_flags, chunk_lines, push);
    default: set_error("unknown kernel id %d", c.kernel); return FSB_EINVAL;
    }
}

}  // namespace

int tau_max_fused_lines() { return kMaxTauLines; }

namespace {
struct PinnedFlags {  // zero-initialised ints in pinned, device-visible host memory
    int *ptr = nullptr;
    int alloc(size_t n)
    {
        FSB_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&ptr), sizeof(int) * n, cudaHostAllocMapped | cudaHostAllocPortable));
        for (size_t i = 0; i < n; ++i) ptr[i] = 0;
        return FSB_OK;
    }
    ~PinnedFlags() { if (ptr) cudaFreeHost(ptr); }
};
}  // namespace

// c.nlines (1..kMaxTauLines) lines of one ion in one pass; out[l][nlos][nbins].
int launch_tau(const fsb_index *idx, const InterpConsts &c, const float *pos, const float *vel, const float *dens,
               const float *temp, const float *h, const float *cells, double *out, fsb_counters *counters, int precision,
               cudaStream_t stream, HostSink *sink, const fsb_push *push_in)
{
    fsb_push push;
    memset(&push, 0, sizeof(push));
    if (push_in) push = *push_in;
    if (sink) sink->streamed = false;
    if (idx->nlos == 0 || idx->npairs == 0) return FSB_OK;
    if (precision != FSB_PRECISION_FP64 && precision != FSB_PRECISION_FP32) {
        set_error("launch_tau: unknown precision %d", precision);
        return FSB_EINVAL;
    }
    if (c.nlines < 1 || c.nlines > kMaxTauLines) {
        set_error("launch_tau: %d fused lines (1..%d)", c.nlines, kMaxTauLines);
        return FSB_EINVAL;
    }
    ItemPlan plan;
    FSB_TRY(plan_items(idx, c.seg_pairs, c.nbins, c.nlines, stream, plan));
    // One output row per line: hand the lists out in ticketed runs (fsb_items.cuh).  The run length is a multiple
    // of the particle batch, so the batches, and with them the order of every addition, are those of a whole-list pass.
    if (!plan.segmented) {
        int64_t ticket = 256;
        if (const char *env = getenv("FSB200_TICKET_PAIRS")) ticket = std::max(0ll, atoll(env)) / 32 * 32;  // tuning hook; 0 = whole lists
        if (ticket > 0 && idx->max_list > ticket) FSB_TRY(plan_tickets(idx, (int) ticket, c.line0, c.nrange, stream, plan));
        if (plan.items.ticket_pairs == 0) plan.n_items = c.nrange;
    }
    Scratch next_item;
    FSB_TRY(next_item.alloc(sizeof(int), stream));
    FSB_CUDA_TRY(cudaMemsetAsync(next_item.ptr, 0, sizeof(int), stream));
    unsigned long long *ctr = reinterpret_cast<unsigned long long *>(counters);
    // streaming to the host: chunks of sightlines of about 16 MB per fused line
    const size_t row_bytes = sizeof(double) * (size_t) c.nbins;
    size_t chunk_bytes = (size_t) 16 << 20, min_lines = 64;
    if (const char *env = getenv("FSB200_STREAM_CHUNK_BYTES")) {  // test hook: small chunks exercise the streaming path on small cases
        chunk_bytes = (size_t) std::max(1ll, atoll(env));
        min_lines = 1;
    }
    const int chunk_lines = (int) std::min<size_t>((size_t) idx->nlos, std::max<size_t>(min_lines, chunk_bytes / row_bytes));
    const int nchunks = (idx->nlos + chunk_lines - 1) / chunk_lines;
    const bool stream_out = sink && sink->host && !ctr && !plan.segmented && nchunks >= 4 && c.nrange == idx->nlos;
    Scratch chunk_done;
    PinnedFlags flags;
    if (stream_out) {
        FSB_TRY(chunk_done.alloc(sizeof(int) * (size_t) nchunks, stream));
        FSB_CUDA_TRY(cudaMemsetAsync(chunk_done.ptr, 0, sizeof(int) * (size_t) nchunks, stream));
        FSB_TRY(flags.alloc((size_t) nchunks));
    }
    int *cd = stream_out ? chunk_done.as<int>() : nullptr;
    if (c.nrange != idx->nlos && plan.segmented) {
        set_error("launch_tau: a sightline range needs one work row per sightline");
        return FSB_EINVAL;
    }
    if (push.npeers > 0 && (plan.segmented || ctr || (sink && sink->host))) {
        set_error("launch_tau: pushing rows to peers needs one work row per sightline, no counters and no host sink");
        return FSB_EINVAL;
    }
    if (c.nlines == 1) FSB_TRY(launch_tau_nl<1>(idx, c, plan, next_item.as<int>(), pos, vel, dens, temp, h, cells, out, ctr, precision, stream, cd, flags.ptr, chunk_lines, push));
    else FSB_TRY(launch_tau_nl<2>(idx, c, plan, next_item.as<int>(), pos, vel, dens, temp, h, cells, out, ctr, precision, stream, cd, flags.ptr, chunk_lines, push));
    FSB_TRY(reduce_items(plan, idx, c.nbins, c.nlines, out, stream));
    if (stream_out) {
        // follow the kernel: copy each chunk as soon as its flag is up; if the kernel ends first (or fails), the
        // remaining chunks are copied after it in stream order
        const volatile int *vf = flags.ptr;
        cudaEvent_t kernel_done;
        FSB_CUDA_TRY(cudaEventCreateWithFlags(&kernel_done, cudaEventDisableTiming));
        FSB_CUDA_TRY(cudaEventRecord(kernel_done, stream));
        bool finished = false;
        int rc = FSB_OK;
        for (int ch = 0; ch < nchunks && rc == FSB_OK; ++ch) {
            unsigned spins = 0;
            while (!finished && vf[ch] == 0) {
                if ((++spins & 0x3ffu) == 0) {
                    const cudaError_t q = cudaEventQuery(kernel_done);
                    if (q == cudaSuccess) finished = true;
                    else if (q != cudaErrorNotReady) {
                        set_error("tau kernel failed while streaming rows: %s", cudaGetErrorString(q));
                        rc = FSB_ECUDA;
                        break;
                    }
                }
            }
            if (rc != FSB_OK) break;
            if (finished) {
                const cudaError_t w = cudaStreamWaitEvent(sink->copy_stream, kernel_done, 0);
                if (w != cudaSuccess) { set_error("cudaStreamWaitEvent: %s", cudaGetErrorString(w)); rc = FSB_ECUDA; break; }
            }
            const int l0 = ch * chunk_lines, nl = std::min(chunk_lines, idx->nlos - l0);
            for (int l = 0; l < c.nlines && rc == FSB_OK; ++l) {
                const size_t off = ((size_t) l * (size_t) idx->nlos + (size_t) l0) * (size_t) c.nbins;
                const cudaError_t e = cudaMemcpyAsync(sink->host + off, out + off, row_bytes * (size_t) nl, cudaMemcpyDeviceToHost, sink->copy_stream);
                if (e != cudaSuccess) { set_error("cudaMemcpyAsync (row streaming): %s", cudaGetErrorString(e)); rc = FSB_ECUDA; }
            }
        }
        cudaEventDestroy(kernel_done);
        // flags and chunk counters are released below: wait for the kernel that writes them
        const cudaError_t e = cudaStreamSynchronize(stream);
        if (rc == FSB_OK && e != cudaSuccess) { set_error("tau kernel: %s", cudaGetErrorString(e)); rc = FSB_ECUDA; }
        if (rc != FSB_OK) return rc;
        sink->streamed = true;
    }
    return FSB_OK;
}

int launch_voigt(const double *x, const double *y, double *out, int64_t n, int voigt, cudaStream_t stream)
{
    if (n <= 0) return FSB_OK;
    count_launch(); k_voigt_profile<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(x, y, out, n, voigt);
    FSB_CUDA_TRY(cudaGetLastError());
    return FSB_OK;
}

}  // namespace fsb
