// Device-side covariance algebra shared by the stand-alone build kernel
// (mf_cov.cu) and any kernel that evaluates covariance tiles on the fly.
//
// Every block (P, Q) of every framework is
//     sum_k coef_k(P, Q; hyper) * base_k(r) + cW * wn^2 * delta_ij  (+ diagonal terms)
// with the base functions {E, AE, Gdd, H, G4} of SURVEY.md section 2.1 sharing one
// sincos + one exp per (i, j).  Reference: kernels.py:80-617 (the kernel classes),
// BIGgp.py:167-278, BIGgpPlusJitt.py:237-282, SMALLgp.py:83-87, 139-249.
#pragma once
#include "mf_common.cuh"

#define MF_PI 3.141592653589793

struct BlockCoef {
    double cE, cAE, cGdd, cH, cG4, cW;
};

// per-sample constants of one stationary kernel (QP or SE)
struct KernConst {
    double P;        // period (QP)
    double wn2;      // wn^2
    // QP: f2 = ell_p^2, f3 = ell_e^2 ; SE: f2 = ell^2
    double i_f2, i_f3;          // 1/f2, 1/f3
    double i_f2P;               // 1/(f2 P)
    double i_f2P2;              // 1/(f2 P^2)
    double i_f2P3;              // 1/(f2 P^3)
    double i_f2P4;              // 1/(f2 P^4)
    // SE: 1/ell^2 (= i_f2), 1/ell^4, 1/ell^6, 1/ell^8
    double i_l4, i_l6, i_l8;
};

struct SampleCoefs {
    KernConst k;                 // main kernel
    KernConst e;                 // extra kernel Z(t) (SMALLgp)
    double amp2[MF_MAX_SERIES];  // a4^2 per series (SMALLgp.py:152-155, 174)
    double jit2[MF_MAX_SERIES];  // jitter^2 per series (BIGgpPlusJitt.py:263-265)
    BlockCoef c[MF_MAX_SERIES][MF_MAX_SERIES];   // [row block][col block]
};

// the three-series (BIGgp / BIGgpPlusJitt) subset of SampleCoefs, same member names
struct BigCoefs {
    KernConst k;
    KernConst e;                 // unused (amp2 == 0)
    double amp2[3];
    double jit2[3];
    BlockCoef c[3][3];
};

// E, Gdd, G4 are even in the lag and AE is odd.  H is odd for the SE family but NOT
// for the quasi-periodic one: the reference builds it from sin*sin (quirk Q4,
// kernels.py:449), so H(-r) != -H(r) and both are kept: Hf = H(r), Hr = H(-r).
struct Bases {
    double E, AE, Gdd, Hf, Hr, G4;
};

// Explicitly rounded FP64 operations: the compiler neither fuses nor splits these,
// so every inlined instance of the covariance algebra performs the same IEEE
// operations and K(i, j) == K(j, i) bit for bit whichever thread/slot computed it.
__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double df(double a, double b, double c) { return __fma_rn(a, b, c); }

constexpr int MF_KC_FIELDS = 11;
static_assert(sizeof(KernConst) == MF_KC_FIELDS * sizeof(double), "KernConst is an array of doubles");

// Field `f` of the per-sample kernel constants (the order of the struct's members).  One
// function for the serial set-up (all fields by one thread) and the parallel one (a field per
// thread), every operation explicitly rounded: both give the same bits.
__device__ __forceinline__ double mf_kc_field(int kernel, const double* p, int f) {
    if (kernel == MF_KERNEL_QP) {      // (ell_e, ell_p, period, wn), kernels.py:282
        const double le = p[0], lp = p[1], P = p[2], wn = p[3];
        const double f2 = dm(lp, lp), f3 = dm(le, le), P2 = dm(P, P);
        switch (f) {
            case 0: return P;
            case 1: return dm(wn, wn);
            case 2: return __ddiv_rn(1.0, f2);
            case 3: return __ddiv_rn(1.0, f3);
            case 4: return __ddiv_rn(1.0, dm(f2, P));
            case 5: return __ddiv_rn(1.0, dm(f2, P2));
            case 6: return __ddiv_rn(1.0, dm(f2, dm(P2, P)));
            case 7: return __ddiv_rn(1.0, dm(f2, dm(P2, P2)));
            default: return 0.0;
        }
    }
    const double l = p[0], wn = p[1];  // (ell, wn), kernels.py:91
    const double f2 = dm(l, l), f4 = dm(f2, f2);
    switch (f) {
        case 0: return 1.0;
        case 1: return dm(wn, wn);
        case 2: return __ddiv_rn(1.0, f2);
        case 8: return __ddiv_rn(1.0, f4);
        case 9: return __ddiv_rn(1.0, dm(f4, f2));
        case 10: return __ddiv_rn(1.0, dm(f4, f4));
        default: return 0.0;
    }
}

__device__ inline void mf_kern_const(int kernel, const double* p, KernConst& o) {
    double* out = reinterpret_cast<double*>(&o);
#pragma unroll
    for (int f = 0; f < MF_KC_FIELDS; ++f) out[f] = mf_kc_field(kernel, p, f);
}

#define MF_4PI   (4.0 * MF_PI)
#define MF_4PI2  (4.0 * MF_PI * MF_PI)
#define MF_16PI3 (16.0 * MF_PI * MF_PI * MF_PI)
#define MF_64PI3 (64.0 * MF_PI * MF_PI * MF_PI)
#define MF_16PI4 (16.0 * MF_PI * MF_PI * MF_PI * MF_PI)

// Base functions at lag r = t_i - t_j, white noise NOT included.
// Everything is evaluated at |r| and the parity of each function applied
// afterwards, so K(i, j) and K(j, i) come out of identical arithmetic and the
// matrix is symmetric to the last bit (the reference's is: its lower blocks are
// transposes, BIGgp.py:270-271, SMALLgp.py:241).
// The sin/cos argument is formed exactly as the reference forms it,
// x = (pi * r) / P with a true division (kernels.py:295): E is sensitive to the
// last bit of x (dE/E ~ 4 s/ell_p^2 dx), everything else tolerates a few ulp.
template <int KERNEL, bool HIGH>
__device__ __forceinline__ void mf_bases(double r_signed, const KernConst& k, Bases& o) {
    const bool neg = r_signed < 0.0;
    const double r = fabs(r_signed);
    double AEa, Hp = 0.0, Hm = 0.0;
    if (KERNEL == MF_KERNEL_QP) {
        const double x = __ddiv_rn(dm(MF_PI, r), k.P);
        double s, c;
        sincos(x, &s, &c);
        const double ss = dm(s, s), cc = dm(c, c), sc = dm(s, c);
        const double rr3 = dm(r, k.i_f3);                         // r / f3
        // kernels.py:297  exp(-(2 s s / f2) - 0.5 r r / f3)
        const double E = exp(df(dm(-0.5, r), rr3, -dm(dm(2.0, ss), k.i_f2)));
        // kernels.py:365  A = 4 pi s c / (f2 P) + r / f3 ; dK/dt2 = A E, dK/dt1 = -A E
        const double A = df(dm(MF_4PI, sc), k.i_f2P, rr3);
        // kernels.py:400-405  q0 = 1/f3 + 4 pi^2 (c^2 - s^2) / (f2 P^2)
        const double q0 = df(-dm(MF_4PI2, ss), k.i_f2P2, df(dm(MF_4PI2, cc), k.i_f2P2, k.i_f3));
        o.E = E;
        AEa = dm(A, E);
        o.Gdd = dm(df(-A, A, q0), E);
        if (HIGH) {
            // kernels.py:448-456: (j1 j2 + j3 j4 + 2 j5 j6 - j8) = j2^3 - 3 q0 j2 - j8 with
            // j2 = r/f3 + 4 pi s s/(f2 P)  (sin*sin as shipped) and j8 = 16 pi^3 c s/(f2 P^3)
            const double j2e = dm(dm(MF_4PI, ss), k.i_f2P);      // even part of j2
            const double j8 = dm(dm(MF_16PI3, sc), k.i_f2P3);
            const double j2p = da(j2e, rr3), j2m = da(j2e, -rr3);
            const double q3 = dm(3.0, q0);
            Hp = dm(df(j2p, dm(j2p, j2p), -df(q3, j2p, j8)), E);   // at +|r|
            Hm = dm(df(j2m, dm(j2m, j2m), -df(q3, j2m, -j8)), E);  // at -|r|
            // kernels.py:573-587: A^4 - 6 q0 A^2 + 3 q0^2 - 64 pi^3 c s A/(f2 P^3)
            //                     - 16 pi^4 s^2/(f2 P^4) + 16 pi^4 c^2/(f2 P^4)
            const double A2 = dm(A, A);
            double p4 = df(A2, A2, dm(q3, q0));
            p4 = df(-dm(6.0, q0), A2, p4);
            p4 = df(-dm(dm(MF_64PI3, sc), A), k.i_f2P3, p4);
            p4 = df(-dm(MF_16PI4, ss), k.i_f2P4, p4);
            p4 = df(dm(MF_16PI4, cc), k.i_f2P4, p4);
            o.G4 = dm(p4, E);
        } else {
            o.G4 = 0.0;
        }
    } else {
        // kernels.py:96-259
        const double r2 = dm(r, r);
        const double E = exp(dm(dm(-0.5, r2), k.i_f2));
        o.E = E;
        AEa = dm(dm(r, k.i_f2), E);                             // dK/dt2 (kernels.py:143)
        o.Gdd = dm(df(-r2, k.i_l4, k.i_f2), E);                 // kernels.py:165
        if (HIGH) {
            Hp = dm(df(dm(r2, r), k.i_l6, -dm(dm(3.0, r), k.i_l4)), E);         // kernels.py:191
            Hm = -Hp;
            double p4 = df(dm(r2, r2), k.i_l8, dm(3.0, k.i_l4));                // kernels.py:251
            p4 = df(-dm(6.0, r2), k.i_l6, p4);
            o.G4 = dm(p4, E);
        } else {
            o.G4 = 0.0;
        }
    }
    o.AE = neg ? -AEa : AEa;
    o.Hf = neg ? Hm : Hp;
    o.Hr = neg ? Hp : Hm;
}

// plain K(r) of a kernel family selected at run time (extra kernel Z(t))
__device__ __forceinline__ double mf_kernel_value(int kernel, double r_signed, const KernConst& k) {
    const double r = fabs(r_signed);
    if (kernel == MF_KERNEL_QP) {
        const double s = sin(__ddiv_rn(dm(MF_PI, r), k.P));
        return exp(df(dm(-0.5, r), dm(r, k.i_f3), -dm(dm(2.0, dm(s, s)), k.i_f2)));
    }
    return exp(dm(dm(-0.5, dm(r, r)), k.i_f2));
}

template <typename SC>
__device__ inline void mf_set_pair(SC& sc, int p, int q, double cE, double cAE,
                                   double cGdd, double cH, double cG4, double cW) {
    BlockCoef u = {cE, cAE, cGdd, cH, cG4, cW};
    sc.c[p][q] = u;
    if (p != q) {   // lower block (q, p) is the transpose: AE changes sign, H is read at -r
        BlockCoef l = {cE, -cAE, cGdd, cH, cG4, cW};
        sc.c[q][p] = l;
    }
}

// BIGgp.py:167-234, block order [rv, rhk, bis]: coefficients of the upper-form pair `idx`
// (0: (0,0), 1: (1,1), 2: (2,2), 3: (0,1), 4: (0,2), 5: (1,2)) from the scaling parameters.
// Explicitly rounded, shared by the serial and the parallel set-up.
__device__ __forceinline__ BlockCoef mf_big_pair_coef(int idx, const double* s5) {
    const double vc = s5[0], vr = s5[1], lc = s5[2], bc = s5[3], br = s5[4];
    BlockCoef u = {0, 0, 0, 0, 0, 0};
    switch (idx) {
        case 0: u.cE = dm(vc, vc); u.cGdd = dm(vr, vr); u.cW = dm(da(vc, vr), da(vc, vr)); break;
        case 1: u.cE = dm(lc, lc); u.cW = dm(lc, lc); break;
        case 2: u.cE = dm(bc, bc); u.cGdd = dm(br, br); u.cW = dm(da(bc, br), da(bc, br)); break;
        case 3: u.cE = dm(vc, lc); u.cAE = -dm(vr, lc); u.cW = da(dm(vc, lc), dm(vr, lc)); break;
        case 4:
            u.cE = dm(vc, bc); u.cAE = da(dm(vc, br), -dm(vr, bc)); u.cGdd = dm(vr, br);
            u.cW = da(da(da(dm(vc, bc), dm(vr, br)), dm(vc, br)), dm(vr, bc));
            break;
        default: u.cE = dm(lc, bc); u.cAE = dm(lc, br); u.cW = da(dm(lc, bc), dm(lc, br)); break;
    }
    return u;
}
__device__ __forceinline__ void mf_big_pair_index(int idx, int& p, int& q) {
    p = idx < 3 ? idx : (idx == 5 ? 1 : 0);
    q = idx < 3 ? idx : (idx == 3 ? 1 : 2);
}

// Per-sample coefficient table from the raw hyper-parameter rows.
__device__ inline void mf_make_coefs(const mf_problem_t& pb, const double* hyper,
                                     const double* extra, SampleCoefs& sc) {
    const int M = pb.n_series;
    const int npar = (pb.kernel == MF_KERNEL_QP) ? 4 : 2;
    mf_kern_const(pb.kernel, hyper, sc.k);
    for (int p = 0; p < MF_MAX_SERIES; ++p) {
        sc.amp2[p] = 0.0;
        sc.jit2[p] = 0.0;
    }
    for (int p = 0; p < M; ++p)
        for (int q = 0; q < M; ++q) {
            BlockCoef z = {0, 0, 0, 0, 0, 0};
            sc.c[p][q] = z;
        }
    sc.e = sc.k;
    if (pb.framework == MF_FRAMEWORK_SMALLGP) {
        if (M == 1) return;   // quirk Q5: SMALLgp.py:227-243 leaves only diag(yerr^2)
        for (int p = 0; p < M; ++p) {
            const double a1 = hyper[npar + 3 * p], a2 = hyper[npar + 3 * p + 1],
                         a3 = hyper[npar + 3 * p + 2];
            for (int q = p; q < M; ++q) {
                const double b1 = hyper[npar + 3 * q], b2 = hyper[npar + 3 * q + 1],
                             b3 = hyper[npar + 3 * q + 2];
                // SMALLgp.py:204-208 with dgg = -AE, gdg = +AE, gddg = ddgg = -Gdd,
                // dgddg = -H, ddgdg = +H; every class adds wn^2 on the diagonal
                mf_set_pair(sc, p, q, a1 * b1, a1 * b2 - a2 * b1, a2 * b2 - a1 * b3 - a3 * b1,
                            a3 * b2 - a2 * b3, a3 * b3,
                            a1 * b1 + a2 * b2 + a3 * b3 + a2 * b1 + a1 * b2 - a1 * b3 - a3 * b1 +
                                a2 * b3 + a3 * b2);
            }
        }
        if (pb.extra_kernel != MF_KERNEL_NONE && extra != nullptr) {
            mf_kern_const(pb.extra_kernel, extra, sc.e);
            for (int p = 0; p < M; ++p) {
                double a4 = extra[pb.extra_len - M + p];       // SMALLgp.py:155
                sc.amp2[p] = a4 * a4;
            }
        }
    } else {
        for (int idx = 0; idx < 6; ++idx) {
            int p, q;
            mf_big_pair_index(idx, p, q);
            const BlockCoef u = mf_big_pair_coef(idx, hyper + npar);
            mf_set_pair(sc, p, q, u.cE, u.cAE, u.cGdd, u.cH, u.cG4, u.cW);
        }
        if (pb.framework == MF_FRAMEWORK_BIGGP_JITT && extra != nullptr)
            for (int p = 0; p < 3; ++p) sc.jit2[p] = dm(extra[p], extra[p]);
    }
}

// One covariance entry: block (P, Q), epochs (i, j), from the shared bases.
// `ez` = extra-kernel value at this lag (only read when P == Q and amp2 != 0).
template <typename SC>
__device__ __forceinline__ double mf_entry(const SC& sc, const Bases& b, int P, int Q,
                                           bool same_epoch, double ez, double diag_add,
                                           double nugget) {
    const BlockCoef& c = sc.c[P][Q];
    double v = dm(c.cE, b.E);
    v = df(c.cGdd, b.Gdd, v);
    v = df(c.cG4, b.G4, v);
    v = df(c.cAE, b.AE, v);
    v = df(c.cH, (P > Q ? b.Hr : b.Hf), v);
    if (same_epoch) v = df(c.cW, sc.k.wn2, v);
    if (P == Q) {
        if (sc.amp2[P] != 0.0) v = df(sc.amp2[P], da(ez, same_epoch ? sc.e.wn2 : 0.0), v);
        if (same_epoch) {
            v = da(da(v, diag_add), sc.jit2[P]);
            if (nugget != 0.0) return df(1.0 - nugget, v, dm(nugget, v));
        }
    }
    if (nugget != 0.0) v = dm(1.0 - nugget, v);
    return v;
}
