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

struct Bases {
    double E, AE, Gdd, H, G4;
};

__device__ inline void mf_kern_const(int kernel, const double* p, KernConst& o) {
    if (kernel == MF_KERNEL_QP) {      // (ell_e, ell_p, period, wn), kernels.py:282
        double le = p[0], lp = p[1], P = p[2], wn = p[3];
        double f2 = lp * lp, f3 = le * le;
        o.P = P;
        o.wn2 = wn * wn;
        o.i_f2 = 1.0 / f2;
        o.i_f3 = 1.0 / f3;
        o.i_f2P = 1.0 / (f2 * P);
        o.i_f2P2 = 1.0 / (f2 * (P * P));
        o.i_f2P3 = 1.0 / (f2 * (P * P * P));
        o.i_f2P4 = 1.0 / (f2 * ((P * P) * (P * P)));
        o.i_l4 = o.i_l6 = o.i_l8 = 0.0;
    } else {                           // (ell, wn), kernels.py:91
        double l = p[0], wn = p[1];
        double f2 = l * l;
        o.P = 1.0;
        o.wn2 = wn * wn;
        o.i_f2 = 1.0 / f2;
        o.i_f3 = 0.0;
        o.i_f2P = o.i_f2P2 = o.i_f2P3 = o.i_f2P4 = 0.0;
        o.i_l4 = 1.0 / (f2 * f2);
        o.i_l6 = 1.0 / (f2 * f2 * f2);
        o.i_l8 = 1.0 / ((f2 * f2) * (f2 * f2));
    }
}

// Base functions at lag r = t_i - t_j, white noise NOT included.
// The sin/cos argument is formed exactly as the reference forms it,
// x = (pi * r) / P with a true division (kernels.py:295): E is sensitive to the
// last bit of x (dE/E ~ 4 s/ell_p^2 dx), everything else tolerates a few ulp.
template <int KERNEL, bool HIGH>
__device__ __forceinline__ void mf_bases(double r, const KernConst& k, Bases& o) {
    if (KERNEL == MF_KERNEL_QP) {
        double x = (MF_PI * r) / k.P;
        double s, c;
        sincos(x, &s, &c);
        double ss = s * s, cc = c * c, sc = s * c;
        double rr3 = r * k.i_f3;                        // r / f3
        // kernels.py:297  exp(-(2 s s / f2) - 0.5 r r / f3)
        double E = exp(-(2.0 * ss) * k.i_f2 - 0.5 * r * rr3);
        // kernels.py:365  A = 4 pi s c / (f2 P) + r / f3 ; dK/dt2 = A E, dK/dt1 = -A E
        double A = (4.0 * MF_PI) * sc * k.i_f2P + rr3;
        // kernels.py:400-405  q0 = 1/f3 + 4 pi^2 (c^2 - s^2) / (f2 P^2)
        double q0 = k.i_f3 + (4.0 * MF_PI * MF_PI) * cc * k.i_f2P2 - (4.0 * MF_PI * MF_PI) * ss * k.i_f2P2;
        o.E = E;
        o.AE = A * E;
        o.Gdd = (q0 - A * A) * E;
        if (HIGH) {
            // kernels.py:448-456 (third derivative; j2 uses sin*sin as the
            // reference does - quirk Q4) and kernels.py:573-587 (fourth)
            double j2 = rr3 + (4.0 * MF_PI) * ss * k.i_f2P;
            double j8 = (16.0 * MF_PI * MF_PI * MF_PI) * sc * k.i_f2P3;
            o.H = (j2 * j2 * j2 - 3.0 * q0 * j2 - j8) * E;
            double A2 = A * A;
            double p4 = A2 * A2 - 6.0 * q0 * A2 + 3.0 * q0 * q0
                        - (64.0 * MF_PI * MF_PI * MF_PI) * sc * A * k.i_f2P3
                        - (16.0 * MF_PI * MF_PI * MF_PI * MF_PI) * ss * k.i_f2P4
                        + (16.0 * MF_PI * MF_PI * MF_PI * MF_PI) * cc * k.i_f2P4;
            o.G4 = p4 * E;
        } else {
            o.H = 0.0;
            o.G4 = 0.0;
        }
    } else {
        // kernels.py:96-259
        double r2 = r * r;
        double E = exp(-0.5 * r2 * k.i_f2);
        o.E = E;
        o.AE = r * k.i_f2 * E;                          // dK/dt2 (kernels.py:143)
        o.Gdd = (k.i_f2 - r2 * k.i_l4) * E;             // kernels.py:165
        if (HIGH) {
            o.H = (r2 * r * k.i_l6 - 3.0 * r * k.i_l4) * E;                   // kernels.py:191
            o.G4 = (r2 * r2 * k.i_l8 - 6.0 * r2 * k.i_l6 + 3.0 * k.i_l4) * E; // kernels.py:251
        } else {
            o.H = 0.0;
            o.G4 = 0.0;
        }
    }
}

// plain K(r) of a kernel family selected at run time (extra kernel Z(t))
__device__ __forceinline__ double mf_kernel_value(int kernel, double r, const KernConst& k) {
    if (kernel == MF_KERNEL_QP) {
        double s = sin((MF_PI * r) / k.P);
        return exp(-(2.0 * s * s) * k.i_f2 - 0.5 * r * r * k.i_f3);
    }
    return exp(-0.5 * r * r * k.i_f2);
}

__device__ inline void mf_set_pair(SampleCoefs& sc, int p, int q, double cE, double cAE,
                                   double cGdd, double cH, double cG4, double cW) {
    BlockCoef u = {cE, cAE, cGdd, cH, cG4, cW};
    sc.c[p][q] = u;
    if (p != q) {   // lower block (q, p) is the transpose: odd functions change sign
        BlockCoef l = {cE, -cAE, cGdd, -cH, cG4, cW};
        sc.c[q][p] = l;
    }
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
        // BIGgp.py:167-234; block order [rv, rhk, bis]
        const double vc = hyper[npar], vr = hyper[npar + 1], lc = hyper[npar + 2],
                     bc = hyper[npar + 3], br = hyper[npar + 4];
        mf_set_pair(sc, 0, 0, vc * vc, 0, vr * vr, 0, 0, (vc + vr) * (vc + vr));
        mf_set_pair(sc, 1, 1, lc * lc, 0, 0, 0, 0, lc * lc);
        mf_set_pair(sc, 2, 2, bc * bc, 0, br * br, 0, 0, (bc + br) * (bc + br));
        mf_set_pair(sc, 0, 1, vc * lc, -(vr * lc), 0, 0, 0, vc * lc + vr * lc);
        mf_set_pair(sc, 0, 2, vc * bc, vc * br - vr * bc, vr * br, 0, 0,
                    vc * bc + vr * br + vc * br + vr * bc);
        mf_set_pair(sc, 1, 2, lc * bc, lc * br, 0, 0, 0, lc * bc + lc * br);
        if (pb.framework == MF_FRAMEWORK_BIGGP_JITT && extra != nullptr)
            for (int p = 0; p < 3; ++p) sc.jit2[p] = extra[p] * extra[p];
    }
}

// One covariance entry: block (P, Q), epochs (i, j), from the shared bases.
// `ez` = extra-kernel value at this lag (only read when P == Q and amp2 != 0).
__device__ __forceinline__ double mf_entry(const SampleCoefs& sc, const Bases& b, int P, int Q,
                                           bool same_epoch, double ez, double diag_add,
                                           double nugget) {
    const BlockCoef& c = sc.c[P][Q];
    double v = c.cE * b.E + c.cGdd * b.Gdd + c.cG4 * b.G4 + c.cAE * b.AE + c.cH * b.H;
    if (same_epoch) v += c.cW * sc.k.wn2;
    if (P == Q) {
        if (sc.amp2[P] != 0.0) v += sc.amp2[P] * (ez + (same_epoch ? sc.e.wn2 : 0.0));
        if (same_epoch) {
            v = v + diag_add + sc.jit2[P];
            if (nugget != 0.0) return (1.0 - nugget) * v + nugget * v;
        }
    }
    if (nugget != 0.0) v = (1.0 - nugget) * v;
    return v;
}
