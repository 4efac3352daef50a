// Device-side projected-Newton box QP (Tassa 2014), the batched replacement of boxQP(H,g,lower,upper,x0)
// in src/boxQP.jl:29-188 of the reference.
//
// Arithmetic contract: every sum is sequential in index order with separate IEEE multiply and
// add (__dmul_rn/__dadd_rn; ptxas never contracts those into FMA), IEEE sqrt and divide.  The CPU
// oracle states the same order, so every branch decision (clamped set, Armijo steps, result code)
// and every output bit is reproducible between the two.
//
// One thread runs one QP.  MM is the compile-time capacity (arrays live in registers when the
// loops unroll, i.e. for the specialised small-m kernels; in local memory for the generic MM=16).
#pragma once
#include "ddp_common.cuh"

#define DMUL(a, b) __dmul_rn((a), (b))
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))
#define DDIV(a, b) __ddiv_rn((a), (b))

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return clamp_jl(v, lo, hi); }   // NaN-preserving, boxQP.jl:58

// Upper Cholesky factor of A[idx,idx] (reads the upper triangle only; R'R = A), nf x nf, written to
// R with leading dimension ldr.  Returns false when a pivot is <= 0 or NaN (LAPACK dpotrf's test).
// CMP (compact): the loops are kept rolled.  One QP per WARP on a single lane (how the n=32, m=8 tile kernel first ran its QP; it now uses boxqp_warp8) makes the fully
// unrolled m = 8 code an instruction-cache problem (55 % of the stall samples were instruction fetches); the rolled code is the
// same sequence of operations.
template <int MM, bool CMP = false>
__device__ __forceinline__ bool chol_upper_sub(const double* A, int lda, const int* idx, int nf, double* R, int ldr) {
    constexpr int UF = CMP ? 1 : MM;
#pragma unroll UF
    for (int j = 0; j < MM; j++) {
        if (j < nf) {
#pragma unroll UF
            for (int i = 0; i < MM; i++) {
                if (i < j) {
                    double s = A[idx[i] + lda * idx[j]];
                    for (int p = 0; p < i; p++) s = DSUB(s, DMUL(R[p + ldr * i], R[p + ldr * j]));
                    R[i + ldr * j] = DDIV(s, R[i + ldr * i]);
                }
            }
            double d = A[idx[j] + lda * idx[j]];
            for (int p = 0; p < j; p++) d = DSUB(d, DMUL(R[p + ldr * j], R[p + ldr * j]));
            if (!(d > 0.0)) return false;
            R[j + ldr * j] = __dsqrt_rn(d);
        }
    }
    return true;
}

// y = R' \ b ; x = R \ y   (in place in v), R upper nf x nf
template <int MM, bool CMP = false>
__device__ __forceinline__ void chol_solve(const double* R, int ldr, int nf, double* v) {
    constexpr int UF = CMP ? 1 : MM;
#pragma unroll UF
    for (int i = 0; i < MM; i++) {
        if (i < nf) {
            double s = v[i];
            for (int p = 0; p < i; p++) s = DSUB(s, DMUL(R[p + ldr * i], v[p]));
            v[i] = DDIV(s, R[i + ldr * i]);
        }
    }
#pragma unroll UF
    for (int ii = 0; ii < MM; ii++) {
        int i = nf - 1 - ii;
        if (i >= 0) {
            double s = v[i];
            for (int p = i + 1; p < nf; p++) s = DSUB(s, DMUL(R[i + ldr * p], v[p]));
            v[i] = DDIV(s, R[i + ldr * i]);
        }
    }
}

// x'g + ((0.5 x') H) x   -- the way Julia parses `x'g + 0.5x'H*x` (boxQP.jl:63)
template <int MM, bool CMP = false>
__device__ __forceinline__ double qp_value(int m, const double* H, int ldh, const double* g, const double* x) {
    constexpr int UF = CMP ? 1 : MM;
    double s1 = 0.0;
#pragma unroll UF
    for (int i = 0; i < MM; i++)
        if (i < m) s1 = DADD(s1, DMUL(x[i], g[i]));
    double s2 = 0.0;
#pragma unroll UF
    for (int j = 0; j < MM; j++) {
        if (j < m) {
            double t = 0.0;
#pragma unroll UF
            for (int i = 0; i < MM; i++)
                if (i < m) t = DADD(t, DMUL(DMUL(0.5, x[i]), H[i + ldh * j]));
            s2 = DADD(s2, DMUL(t, x[j]));
        }
    }
    return DADD(s1, s2);
}

// Returns the reference's result code 0..6, or -1 where the reference's `cholesky` would throw.
// Outputs: x[m]; Rf (nf x nf upper factor of the last factorisation, leading dim ldr); free_mask; nfactor.
template <int MM, bool CMP = false>
__device__ int boxqp_seq(int m, const double* H, int ldh, const double* g, const double* lower, const double* upper,
                         const double* x0, const QPOpts& o, double* x, double* Rf, int ldr, unsigned* free_mask_out,
                         int* nfactor_out, int* nf_out = nullptr) {
    constexpr int UF = CMP ? 1 : MM;
    unsigned clamped = 0u, free_mask = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
    const unsigned all_mask = free_mask;
    double oldvalue = 0.0;
    int result = 0, nfactor = 0, nf = 0;
    int idx[MM];
    double grad[MM], search[MM], xc[MM], tmp[MM];
#pragma unroll UF
    for (int i = 0; i < MM; i++)
        if (i < m) x[i] = clampd(x0[i], lower[i], upper[i]);            // boxQP.jl:58
    double value = qp_value<MM, CMP>(m, H, ldh, g, x);                        // :63
    int iter = 1;
    while (iter <= o.max_iter) {                                         // :71
        if (result != 0) break;                                          // :73
        if (iter > 1 && DSUB(oldvalue, value) < DMUL(o.min_rel_improve, fabs(oldvalue))) {   // :78
            result = 4;
            break;
        }
        oldvalue = value;
#pragma unroll UF
        for (int i = 0; i < MM; i++) {                                   // grad = g + H*x  :85
            if (i < m) {
                double s = 0.0;
#pragma unroll UF
                for (int j = 0; j < MM; j++)
                    if (j < m) s = DADD(s, DMUL(H[i + ldh * j], x[j]));
                grad[i] = DADD(g[i], s);
            }
        }
        unsigned old_clamped = clamped;
        clamped = 0u;
#pragma unroll UF
        for (int i = 0; i < MM; i++)                                     // :92-94 exact equality on bounds
            if (i < m && ((x[i] == lower[i] && grad[i] > 0.0) || (x[i] == upper[i] && grad[i] < 0.0))) clamped |= 1u << i;
        free_mask = all_mask & ~clamped;
        if (clamped == all_mask) {                                       // :98
            result = 6;
            break;
        }
        if (iter == 1 || old_clamped != clamped) {                       // :104-117
            nf = 0;
#pragma unroll UF
            for (int i = 0; i < MM; i++)
                if (i < m && ((free_mask >> i) & 1u)) idx[nf++] = i;
            if (!chol_upper_sub<MM, CMP>(H, ldh, idx, nf, Rf, ldr)) {
                *free_mask_out = free_mask;
                *nfactor_out = nfactor;
                if (nf_out) *nf_out = 0;
                return -1;                                               // PosDefException
            }
            nfactor++;
        }
        double gs = 0.0;
        for (int p = 0; p < nf; p++) gs = DADD(gs, DMUL(grad[idx[p]], grad[idx[p]]));   // norm(grad[free]) :120
        // One free variable: sqrt(fl(g*g)) == |g| exactly in binary floating point with round-to-nearest as long as g*g
        // neither underflows nor overflows, so the (expensive) square root is skipped without changing a single bit.
        const double gnorm = (MM == 1 && nf == 1 && gs > 1e-280 && gs < 1e280) ? fabs(grad[idx[0]]) : __dsqrt_rn(gs);
        if (gnorm < o.min_grad) {
            result = 5;
            break;
        }
        // grad_clamped = g + H*(x.*clamped)  :127 ; only the free entries are needed
        for (int p = 0; p < nf; p++) {
            int i = idx[p];
            double s = 0.0;
#pragma unroll UF
            for (int j = 0; j < MM; j++)
                if (j < m) s = DADD(s, DMUL(H[i + ldh * j], ((clamped >> j) & 1u) ? x[j] : DMUL(x[j], 0.0)));
            tmp[p] = DADD(g[i], s);
        }
        chol_solve<MM, CMP>(Rf, ldr, nf, tmp);                                // Hfree\(Hfree'\grad_clamped[free])
#pragma unroll UF
        for (int i = 0; i < MM; i++) search[i] = 0.0;
        for (int p = 0; p < nf; p++) search[idx[p]] = DSUB(-tmp[p], x[idx[p]]);          // :129
        double sdotg = 0.0;
#pragma unroll UF
        for (int i = 0; i < MM; i++)
            if (i < m) sdotg = DADD(sdotg, DMUL(search[i], grad[i]));    // :132
        if (sdotg >= 0.0) break;                                         // :133 leaves result == 0
        double step = 1.0;                                               // :138
#pragma unroll UF
        for (int i = 0; i < MM; i++)
            if (i < m) xc[i] = clampd(DADD(x[i], DMUL(step, search[i])), lower[i], upper[i]);
        double vc = qp_value<MM, CMP>(m, H, ldh, g, xc);
        while (DDIV(DSUB(vc, oldvalue), DMUL(step, sdotg)) < o.armijo) { // :142
            step = DMUL(step, o.step_dec);
#pragma unroll UF
            for (int i = 0; i < MM; i++)
                if (i < m) xc[i] = clampd(DADD(x[i], DMUL(step, search[i])), lower[i], upper[i]);
            vc = qp_value<MM, CMP>(m, H, ldh, g, xc);
            if (step < o.min_step) {
                result = 2;
                break;
            }
        }
#pragma unroll UF
        for (int i = 0; i < MM; i++)
            if (i < m) x[i] = xc[i];                                     // :161
        value = vc;
        iter++;
    }
    if (iter == o.max_iter) result = 1;                                  // :167 (quirk Q4)
    *free_mask_out = free_mask;
    *nfactor_out = nfactor;
    if (nf_out) *nf_out = nf;                                            // size of the factor held in Rf
    return result;
}


// sqrt(gs) for the rare |grad| outside the range where sqrt(fl(g*g)) == |g| holds exactly; a real call, so that the compiler cannot
// turn the range test into a select that evaluates the square root on every iteration
static __device__ __noinline__ double qp_sqrt_slow(double v) { return __dsqrt_rn(v); }

// The same projected-Newton iteration for ONE variable (m = 1: config 3), written on scalars: every operation, comparison and
// rounding is the one boxqp_seq<1> performs in the same order (so result code, free set, factor and x agree bit for bit with it and
// with the oracle), but without index sets, mask loops and runtime-bounded solves.  h = H[0,0].
__device__ __forceinline__ double qp1_value(double h, double g, double x) {          // x'g + ((0.5 x') H) x, sums started from 0.0
    const double s1 = DADD(0.0, DMUL(x, g));
    const double t = DADD(0.0, DMUL(DMUL(0.5, x), h));
    const double s2 = DADD(0.0, DMUL(t, x));
    return DADD(s1, s2);
}

__device__ __forceinline__ int boxqp_scalar(double h, double g, double lower, double upper, double x0, const QPOpts& o, double* x_out,
                                            double* R_out, unsigned* free_mask_out, int* nfactor_out) {
    bool clamped = false;
    unsigned free_mask = 1u;
    double oldvalue = 0.0, R = 0.0;
    int result = 0, nfactor = 0;
    double x = clampd(x0, lower, upper);                                 // boxQP.jl:58
    double value = qp1_value(h, g, x);                                   // :63
    int iter = 1;
    while (iter <= o.max_iter) {                                         // :71
        if (result != 0) break;                                          // :73
        if (iter > 1 && DSUB(oldvalue, value) < DMUL(o.min_rel_improve, fabs(oldvalue))) {   // :78
            result = 4;
            break;
        }
        oldvalue = value;
        const double grad = DADD(g, DADD(0.0, DMUL(h, x)));              // :85
        const bool old_clamped = clamped;
        clamped = (x == lower && grad > 0.0) || (x == upper && grad < 0.0);      // :92-94
        free_mask = clamped ? 0u : 1u;
        if (clamped) {                                                   // :98 all clamped
            result = 6;
            break;
        }
        if (iter == 1 || old_clamped != clamped) {                       // :104-117
            if (!(h > 0.0)) {                                            // PosDefException
                *x_out = x; *R_out = R; *free_mask_out = free_mask; *nfactor_out = nfactor;
                return -1;
            }
            R = __dsqrt_rn(h);
            nfactor++;
        }
        const double gs = DADD(0.0, DMUL(grad, grad));                   // norm(grad[free]) :120 (see boxqp_seq for the |g| shortcut)
        // gnorm only decides `gnorm < min_grad`: outside the exact range both |grad| and sqrt(gs) are below 1e-140 or above 1e140
        // (or non-finite), so for any min_grad strictly inside (1e-139, 1e139) the decision is the same and no square root is needed
        double gnorm = fabs(grad);
        if (!(gs > 1e-280 && gs < 1e280) && !(o.min_grad > 1e-139 && o.min_grad < 1e139)) gnorm = qp_sqrt_slow(gs);
        if (gnorm < o.min_grad) {
            result = 5;
            break;
        }
        const double tmp0 = DADD(g, DADD(0.0, DMUL(h, DMUL(x, 0.0))));   // grad_clamped = g + H*(x.*clamped), nothing clamped here  :127
        const double sol = DDIV(DDIV(tmp0, R), R);                        // Hfree \ (Hfree' \ grad_clamped)
        const double search = DSUB(-sol, x);                             // :129
        const double sdotg = DADD(0.0, DMUL(search, grad));              // :132
        if (sdotg >= 0.0) break;                                         // :133 leaves result == 0
        double step = 1.0;                                               // :138
        double xc = clampd(DADD(x, DMUL(step, search)), lower, upper);
        double vc = qp1_value(h, g, xc);
        while (DDIV(DSUB(vc, oldvalue), DMUL(step, sdotg)) < o.armijo) { // :142
            step = DMUL(step, o.step_dec);
            xc = clampd(DADD(x, DMUL(step, search)), lower, upper);
            vc = qp1_value(h, g, xc);
            if (step < o.min_step) {
                result = 2;
                break;
            }
        }
        x = xc;                                                          // :161
        value = vc;
        iter++;
    }
    if (iter == o.max_iter) result = 1;                                  // :167 (quirk Q4)
    *x_out = x;
    *R_out = R;
    *free_mask_out = free_mask;
    *nfactor_out = nfactor;
    return result;
}
