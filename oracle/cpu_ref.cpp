// CPU restatement (C++17 + OpenMP) of the reference's hot path, used ONLY as the timed CPU
// baseline of bench.py (`cpu_baseline`, `--impl reference`) and by tests/ to cross-check the
// NumPy oracle.  TEST / MEASUREMENT INFRASTRUCTURE -- never linked into libddp.so.
//
// Julia is not installed in the build container and the reference has no C sources, so
// oracle/_ref cannot exist for this project; this file is the "port" baseline (kind = "port"),
// labelled "restated reference, not Julia" wherever its numbers are reported.
//
// It follows the reference's operation structure, one trajectory per OpenMP thread:
//   back_pass      src/backward_pass.jl:217-252 (LTI) / :162-215 (LTV) + @end_backward_pass :28-79
//                  -- like the reference it forms fu'Vxx, fx'Vxx and fu'Vxx_reg as separate
//                  products (backward_pass.jl:242-247) instead of sharing W = Vxx*[fx fu]
//   boxQP          src/boxQP.jl:29-188, same fixed summation order as oracle/ddp_oracle.py
//                  (compile with -ffp-contract=off so it stays bit-identical to it)
//   forward_pass   src/forward_pass.jl:9-33 with the linear model of demo_linear.jl:35-50
//
// Layout: the reference's column-major arrays with batch trailing, i.e. [B][T][col][row], exactly
// what libddp.so takes.  Build: oracle/Makefile.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// boxQP arithmetic must not be contracted into FMAs (bit-identical to oracle/ddp_oracle.py)
#define NOFMA __attribute__((optimize("fp-contract=off")))

struct QPOut { int result, nfactor, nf; unsigned free_mask; };

NOFMA double qp_value(int m, const double* H, const double* g, const double* x) {
    double s1 = 0.0;
    for (int i = 0; i < m; i++) s1 = s1 + x[i] * g[i];
    double s2 = 0.0;
    for (int j = 0; j < m; j++) {
        double t = 0.0;
        for (int i = 0; i < m; i++) t = t + (0.5 * x[i]) * H[i + m * j];
        s2 = s2 + t * x[j];
    }
    return s1 + s2;
}

NOFMA bool chol_sub(int m, const double* A, const int* idx, int nf, double* R) {   // R ld m
    for (int j = 0; j < nf; j++) {
        for (int i = 0; i < j; i++) {
            double s = A[idx[i] + m * idx[j]];
            for (int p = 0; p < i; p++) s = s - R[p + m * i] * R[p + m * j];
            R[i + m * j] = s / R[i + m * i];
        }
        double d = A[idx[j] + m * idx[j]];
        for (int p = 0; p < j; p++) d = d - R[p + m * j] * R[p + m * j];
        if (!(d > 0.0)) return false;
        R[j + m * j] = std::sqrt(d);
    }
    return true;
}

NOFMA void chol_solve(int m, const double* R, int nf, double* v) {
    for (int i = 0; i < nf; i++) {
        double s = v[i];
        for (int p = 0; p < i; p++) s = s - R[p + m * i] * v[p];
        v[i] = s / R[i + m * i];
    }
    for (int i = nf - 1; i >= 0; i--) {
        double s = v[i];
        for (int p = i + 1; p < nf; p++) s = s - R[i + m * p] * v[p];
        v[i] = s / R[i + m * i];
    }
}

inline double clampd(double v, double lo, double hi) { return (v > hi) ? hi : ((v < lo) ? lo : v); }   // Julia's clamp: NaN stays NaN

// boxQP.jl:29-188; returns result (or -1 where cholesky throws)
NOFMA QPOut boxqp(int m, const double* H, const double* g, const double* lower, const double* upper, const double* x0,
            double* x, double* Rf, int maxIter = 100, double minGrad = 1e-8, double minRelImprove = 1e-8,
            double stepDec = 0.6, double minStep = 1e-22, double Armijo = 0.1) {
    const unsigned all = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
    unsigned clamped = 0, free_mask = all;
    double oldvalue = 0.0;
    int result = 0, nfactor = 0, nf = 0;
    int idx[32];
    double grad[32], search[32], xc[32], tmp[32];
    for (int i = 0; i < m; i++) x[i] = clampd(x0[i], lower[i], upper[i]);
    double value = qp_value(m, H, g, x);
    int iter = 1;
    while (iter <= maxIter) {
        if (result != 0) break;
        if (iter > 1 && (oldvalue - value) < minRelImprove * std::fabs(oldvalue)) { result = 4; break; }
        oldvalue = value;
        for (int i = 0; i < m; i++) {
            double s = 0.0;
            for (int j = 0; j < m; j++) s = s + H[i + m * j] * x[j];
            grad[i] = g[i] + s;
        }
        unsigned old_clamped = clamped;
        clamped = 0;
        for (int i = 0; i < m; i++)
            if ((x[i] == lower[i] && grad[i] > 0.0) || (x[i] == upper[i] && grad[i] < 0.0)) clamped |= 1u << i;
        free_mask = all & ~clamped;
        if (clamped == all) { result = 6; break; }
        if (iter == 1 || old_clamped != clamped) {
            nf = 0;
            for (int i = 0; i < m; i++)
                if ((free_mask >> i) & 1u) idx[nf++] = i;
            if (!chol_sub(m, H, idx, nf, Rf)) return QPOut{-1, nfactor, 0, free_mask};
            nfactor++;
        }
        double gs = 0.0;
        for (int p = 0; p < nf; p++) gs = gs + grad[idx[p]] * grad[idx[p]];
        if (std::sqrt(gs) < minGrad) { result = 5; break; }
        for (int p = 0; p < nf; p++) {
            int i = idx[p];
            double s = 0.0;
            for (int j = 0; j < m; j++) s = s + H[i + m * j] * (((clamped >> j) & 1u) ? x[j] : x[j] * 0.0);
            tmp[p] = g[i] + s;
        }
        chol_solve(m, Rf, nf, tmp);
        for (int i = 0; i < m; i++) search[i] = 0.0;
        for (int p = 0; p < nf; p++) search[idx[p]] = -tmp[p] - x[idx[p]];
        double sdotg = 0.0;
        for (int i = 0; i < m; i++) sdotg = sdotg + search[i] * grad[i];
        if (sdotg >= 0.0) break;
        double step = 1.0;
        for (int i = 0; i < m; i++) xc[i] = clampd(x[i] + step * search[i], lower[i], upper[i]);
        double vc = qp_value(m, H, g, xc);
        while ((vc - oldvalue) / (step * sdotg) < Armijo) {
            step = step * stepDec;
            for (int i = 0; i < m; i++) xc[i] = clampd(x[i] + step * search[i], lower[i], upper[i]);
            vc = qp_value(m, H, g, xc);
            if (step < minStep) { result = 2; break; }
        }
        for (int i = 0; i < m; i++) x[i] = xc[i];
        value = vc;
        iter++;
    }
    if (iter == maxIter) result = 1;
    return QPOut{result, nfactor, nf, free_mask};
}

// C(r x c) = A'(k x r)' * B(k x c)   all column-major, leading dims = rows
// C(r x c) = A(r x k) * B(k x c)
inline void gemm_nn(int r, int c, int k, const double* A, const double* B, double* C);
inline void gemm_tn(int r, int c, int k, const double* A, const double* B, double* C) {
    double At[64 * 64];                       // A' (r x k), so that the product vectorises as axpys
    for (int i = 0; i < r; i++)
        for (int p = 0; p < k; p++) At[i + r * p] = A[p + (size_t)k * i];
    gemm_nn(r, c, k, At, B, C);
}
inline void gemm_nn(int r, int c, int k, const double* A, const double* B, double* C) {
    for (int j = 0; j < c; j++) {
        double* cj = C + (size_t)r * j;
        for (int i = 0; i < r; i++) cj[i] = 0.0;
        for (int p = 0; p < k; p++) {
            const double bpj = B[p + (size_t)k * j];
            const double* ap = A + (size_t)r * p;
            for (int i = 0; i < r; i++) cj[i] += ap[i] * bpj;
        }
    }
}

struct Work {
    std::vector<double> V, Vn, T1, T2, Qxx, Qux, Quxr, Quu, QuuF, R, Kt, QK, Vreg, tmp;
    Work(int n, int m)
        : V(n * n), Vn(n * n), T1(m * n), T2(n * n), Qxx(n * n), Qux(m * n), Quxr(m * n), Quu(m * m), QuuF(m * m),
          R(m * m), Kt(m * n), QK(m * n), Vreg(n * n), tmp(n * n) {}
};

int back_pass_one(int n, int m, int N, const double* cx, const double* cu, const double* cxx, long cxx_st,
                  const double* cxu, long cxu_st, const double* cuu, long cuu_st, const double* fx, long fx_st,
                  const double* fu, long fu_st, double lam, int regType, const double* lims, const double* u, double* K,
                  double* k, double* Vx, double* Vxx, double* Vxx1, double* Quu_out, double* dV, Work& w) {
    const size_t nn = (size_t)n * n, mn = (size_t)m * n, mm = (size_t)m * m;
    std::memset(K, 0, sizeof(double) * mn * N);
    std::memset(k, 0, sizeof(double) * (size_t)m * N);
    std::memset(Vx, 0, sizeof(double) * (size_t)n * N);
    if (Vxx) std::memset(Vxx, 0, sizeof(double) * nn * N);
    dV[0] = dV[1] = 0.0;
    std::memcpy(Vx + (size_t)n * (N - 1), cx + (size_t)n * (N - 1), sizeof(double) * n);
    std::memcpy(w.V.data(), cxx + cxx_st * (N - 1), sizeof(double) * nn);
    if (Vxx) std::memcpy(Vxx + nn * (N - 1), w.V.data(), sizeof(double) * nn);
    if (Quu_out) std::memcpy(Quu_out + mm * (N - 1), cuu + cuu_st * (N - 1), sizeof(double) * mm);
    const bool use_qp = lims && !(lims[0] > lims[m]);
    std::vector<double> Qx(n), Qu(m), ki(m), kw(m, 0.0), Quuk(m), lo(m), up(m), col(m);
    int diverge = 0;
    for (int i = N - 2; i >= 0; i--) {
        const double* fxi = fx + fx_st * i;
        const double* fui = fu + fu_st * i;
        const double* cxxi = cxx + cxx_st * i;
        const double* cxui = cxu + cxu_st * i;
        const double* cuui = cuu + cuu_st * i;
        const double* Vxn = Vx + (size_t)n * (i + 1);
        double* V = w.V.data();
        for (int a = 0; a < m; a++) { double s = 0; for (int p = 0; p < n; p++) s += fui[p + n * a] * Vxn[p]; Qu[a] = cu[(size_t)m * i + a] + s; }
        for (int r = 0; r < n; r++) { double s = 0; for (int p = 0; p < n; p++) s += fxi[p + n * r] * Vxn[p]; Qx[r] = cx[(size_t)n * i + r] + s; }
        gemm_tn(m, n, n, fui, V, w.T1.data());                       // fu'Vxx
        gemm_nn(m, n, n, w.T1.data(), fxi, w.Qux.data());            // (fu'Vxx) fx
        for (int j = 0; j < n; j++) for (int a = 0; a < m; a++) w.Qux[a + m * j] += cxui[j + n * a];
        gemm_nn(m, m, n, w.T1.data(), fui, w.Quu.data());            // (fu'Vxx) fu
        for (size_t e = 0; e < mm; e++) w.Quu[e] += cuui[e];
        gemm_tn(n, n, n, fxi, V, w.T2.data());                       // fx'Vxx
        gemm_nn(n, n, n, w.T2.data(), fxi, w.Qxx.data());
        for (size_t e = 0; e < nn; e++) w.Qxx[e] += cxxi[e];
        const double* Vr = V;
        if (regType == 2) {
            std::memcpy(w.Vreg.data(), V, sizeof(double) * nn);
            for (int d = 0; d < n; d++) w.Vreg[d + n * d] += lam;
            Vr = w.Vreg.data();
        }
        gemm_tn(m, n, n, fui, Vr, w.T1.data());                      // fu'Vxx_reg (recomputed, as the reference does)
        gemm_nn(m, n, n, w.T1.data(), fxi, w.Quxr.data());
        for (int j = 0; j < n; j++) for (int a = 0; a < m; a++) w.Quxr[a + m * j] += cxui[j + n * a];
        gemm_nn(m, m, n, w.T1.data(), fui, w.QuuF.data());
        for (size_t e = 0; e < mm; e++) w.QuuF[e] += cuui[e];
        if (regType == 1) for (int d = 0; d < m; d++) w.QuuF[d + m * d] += lam;
        // ---- @end_backward_pass
        unsigned fm = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
        int nf = m;
        int idx[32];
        if (!use_qp) {
            for (int a = 0; a < m; a++) idx[a] = a;
            if (!chol_sub(m, w.QuuF.data(), idx, m, w.R.data())) { diverge = i + 1; break; }
            for (int a = 0; a < m; a++) ki[a] = Qu[a];
            chol_solve(m, w.R.data(), m, ki.data());
            for (int a = 0; a < m; a++) ki[a] = -ki[a];
        } else {
            for (int a = 0; a < m; a++) { lo[a] = lims[a] - u[(size_t)m * i + a]; up[a] = lims[m + a] - u[(size_t)m * i + a]; }
            QPOut o = boxqp(m, w.QuuF.data(), Qu.data(), lo.data(), up.data(), kw.data(), ki.data(), w.R.data());
            if (o.result < 1) { diverge = i + 1; break; }
            fm = o.free_mask;
            nf = __builtin_popcount(fm);
        }
        nf = 0;
        for (int a = 0; a < m; a++) if ((fm >> a) & 1u) idx[nf++] = a;
        for (int j = 0; j < n; j++) {
            for (int p = 0; p < nf; p++) col[p] = w.Quxr[idx[p] + m * j];
            if (nf > 0) chol_solve(m, w.R.data(), nf, col.data());
            for (int a = 0; a < m; a++) w.Kt[a + m * j] = 0.0;
            for (int p = 0; p < nf; p++) w.Kt[idx[p] + m * j] = -col[p];
        }
        // ---- value backup
        for (int a = 0; a < m; a++) { double s = 0; for (int c = 0; c < m; c++) s += w.Quu[a + m * c] * ki[c]; Quuk[a] = s; }
        double d0 = 0, d1 = 0;
        for (int a = 0; a < m; a++) { d0 += ki[a] * Qu[a]; d1 += ki[a] * Quuk[a]; }
        dV[0] += d0; dV[1] += 0.5 * d1;
        double* Vxi = Vx + (size_t)n * i;
        for (int r = 0; r < n; r++) {
            double t1 = 0, t2 = 0, t3 = 0;
            for (int a = 0; a < m; a++) { t1 += w.Kt[a + m * r] * Quuk[a]; t2 += w.Kt[a + m * r] * Qu[a]; t3 += w.Qux[a + m * r] * ki[a]; }
            Vxi[r] = ((Qx[r] + t1) + t2) + t3;
        }
        gemm_nn(m, n, m, w.Quu.data(), w.Kt.data(), w.QK.data());
        for (int c = 0; c < n; c++)
            for (int r = 0; r < n; r++) {
                double t1 = 0, t2 = 0, t3 = 0;
                for (int a = 0; a < m; a++) {
                    t1 += w.Kt[a + m * r] * w.QK[a + m * c];
                    t2 += w.Kt[a + m * r] * w.Qux[a + m * c];
                    t3 += w.Qux[a + m * r] * w.Kt[a + m * c];
                }
                w.tmp[r + (size_t)n * c] = ((w.Qxx[r + (size_t)n * c] + t1) + t2) + t3;
            }
        for (int c = 0; c < n; c++)
            for (int r = 0; r < n; r++) V[r + (size_t)n * c] = 0.5 * (w.tmp[r + (size_t)n * c] + w.tmp[c + (size_t)n * r]);
        if (Vxx) std::memcpy(Vxx + nn * i, V, sizeof(double) * nn);
        std::memcpy(K + mn * i, w.Kt.data(), sizeof(double) * mn);
        for (int a = 0; a < m; a++) { k[(size_t)m * i + a] = ki[a]; kw[a] = ki[a]; }
        if (Quu_out) std::memcpy(Quu_out + mm * i, w.Quu.data(), sizeof(double) * mm);
    }
    if (Vxx1) {
        if (diverge > 0) std::memset(Vxx1, 0, sizeof(double) * nn);
        else std::memcpy(Vxx1, w.V.data(), sizeof(double) * nn);
    }
    return diverge;
}

}  // namespace

extern "C" {

int cpu_ref_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// strides in elements, as ddp_tensor; cx/cu/u are dense (n,T,B)/(m,T,B).
void cpu_back_pass(int n, int m, int N, long B, const double* cx, const double* cu, const double* cxx, long cxx_sb,
                   long cxx_st, const double* cxu, long cxu_sb, long cxu_st, const double* cuu, long cuu_sb, long cuu_st,
                   const double* fx, long fx_sb, long fx_st, const double* fu, long fu_sb, long fu_st, const double* lambda,
                   int regType, const double* lims, const double* u, int* diverge, double* K, double* k, double* Vx,
                   double* Vxx, double* Vxx1, double* Quu, double* dV, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        Work w(n, m);
#pragma omp for schedule(dynamic, 1)
        for (long b = 0; b < B; b++) {
            diverge[b] = back_pass_one(
                n, m, N, cx + (size_t)b * n * N, cu + (size_t)b * m * N, cxx + cxx_sb * b, cxx_st, cxu + cxu_sb * b, cxu_st,
                cuu + cuu_sb * b, cuu_st, fx + fx_sb * b, fx_st, fu + fu_sb * b, fu_st, lambda[b], regType, lims,
                u ? u + (size_t)b * m * N : nullptr, K + (size_t)b * m * n * N, k + (size_t)b * m * N, Vx + (size_t)b * n * N,
                Vxx ? Vxx + (size_t)b * n * n * N : nullptr, Vxx1 ? Vxx1 + (size_t)b * n * n : nullptr,
                Quu ? Quu + (size_t)b * m * m * N : nullptr, dV + 2 * b, w);
        }
    }
}

// forward_pass.jl:9-33 with f = A x + B u, cost = ½Σx'Qx + ½Σu'Ru (demo_linear.jl:35-50)
void cpu_forward_pass_linear(int n, int m, int N, long B, const double* K, const double* k, const double* x0,
                             const double* x, const double* u, const double* alpha, const double* lims, const double* A,
                             long A_sb, const double* Bm, long B_sb, const double* Q, const double* R, double* xnew,
                             double* unew, double* cost, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (long b = 0; b < B; b++) {
        const double* Ab = A + A_sb * b;
        const double* Bb = Bm + B_sb * b;
        std::vector<double> xc(x0 + (size_t)b * n, x0 + (size_t)b * n + n), xn(n), dx(n), un(m);
        double c = 0.0;
        for (int t = 0; t < N; t++) {
            double* xo = xnew + ((size_t)b * N + t) * n;
            for (int i = 0; i < n; i++) xo[i] = xc[i];
            for (int a = 0; a < m; a++) un[a] = u[((size_t)b * N + t) * m + a];
            if (K) {
                const double* Kt = K + ((size_t)b * N + t) * m * n;
                for (int i = 0; i < n; i++) dx[i] = xc[i] - x[((size_t)b * N + t) * n + i];
                for (int a = 0; a < m; a++) un[a] = un[a] + k[((size_t)b * N + t) * m + a] * alpha[b];
                for (int a = 0; a < m; a++) { double s = 0; for (int j = 0; j < n; j++) s += Kt[a + m * j] * dx[j]; un[a] = un[a] + s; }
            }
            if (lims) for (int a = 0; a < m; a++) un[a] = clampd(un[a], lims[a], lims[m + a]);
            for (int a = 0; a < m; a++) { if (un[a] != un[a]) un[a] = 0.0; unew[((size_t)b * N + t) * m + a] = un[a]; }
            for (int i = 0; i < n; i++) { double s = 0; for (int j = 0; j < n; j++) s += Q[i + n * j] * xc[j]; c += 0.5 * xc[i] * s; }
            for (int a = 0; a < m; a++) { double s = 0; for (int cc = 0; cc < m; cc++) s += R[a + m * cc] * un[cc]; c += 0.5 * un[a] * s; }
            if (t < N - 1) {
                for (int i = 0; i < n; i++) {
                    double s1 = 0, s2 = 0;
                    for (int j = 0; j < n; j++) s1 += Ab[i + n * j] * xc[j];
                    for (int a = 0; a < m; a++) s2 += Bb[i + n * a] * un[a];
                    xn[i] = s1 + s2;
                }
                xc.swap(xn);
            }
        }
        cost[b] = c;
    }
}

// standalone boxQP (bit-for-bit the oracle's arithmetic)
void cpu_boxqp(int m, long B, const double* H, const double* g, const double* lower, const double* upper, const double* x0,
               double* x, int* result, double* Hfree, unsigned* free_mask, int* nfactor) {
    for (long b = 0; b < B; b++) {
        std::vector<double> Rf((size_t)m * m, 0.0);
        QPOut o = boxqp(m, H + (size_t)b * m * m, g + (size_t)b * m, lower + (size_t)b * m, upper + (size_t)b * m,
                        x0 + (size_t)b * m, x + (size_t)b * m, Rf.data());
        result[b] = o.result;
        free_mask[b] = o.free_mask;
        nfactor[b] = o.nfactor;
        for (int j = 0; j < m; j++)
            for (int i = 0; i < m; i++) Hfree[(size_t)b * m * m + i + m * j] = (i <= j && j < o.nf) ? Rf[i + m * j] : 0.0;
    }
}

}  // extern "C"
