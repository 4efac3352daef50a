// Generic backward sweep: any n <= 64, m <= 16, all branches (Cholesky / boxQP / KL-augmented).
// One CTA (128 threads) walks one trajectory backwards in time with every per-step block resident
// in shared memory.  This is the coverage kernel: the specialised kernels in back_pass_tile.cu
// (n=32,m=8 class, DMMA tiles) and back_pass_small.cu (n<=8, thread-per-trajectory) take the
// benchmarked shapes; everything else lands here.
//
// WPT (warp per trajectory): the same code with one WARP per trajectory -- four trajectories per CTA, each warp on its own
// slice of shared memory, __syncwarp instead of __syncthreads.  For small shapes (n <= 16: the reference's demo_linear /
// test_readme shape n=10, m=2) a phase has ~100 independent elements, so 128 threads only wait for each other
// (ncu at n=10, m=2: 55 % of the stalls on the barrier, 4 CTAs per SM); the arithmetic is the same, element by element.
// The serial factorisation / QP is instantiated for the smallest bound MM in {2,4,8,16} that holds m (the fully unrolled
// MM = 16 code was 25 % instruction-fetch stalls at m = 2).
//
// Replaces back_pass / back_pass_gps of src/backward_pass.jl:162-252, :259-350 (+ macros :3-79).
#include <cstdlib>
#include "boxqp.cuh"

namespace {

constexpr int NT = 128;

struct Smem {
    double *V, *Fx, *Fu, *W, *Z, *Qxx, *Qux, *Quxr, *K, *QK, *Kp, *S;
    double *Quu, *QuuF, *R, *Si, *Inv;
    double *Vx, *Qx, *Qu, *k, *kw, *Quuk, *lo, *up, *kp, *Sik, *VxN;
    int* flags;   // [0] status (0 ok / 1 diverged), [1] free mask, [2] nfree
};

__host__ __device__ inline size_t smem_doubles(int n, int m, bool gps) {
    int ldn = n | 1, ldm = m | 1;
    size_t d = 0;
    d += (size_t)ldn * n * 4;           // V, Fx, W, Qxx
    d += (size_t)ldn * m * 2;           // Fu, Z
    d += (size_t)ldm * n * 4;           // Qux, Quxr, K, QK
    if (gps) d += (size_t)ldm * n * 2;  // Kp, S
    d += (size_t)m * m * 5;             // Quu, QuuF, R, Si, Inv
    d += (size_t)n * 3 + (size_t)m * 8 + 8;
    return d;
}

__device__ inline void carve(double* base, int n, int m, bool gps, Smem& s) {
    int ldn = n | 1, ldm = m | 1;
    double* p = base;
    s.V = p; p += ldn * n;
    s.Fx = p; p += ldn * n;
    s.W = p; p += ldn * n;
    s.Qxx = p; p += ldn * n;
    s.Fu = p; p += ldn * m;
    s.Z = p; p += ldn * m;
    s.Qux = p; p += ldm * n;
    s.Quxr = p; p += ldm * n;
    s.K = p; p += ldm * n;
    s.QK = p; p += ldm * n;
    if (gps) { s.Kp = p; p += ldm * n; s.S = p; p += ldm * n; } else { s.Kp = s.S = nullptr; }
    s.Quu = p; p += m * m;
    s.QuuF = p; p += m * m;
    s.R = p; p += m * m;
    s.Si = p; p += m * m;
    s.Inv = p; p += m * m;
    s.Vx = p; p += n;
    s.Qx = p; p += n;
    s.VxN = p; p += n;
    s.Qu = p; p += m;
    s.k = p; p += m;
    s.kw = p; p += m;
    s.Quuk = p; p += m;
    s.lo = p; p += m;
    s.up = p; p += m;
    s.kp = p; p += m;
    s.Sik = p; p += m;
    s.flags = reinterpret_cast<int*>(p);
}

// general inverse by Gauss-Jordan with partial pivoting (stands for Julia's `inv`, LU based,
// backward_pass.jl:283,346).  A (m x m, ld m) is destroyed; Ainv receives the inverse.
__device__ void inv_gj(int m, double* A, double* Ainv) {
    for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++) Ainv[i + m * j] = (i == j) ? 1.0 : 0.0;
    for (int c = 0; c < m; c++) {
        int piv = c;
        double best = fabs(A[c + m * c]);
        for (int r = c + 1; r < m; r++)
            if (fabs(A[r + m * c]) > best) { best = fabs(A[r + m * c]); piv = r; }
        if (piv != c)
            for (int j = 0; j < m; j++) {
                double t = A[c + m * j]; A[c + m * j] = A[piv + m * j]; A[piv + m * j] = t;
                t = Ainv[c + m * j]; Ainv[c + m * j] = Ainv[piv + m * j]; Ainv[piv + m * j] = t;
            }
        double d = 1.0 / A[c + m * c];
        for (int j = 0; j < m; j++) { A[c + m * j] *= d; Ainv[c + m * j] *= d; }
        for (int r = 0; r < m; r++) {
            if (r == c) continue;
            double f = A[r + m * c];
            if (f != 0.0)
                for (int j = 0; j < m; j++) { A[r + m * j] -= f * A[c + m * j]; Ainv[r + m * j] -= f * Ainv[c + m * j]; }
        }
    }
}

// factorisation / QP of one step by one thread (fixed sequential order shared with the oracle): returns 1 when the sweep diverges
template <int MM>
__device__ __noinline__ int factor_step(const int m, const bool use_qp, const QPOpts& qp, const double* QuuF, const double* Qu,
                                        const double* lo, const double* up, const double* kw, double* k, double* R, int* flags) {
    if (!use_qp) {
        int idx[MM];
        for (int a = 0; a < m; a++) idx[a] = a;
        if (!chol_upper_sub<MM>(QuuF, m, idx, m, R, m)) return 1;
        for (int a = 0; a < m; a++) k[a] = Qu[a];
        chol_solve<MM>(R, m, m, k);
        for (int a = 0; a < m; a++) k[a] = -k[a];
        flags[1] = (m >= 32) ? -1 : (int)((1u << m) - 1u);
        flags[2] = m;
        return 0;
    }
    unsigned fm = 0;
    int nfac = 0;
    const int res = boxqp_seq<MM>(m, QuuF, m, Qu, lo, up, kw, qp, k, R, m, &fm, &nfac);
    flags[1] = (int)fm;
    flags[2] = __popc(fm);
    return res < 1 ? 1 : 0;                                // :50-56
}

// K[free, j] = -R \ (R' \ Qux_reg[free, j]), clamped rows zero (:42 / :57-61)
template <int MM>
__device__ __noinline__ void gain_column(const int m, const unsigned fm, const int nf, const double* R, const double* qcol, double* kcol) {
    double v[MM];
    int p = 0;
    for (int a = 0; a < m; a++)
        if ((fm >> a) & 1u) v[p++] = qcol[a];
    if (nf > 0) chol_solve<MM>(R, m, nf, v);
    p = 0;
    for (int a = 0; a < m; a++) kcol[a] = ((fm >> a) & 1u) ? -v[p++] : 0.0;
}

template <bool GPS, bool WPT>
__global__ void __launch_bounds__(NT, WPT ? 8 : 4) bp_generic_kernel(BackParams P) {
    extern __shared__ double smem_raw[];
    const int n = P.n, m = P.m, N = P.T;
    const int ldn = n | 1, ldm = m | 1;
    constexpr int NTL = WPT ? 32 : NT;                     // threads that share one trajectory
    const int tid = WPT ? (threadIdx.x & 31) : threadIdx.x;
    Smem s;
    carve(smem_raw + (WPT ? (size_t)(threadIdx.x >> 5) * smem_doubles(n, m, GPS) : 0), n, m, GPS, s);
    auto sync = [&]() { if (WPT) __syncwarp(); else __syncthreads(); };
    // element e = tid + k NTL of a column-major block with d rows: (row, column) advanced without a division per element
    const int rc0_n = tid % n, rc1_n = tid / n, rcd0_n = NTL % n, rcd1_n = NTL / n;
    const int rc0_m = tid % m, rc1_m = tid / m, rcd0_m = NTL % m, rcd1_m = NTL / m;
#define FOR_RC(e, r, c, count, d) \
    for (int e = tid, r = rc0_##d, c = rc1_##d; e < (count); e += NTL, r += rcd0_##d, c += rcd1_##d, (r >= d ? (r -= d, ++c) : 0))
    const bool use_qp = (P.lims != nullptr) && !(P.lims[0] > P.lims[m]);   // backward_pass.jl:31
    const long long nn = (long long)n * n, mn = (long long)m * n, mm = (long long)m * m;
    if (P.redo_count && *P.redo_count == 0) return;          // follow-up launch of the tile kernel with nothing handed over

    const long long b_first = WPT ? (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5) : (long long)blockIdx.x;
    const long long b_step = WPT ? (long long)gridDim.x * (NT / 32) : (long long)gridDim.x;
    for (long long b = b_first; b < P.B; b += b_step) {
        if (P.active && !P.active[b]) continue;
        if (P.redo && !P.redo[b]) continue;
        sync();
        const double lam = GPS ? 0.0 : P.lambda[b];
        const double eta = GPS ? P.eta[b] : 1.0;
        double* Kb = P.K + b * (long long)N * mn;
        double* kb = P.k + b * (long long)N * m;
        double* Vxb = P.Vx + b * (long long)N * n;
        double* Vxxb = P.Vxx ? P.Vxx + b * (long long)N * nn : nullptr;
        double* Quub = P.Quu ? P.Quu + b * (long long)N * mm : nullptr;
        double* Quuib = (GPS && P.Quui) ? P.Quui + b * (long long)N * mm : nullptr;
        const long long trn = (long long)n * (n + 1) / 2, trm = (long long)m * (m + 1) / 2;      // packed upper triangles
        double* Vtb = P.Vxx_tri ? P.Vxx_tri + b * (long long)N * trn : nullptr;
        double* Qtb = P.Quu_tri ? P.Quu_tri + b * (long long)N * trm : nullptr;
        double* Qitb = (GPS && P.Quui_tri) ? P.Quui_tri + b * (long long)N * trm : nullptr;
        // ---- terminal step (backward_pass.jl:21-23 / :280-283)
        {
            const double* cxN = tp(P.cx, b, N - 1);
            const double* cxxN = tp(P.cxx, b, N - 1);
            const double* cuuN = tp(P.cuu, b, N - 1);
            for (int i = tid; i < n; i += NTL) { s.Vx[i] = cxN[i]; Vxb[(long long)(N - 1) * n + i] = cxN[i]; }
            FOR_RC(e, i, j, n * n, n) {
                double v = cxxN[e];
                s.V[i + ldn * j] = v;
                if (Vxxb) Vxxb[(long long)(N - 1) * nn + e] = v;
                if (Vtb && i <= j) Vtb[(long long)(N - 1) * trn + (long long)j * (j + 1) / 2 + i] = v;
            }
            for (int e = tid; e < m * n; e += NTL) Kb[(long long)(N - 1) * mn + e] = 0.0;
            for (int a = tid; a < m; a += NTL) { kb[(long long)(N - 1) * m + a] = 0.0; s.kw[a] = 0.0; }
            if (GPS) {
                const double* SiN = tp(P.Sip, b, N - 1);
                for (int e = tid; e < m * m; e += NTL) {
                    double v = cuuN[e] / eta + SiN[e];
                    s.Quu[e] = v;
                    s.QuuF[e] = v;
                    if (Quub) Quub[(long long)(N - 1) * mm + e] = v;
                    if (Qtb && (e % m) <= (e / m)) Qtb[(long long)(N - 1) * trm + (long long)(e / m) * (e / m + 1) / 2 + (e % m)] = v;
                }
                sync();
                if (tid == 0) inv_gj(m, s.QuuF, s.Inv);
                sync();
                for (int e = tid; e < m * m; e += NTL) {
                    if (Quuib) Quuib[(long long)(N - 1) * mm + e] = s.Inv[e];
                    if (Qitb && (e % m) <= (e / m)) Qitb[(long long)(N - 1) * trm + (long long)(e / m) * (e / m + 1) / 2 + (e % m)] = s.Inv[e];
                }
            } else if (Quub || Qtb) {
                for (int e = tid; e < m * m; e += NTL) {
                    if (Quub) Quub[(long long)(N - 1) * mm + e] = cuuN[e];
                    if (Qtb && (e % m) <= (e / m)) Qtb[(long long)(N - 1) * trm + (long long)(e / m) * (e / m + 1) / 2 + (e % m)] = cuuN[e];
                }
            }
            if (tid == 0) s.flags[0] = 0;
        }
        double dV0 = 0.0, dV1 = 0.0;   // thread 0 only
        int diverge = 0;
        const bool lti = (P.fx.st == 0 && P.fu.st == 0);
        for (int i = N - 2; i >= 0; i--) {
            // ---- stage the step's dynamics
            if (!lti || i == N - 2) {
                const double* fxi = tp(P.fx, b, i);
                const double* fui = tp(P.fu, b, i);
                FOR_RC(e, r_, c_, n * n, n) s.Fx[r_ + ldn * c_] = fxi[e];
                FOR_RC(e, r_, c_, n * m, n) s.Fu[r_ + ldn * c_] = fui[e];
            }
            if (GPS) {
                const double* Kpi = tp(P.Kp, b, i);
                const double* Sii = tp(P.Sip, b, i);
                FOR_RC(e, r_, c_, m * n, m) s.Kp[r_ + ldm * c_] = Kpi[e];
                for (int e = tid; e < m * m; e += NTL) s.Si[e] = Sii[e];
                for (int a = tid; a < m; a += NTL) s.kp[a] = P.kp.p ? tp(P.kp, b, i)[a] : 0.0;
            }
            sync();
            // ---- W = Vxx fx, Z = Vxx fu
            FOR_RC(e, r, c, n * (n + m), n) {
                const double* col = (c < n) ? (s.Fx + ldn * c) : (s.Fu + ldn * (c - n));
                double acc = 0.0;
                for (int q = 0; q < n; q++) acc = fma(s.V[r + ldn * q], col[q], acc);
                if (c < n) s.W[r + ldn * c] = acc; else s.Z[r + ldn * (c - n)] = acc;
            }
            if (GPS) {   // S = Σi K_prev ; Sik = Σi k_prev
                FOR_RC(e, a, j, m * n, m) {
                    double acc = 0.0;
                    for (int q = 0; q < m; q++) acc = fma(s.Si[a + m * q], s.Kp[q + ldm * j], acc);
                    s.S[a + ldm * j] = acc;
                }
                for (int a = tid; a < m; a += NTL) {
                    double acc = 0.0;
                    for (int q = 0; q < m; q++) acc = fma(s.Si[a + m * q], s.kp[q], acc);
                    s.Sik[a] = acc;
                }
            }
            sync();
            // ---- Q expansion (backward_pass.jl:240-247)
            {
                const double* cxi = tp(P.cx, b, i);
                const double* cui = tp(P.cu, b, i);
                const double* cxxi = tp(P.cxx, b, i);
                const double* cxui = tp(P.cxu, b, i);
                const double* cuui = tp(P.cuu, b, i);
                FOR_RC(e, r, c, n * n, n) {          // Qxx = cxx + fx' W
                    double acc = 0.0;
                    for (int q = 0; q < n; q++) acc = fma(s.Fx[q + ldn * r], s.W[q + ldn * c], acc);
                    double v = cxxi[e] + acc;
                    if (!GPS && P.fxx.p) {                        // Qxx[p,q] += sum_k Vx_k fxx[k,q,p]  (backward_pass.jl:118)
                        const double* t3 = tp(P.fxx, b, i) + (long long)n * c + nn * r;
                        double so = 0.0;
                        for (int q = 0; q < n; q++) so = fma(s.Vx[q], t3[q], so);
                        v += so;
                    }
                    if (GPS) {
                        double kl = 0.0;                          // cxxkl = K' Σi K
                        for (int q = 0; q < m; q++) kl = fma(s.Kp[q + ldm * r], s.S[q + ldm * c], kl);
                        v = v / eta + kl;
                    }
                    s.Qxx[r + ldn * c] = v;
                }
                FOR_RC(e, a, j, m * n, m) {          // Qux = cxu' + fu' W
                    double acc = 0.0;
                    for (int q = 0; q < n; q++) acc = fma(s.Fu[q + ldn * a], s.W[q + ldn * j], acc);
                    double v = cxui[j + n * a] + acc;
                    double vr = v;
                    if (GPS) {
                        v = v / eta - s.S[a + ldm * j];           // cxukl = -Σi K
                        vr = v;
                    } else if (P.reg_type == 2) {                 // fu'(Vxx + λI)fx = Qux + λ fu'fx
                        double ff = 0.0;
                        for (int q = 0; q < n; q++) ff = fma(s.Fu[q + ldn * a], s.Fx[q + ldn * j], ff);
                        vr = v + lam * ff;
                    }
                    if (!GPS && P.fxu.p) {                        // Qux[a,j] += sum_k Vx_k fxu[k,j,a]  (:106-109, :121)
                        const double* t3 = tp(P.fxu, b, i) + (long long)n * j + nn * a;
                        double so = 0.0;
                        for (int q = 0; q < n; q++) so = fma(s.Vx[q], t3[q], so);
                        v += so;
                        vr += so;
                    }
                    s.Qux[a + ldm * j] = v;
                    s.Quxr[a + ldm * j] = vr;
                }
                FOR_RC(e, a, c, m * m, m) {          // Quu = cuu + fu' Z
                    double acc = 0.0;
                    for (int q = 0; q < n; q++) acc = fma(s.Fu[q + ldn * a], s.Z[q + ldn * c], acc);
                    double v = cuui[e] + acc;
                    double vf = v;
                    if (GPS) {
                        v = v / eta + s.Si[e];                    // cuukl = Σi
                        vf = v;
                    } else if (P.reg_type == 2) {
                        double ff = 0.0;
                        for (int q = 0; q < n; q++) ff = fma(s.Fu[q + ldn * a], s.Fu[q + ldn * c], ff);
                        vf = v + lam * ff;
                    } else if (P.reg_type == 1 && a == c) {
                        vf = v + lam;
                    }
                    if (!GPS && P.fuu.p) {                        // Quu[a,c] += sum_k Vx_k fuu[k,c,a]  (:112-115, :123)
                        const double* t3 = tp(P.fuu, b, i) + (long long)n * c + (long long)n * m * a;
                        double so = 0.0;
                        for (int q = 0; q < n; q++) so = fma(s.Vx[q], t3[q], so);
                        v += so;
                        vf += so;
                    }
                    s.Quu[e] = v;
                    s.QuuF[e] = vf;
                }
                for (int r = tid; r < n; r += NTL) {              // Qx = cx + fx' Vx
                    double acc = 0.0;
                    for (int q = 0; q < n; q++) acc = fma(s.Fx[q + ldn * r], s.Vx[q], acc);
                    double v = cxi[r] + acc;
                    if (GPS) {
                        double kl = 0.0;                          // cxkl = K' Σi k
                        for (int q = 0; q < m; q++) kl = fma(s.Kp[q + ldm * r], s.Sik[q], kl);
                        v = v / eta + kl;
                    }
                    s.Qx[r] = v;
                }
                for (int a = tid; a < m; a += NTL) {              // Qu = cu + fu' Vx
                    double acc = 0.0;
                    for (int q = 0; q < n; q++) acc = fma(s.Fu[q + ldn * a], s.Vx[q], acc);
                    double v = cui[a] + acc;
                    if (GPS) v = v / eta - s.Sik[a];              // cukl = -Σi k
                    s.Qu[a] = v;
                    if (use_qp) {                                  // :45-46
                        double ui = tp(P.u, b, i)[a];
                        const double* li = P.lims + (long long)i * P.lims_st;      // time-varying limits: block i
                        s.lo[a] = li[a] - ui;
                        s.up[a] = li[m + a] - ui;
                    }
                }
            }
            sync();
            if (GPS) {                                            // Quu = ½(Quu + Quu')  :301
                FOR_RC(e, a, c, m * m, m) {
                    if (a < c) {                                  // each unordered pair owned by one thread
                        double w = 0.5 * (s.Quu[a + m * c] + s.Quu[c + m * a]);
                        s.Quu[a + m * c] = w; s.Quu[c + m * a] = w;
                        s.QuuF[a + m * c] = w; s.QuuF[c + m * a] = w;
                    }
                }
                sync();
            }
            // ---- factorisation / QP (one thread, fixed sequential order shared with the oracle)
            if (tid == 0) {
                int st;
                if (m <= 2) st = factor_step<2>(m, use_qp, P.qp, s.QuuF, s.Qu, s.lo, s.up, s.kw, s.k, s.R, s.flags);
                else if (m <= 4) st = factor_step<4>(m, use_qp, P.qp, s.QuuF, s.Qu, s.lo, s.up, s.kw, s.k, s.R, s.flags);
                else if (m <= 8) st = factor_step<8>(m, use_qp, P.qp, s.QuuF, s.Qu, s.lo, s.up, s.kw, s.k, s.R, s.flags);
                else st = factor_step<DDP_MAX_M>(m, use_qp, P.qp, s.QuuF, s.Qu, s.lo, s.up, s.kw, s.k, s.R, s.flags);
                s.flags[0] = st;
            }
            sync();
            if (s.flags[0] != 0) { diverge = i + 1; break; }
            // ---- gains K (:42 / :57-61): one column per thread
            {
                const unsigned fm = (unsigned)s.flags[1];
                const int nf = s.flags[2];
                for (int j = tid; j < n; j += NTL) {
                    if (m <= 2) gain_column<2>(m, fm, nf, s.R, s.Quxr + ldm * j, s.K + ldm * j);
                    else if (m <= 4) gain_column<4>(m, fm, nf, s.R, s.Quxr + ldm * j, s.K + ldm * j);
                    else if (m <= 8) gain_column<8>(m, fm, nf, s.R, s.Quxr + ldm * j, s.K + ldm * j);
                    else gain_column<DDP_MAX_M>(m, fm, nf, s.R, s.Quxr + ldm * j, s.K + ldm * j);
                }
            }
            sync();
            // ---- QK = Quu K, Quuk = Quu k
            FOR_RC(e, a, j, m * n, m) {
                double acc = 0.0;
                for (int q = 0; q < m; q++) acc = fma(s.Quu[a + m * q], s.K[q + ldm * j], acc);
                s.QK[a + ldm * j] = acc;
            }
            for (int a = tid; a < m; a += NTL) {
                double acc = 0.0;
                for (int q = 0; q < m; q++) acc = fma(s.Quu[a + m * q], s.k[q], acc);
                s.Quuk[a] = acc;
            }
            sync();
            // ---- value backup (:64-72)
            if (tid == 0) {
                double a0 = 0.0, a1 = 0.0;
                for (int a = 0; a < m; a++) { a0 = fma(s.k[a], s.Qu[a], a0); a1 = fma(s.k[a], s.Quuk[a], a1); }
                dV0 += a0;
                dV1 += 0.5 * a1;
            }
            for (int r = tid; r < n; r += NTL) {
                double t1 = 0.0, t2 = 0.0, t3 = 0.0;
                for (int q = 0; q < m; q++) {
                    t1 = fma(s.K[q + ldm * r], s.Quuk[q], t1);
                    t2 = fma(s.K[q + ldm * r], s.Qu[q], t2);
                    t3 = fma(s.Qux[q + ldm * r], s.k[q], t3);
                }
                s.VxN[r] = ((s.Qx[r] + t1) + t2) + t3;
            }
            FOR_RC(e, r, c, n * n, n) {
                double t1 = 0.0, t2 = 0.0, t3 = 0.0;
                for (int q = 0; q < m; q++) {
                    t1 = fma(s.K[q + ldm * r], s.QK[q + ldm * c], t1);
                    t2 = fma(s.K[q + ldm * r], s.Qux[q + ldm * c], t2);
                    t3 = fma(s.Qux[q + ldm * r], s.K[q + ldm * c], t3);
                }
                s.W[r + ldn * c] = ((s.Qxx[r + ldn * c] + t1) + t2) + t3;
            }
            if (GPS && tid == (WPT ? 1 : 32)) {                               // Σ = inv(Quu)  :346
                for (int e = 0; e < m * m; e++) s.QuuF[e] = s.Quu[e];
                inv_gj(m, s.QuuF, s.Inv);
            }
            sync();
            // ---- symmetrise, commit, store
            FOR_RC(e, r, c, n * n, n) {
                double v = 0.5 * (s.W[r + ldn * c] + s.W[c + ldn * r]);
                s.V[r + ldn * c] = v;
                if (Vxxb) Vxxb[(long long)i * nn + e] = v;
                if (Vtb && r <= c) Vtb[(long long)i * trn + (long long)c * (c + 1) / 2 + r] = v;
            }
            for (int r = tid; r < n; r += NTL) { s.Vx[r] = s.VxN[r]; Vxb[(long long)i * n + r] = s.VxN[r]; }
            FOR_RC(e, r_, c_, m * n, m) Kb[(long long)i * mn + e] = s.K[r_ + ldm * c_];
            for (int a = tid; a < m; a += NTL) { kb[(long long)i * m + a] = s.k[a]; s.kw[a] = s.k[a]; }
            if (Quub || Qtb || Quuib || Qitb) FOR_RC(e, r_, c_, m * m, m) {
                const long long te = (long long)c_ * (c_ + 1) / 2 + r_;
                const bool up = r_ <= c_;
                if (Quub) Quub[(long long)i * mm + e] = s.Quu[e];
                if (Qtb && up) Qtb[(long long)i * trm + te] = s.Quu[e];
                if (Quuib) Quuib[(long long)i * mm + e] = s.Inv[e];
                if (Qitb && up) Qitb[(long long)i * trm + te] = s.Inv[e];
            }
            sync();
        }
        // ---- epilogue
        if (diverge > 0) {
            // the reference returns with everything below the failed step still zero (quirk Q10)
            const int upto = diverge - 1;   // 0-based failed step; steps 0..upto stay zero
            for (long long e = tid; e < (long long)(upto + 1) * mn; e += NTL) Kb[e] = 0.0;
            for (long long e = tid; e < (long long)(upto + 1) * m; e += NTL) kb[e] = 0.0;
            for (long long e = tid; e < (long long)(upto + 1) * n; e += NTL) Vxb[e] = 0.0;
            if (Vxxb)
                for (long long e = tid; e < (long long)(upto + 1) * nn; e += NTL) Vxxb[e] = 0.0;
            if (Vtb)
                for (long long e = tid; e < (long long)(upto + 1) * trn; e += NTL) Vtb[e] = 0.0;
        }
        if (P.Vxx1)
            for (int e = tid; e < n * n; e += NTL)
                P.Vxx1[b * nn + e] = (diverge > 0) ? 0.0 : s.V[(e % n) + ldn * (e / n)];
        if (tid == 0) {
            P.diverge[b] = diverge;
            P.dV[2 * b] = dV0;
            P.dV[2 * b + 1] = dV1;
        }
    }
}
#undef FOR_RC

}  // namespace

// Hand-over buffers of the specialised kernels (see BackParams::redo): one int counter (16 bytes reserved) + one byte per
// trajectory, kept on the handle; the counter is reset on the handle's stream before the specialised kernel runs.
int prepare_redo(ddp_handle_s* h, BackParams& P) {
    cudaError_t e = cudaSuccess;
    if (h->redo_cap < P.B) {
        if (h->redo) cudaFree(h->redo);
        h->redo = nullptr; h->redo_cap = 0;
        const long long cap = P.B > h->B ? P.B : h->B;
        e = cudaMalloc((void**)&h->redo, (size_t)cap + 16);
        if (e != cudaSuccess) return (int)e;
        h->redo_cap = cap;
    }
    P.redo_count = reinterpret_cast<int*>(h->redo);
    P.redo = h->redo + 16;
    e = cudaMemsetAsync(P.redo_count, 0, sizeof(int), h->stream);
    return (int)e;
}

int launch_back_pass_generic(ddp_handle_s* h, const BackParams& P, bool gps) {
    const size_t per_traj = smem_doubles(P.n, P.m, gps) * sizeof(double);
    if ((long long)per_traj > h->max_smem_optin) return (int)cudaErrorInvalidValue;
    // small shapes: one warp per trajectory, four trajectories per CTA (see the header); DDP_GENERIC_WPT=0 / 1 forces the CTA /
    // the warp form (read at every launch: the tests compare the two)
    const char* env_wpt = getenv("DDP_GENERIC_WPT");
    const bool no_wpt = env_wpt && env_wpt[0] == '0', force_wpt = env_wpt && env_wpt[0] == '1';
    // (a small batch does not fill the SMs either way: there the CTA form's four warps per trajectory give the lower latency --
    // one demo_linear trajectory, whole solve: 89 ms against 133 ms)
    const bool wpt = !no_wpt && P.n <= 16 && (force_wpt || P.B > (long long)h->sm_count * 8) && (long long)(per_traj * (NT / 32)) <= h->max_smem_optin;
    const size_t bytes = wpt ? per_traj * (NT / 32) : per_traj;
    cudaError_t e;
    if (gps) e = wpt ? cudaFuncSetAttribute(bp_generic_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)
                     : cudaFuncSetAttribute(bp_generic_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    else e = wpt ? cudaFuncSetAttribute(bp_generic_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)
                 : cudaFuncSetAttribute(bp_generic_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    int per_sm = (int)((size_t)h->max_smem_optin / (bytes + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16) per_sm = 16;
    long long grid = (long long)h->sm_count * per_sm;
    const long long need = wpt ? (P.B + NT / 32 - 1) / (NT / 32) : P.B;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (gps) {
        if (wpt) bp_generic_kernel<true, true><<<(unsigned)grid, NT, bytes, h->stream>>>(P);
        else bp_generic_kernel<true, false><<<(unsigned)grid, NT, bytes, h->stream>>>(P);
    } else {
        if (wpt) bp_generic_kernel<false, true><<<(unsigned)grid, NT, bytes, h->stream>>>(P);
        else bp_generic_kernel<false, false><<<(unsigned)grid, NT, bytes, h->stream>>>(P);
    }
    h->launches++;
    return (int)cudaGetLastError();
}
