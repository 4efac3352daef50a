// Device-resident iLQG driver and the host-buffer (end-to-end) iteration pipeline.
//
// ddp_ilqg_solve_f64 replaces the outer loop of iLQG (src/iLQG.jl:143-341) for a whole batch:
// every trajectory carries its own {λ, dλ, α index, iter, accepted_iter, status}; the sweeps are
// launched over the batch with activity masks, and the small state-machine kernels below apply
// the reference's rules per trajectory (including quirks Q1, Q5, Q11 of SURVEY.md section 8a).
// The host only reads back three counters per outer iteration.
//
// ddp_ilqg_iter_host_f64 runs one backward + forward sweep on pinned HOST arrays, cutting the
// batch into chunks whose H2D copies, kernels and D2H copies overlap on three streams.
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "ddp_common.cuh"

namespace {

#define CUS(call)                                       \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) { err = e__; goto fail; } \
    } while (0)

// ---------------------------------------------------------------------------------------------
// derivative kernels (STEP 1 of iLQG.jl:225-229 for the built-in models)

// cx = Q (x - goal), cu = R u   (demo_linear.jl:38-39 / system_pendcart.jl:108-112). One warp per (b,t).
__global__ void __launch_bounds__(128) df_cost_kernel(int n, int m, int T, long long B, const double* __restrict__ x,
                                                      const double* __restrict__ u, TensorD Q, TensorD R,
                                                      const double* __restrict__ goal, const unsigned char* __restrict__ mask,
                                                      double* __restrict__ cx, double* __restrict__ cu) {
    __shared__ double sd[4][64 + 16];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long total = B * T;
    for (long long bt = (long long)blockIdx.x * 4 + w; bt < total; bt += (long long)gridDim.x * 4) {
        const long long b = bt / T;
        if (mask && !mask[b]) continue;
        const double* Qm = Q.p + b * Q.sb;
        const double* Rm = R.p + b * R.sb;
        __syncwarp();
        for (int i = lane; i < n; i += 32) sd[w][i] = x[bt * n + i] - (goal ? goal[i] : 0.0);
        for (int a = lane; a < m; a += 32) sd[w][64 + a] = u[bt * m + a];
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            double s = 0.0;
            for (int j = 0; j < n; j++) s = fma(Qm[i + n * j], sd[w][j], s);
            cx[bt * n + i] = s;
        }
        for (int a = lane; a < m; a += 32) {
            double s = 0.0;
            for (int c = 0; c < m; c++) s = fma(Rm[a + m * c], sd[w][64 + c], s);
            cu[bt * m + a] = s;
        }
    }
}

// Same result for cost matrices shared by the batch (the usual case): Q and R are staged once per CTA in shared memory
// (QDIAG: only the diagonal of Q is read, e.g. Q = h*I of demo_linear.jl:18), so the per-(b,t) work is n FMAs per lane
// on operands that never leave the SM.  One warp per (b,t).
template <bool QDIAG>
__global__ void __launch_bounds__(256) df_cost_shared_kernel(int n, int m, int T, long long B, const double* __restrict__ x,
                                                             const double* __restrict__ u, const double* __restrict__ Q,
                                                             const double* __restrict__ R, const double* __restrict__ goal,
                                                             const unsigned char* __restrict__ mask, double* __restrict__ cx,
                                                             double* __restrict__ cu) {
    extern __shared__ double dsm[];
    double* sQ = dsm;                                   // n x n (or n diagonal entries)
    double* sR = sQ + (QDIAG ? n : n * n);              // m x m
    double* sg = sR + m * m;                            // goal (n)
    double* sd = sg + n;                                // per warp: d (n) | u (m)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int e = threadIdx.x; e < (QDIAG ? n : n * n); e += blockDim.x) sQ[e] = QDIAG ? Q[e * (n + 1)] : Q[e];
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) sR[e] = R[e];
    for (int e = threadIdx.x; e < n; e += blockDim.x) sg[e] = goal ? goal[e] : 0.0;
    __syncthreads();
    double* d = sd + w * (n + m);
    const long long total = B * T;
    for (long long bt = (long long)blockIdx.x * nw + w; bt < total; bt += (long long)gridDim.x * nw) {
        if (mask && !mask[bt / T]) continue;
        __syncwarp();
        for (int i = lane; i < n; i += 32) d[i] = x[bt * n + i] - sg[i];
        for (int a = lane; a < m; a += 32) d[n + a] = u[bt * m + a];
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            double s = 0.0;
            if (QDIAG) s = sQ[i] * d[i];
            else
                for (int j = 0; j < n; j++) s = fma(sQ[i + n * j], d[j], s);
            cx[bt * n + i] = s;
        }
        for (int a = lane; a < m; a += 32) {
            double s = 0.0;
            for (int c = 0; c < m; c++) s = fma(sR[a + m * c], d[n + c], s);
            cu[bt * m + a] = s;
        }
    }
}

// Small systems (n <= 8): one THREAD per output element; a warp-per-(b,t) mapping would leave 32 - n lanes idle.
__global__ void __launch_bounds__(256) df_cost_small_kernel(int n, int m, int T, long long B, const double* __restrict__ x,
                                                            const double* __restrict__ u, const double* __restrict__ Q,
                                                            const double* __restrict__ R, const double* __restrict__ goal,
                                                            const unsigned char* __restrict__ mask, double* __restrict__ cx,
                                                            double* __restrict__ cu) {
    __shared__ double sQ[64], sR[64], sg[8];
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) sQ[e] = Q[e];
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) sR[e] = R[e];
    for (int e = threadIdx.x; e < n; e += blockDim.x) sg[e] = goal ? goal[e] : 0.0;
    __syncthreads();
    const long long nx = B * T * n, nu = B * T * m, stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nx; idx += stride) {
        const long long bt = idx / n;
        const int i = (int)(idx - bt * n);
        if (mask && !mask[bt / T]) continue;
        double s = 0.0;
        for (int j = 0; j < n; j++) s = fma(sQ[i + n * j], x[bt * n + j] - sg[j], s);
        cx[idx] = s;
    }
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nu; idx += stride) {
        const long long bt = idx / m;
        const int a = (int)(idx - bt * m);
        if (mask && !mask[bt / T]) continue;
        double s = 0.0;
        for (int c = 0; c < m; c++) s = fma(sR[a + m * c], u[bt * m + c], s);
        cu[idx] = s;
    }
}

__device__ __forceinline__ void mat3_mul(const double* A, const double* Bm, double* C) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++) s = fma(A[i * 3 + k], Bm[k * 3 + j], s);
            C[i * 3 + j] = s;
        }
}

// ZoH-discretised Jacobians of the pendulum on a cart: [fx fu; 0 1] = exp(h [fxc fuc; 0 0])
// (system_pendcart.jl:137-151).  The 5 x 5 generator is block diagonal up to the shared input column:
//   (x1, x2, u):  h [0 1 0; a -d b; 0 0 0],  a = -g/l cos x1 - u/l sin x1,  b = cos x1 / l
//   (x3, x4, u):  h [0 1 0; 0 0 1; 0 0 0]    nilpotent: exp = [1 h h^2/2; 0 1 h; 0 0 1] exactly
// so only a 3 x 3 exponential is computed: scaling-and-squaring (2^-4) with a degree-10 Taylor polynomial; the scaled
// norm is < 0.02, so the truncation error is below 1e-20.  One thread per (b,t).
__global__ void __launch_bounds__(128) df_pendcart_kernel(int T, long long B, const double* __restrict__ x,
                                                          const double* __restrict__ u, double g, double l, double h, double d,
                                                          const unsigned char* __restrict__ mask, double* __restrict__ fx,
                                                          double* __restrict__ fu) {
    const long long bt_raw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (bt_raw - (threadIdx.x & 31) >= B * T) return;               // whole warp out of range
    const bool inrange = bt_raw < B * T;
    const long long bt = inrange ? bt_raw : B * T - 1;              // tail lanes shadow the last item: the write-out below is cooperative
    const bool act = inrange && !(mask && !mask[bt / T]);
    double sn, cs;
    sincos(x[bt * 4], &sn, &cs);
    const double uu = u[bt];
    double M[9], P[9], S[9], Tm[9];
#pragma unroll
    for (int i = 0; i < 9; i++) M[i] = 0.0;
    const double sc = h / 16.0;
    M[0 * 3 + 1] = sc;                                   // fxc[1,2] = 1
    M[1 * 3 + 0] = sc * (-g / l * cs - uu / l * sn);     // fxc[2,1]
    M[1 * 3 + 1] = sc * (-d);                            // fxc[2,2]
    M[1 * 3 + 2] = sc * (cs / l);                        // fuc[2]
    // S = I + M + M^2/2! + ... + M^10/10!
#pragma unroll
    for (int i = 0; i < 9; i++) { S[i] = M[i] + ((i % 4 == 0) ? 1.0 : 0.0); P[i] = M[i]; }
#pragma unroll
    for (int k = 2; k <= 10; k++) {
        mat3_mul(P, M, Tm);
        const double inv = 1.0 / (double)k;
#pragma unroll
        for (int i = 0; i < 9; i++) { P[i] = Tm[i] * inv; S[i] += P[i]; }
    }
#pragma unroll
    for (int s = 0; s < 4; s++) {
        mat3_mul(S, S, Tm);
#pragma unroll
        for (int i = 0; i < 9; i++) S[i] = Tm[i];
    }
    // column-major outputs fx[i + 4 j], fu[i]: the warp's 32 blocks are one contiguous 4 KB (fx) / 1 KB (fu) run of
    // global memory, so they go through padded shared-memory rows and leave as fully coalesced 16-byte stores
    // (a thread storing its own 128-byte block makes every store instruction touch 32 lines).
    __shared__ __align__(16) double s_out[4][32 * 18 + 32 * 4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* row = s_out[w] + lane * 18;
    double* frow = s_out[w] + 32 * 18 + lane * 4;
    auto put2 = [](double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); };
    put2(row + 0, S[0], S[3]);   put2(row + 2, 0.0, 0.0);        // column 1: (S00, S10, 0, 0)
    put2(row + 4, S[1], S[4]);   put2(row + 6, 0.0, 0.0);        // column 2: (S01, S11, 0, 0)
    put2(row + 8, 0.0, 0.0);     put2(row + 10, 1.0, 0.0);       // column 3: (0, 0, 1, 0)
    put2(row + 12, 0.0, 0.0);    put2(row + 14, h, 1.0);         // column 4: (0, 0, h, 1)
    put2(frow, S[2], S[5]);      put2(frow + 2, 0.5 * h * h, h);
    __syncwarp();
    const long long bt0 = bt_raw - lane;                          // first (b,t) of the warp
    const bool vec = (((uintptr_t)fx % 16) == 0) && (((uintptr_t)fu % 16) == 0) && (mask == nullptr);
    if (vec) {
        const long long total = B * T;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int e = k * 32 + lane, r = e >> 3, c = e & 7;
            if (bt0 + r < total) *reinterpret_cast<double2*>(fx + bt0 * 16 + 2 * e) = *reinterpret_cast<const double2*>(s_out[w] + r * 18 + 2 * c);
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int e = k * 32 + lane, r = e >> 1, c = e & 1;
            if (bt0 + r < total) *reinterpret_cast<double2*>(fu + bt0 * 4 + 2 * e) = *reinterpret_cast<const double2*>(s_out[w] + 32 * 18 + r * 4 + 2 * c);
        }
    } else if (act) {
#pragma unroll
        for (int i = 0; i < 16; i++) fx[bt * 16 + i] = row[i];
#pragma unroll
        for (int i = 0; i < 4; i++) fu[bt * 4 + i] = frow[i];
    }
}

// ---------------------------------------------------------------------------------------------
// state-machine kernels

struct SolveState {
    double *lambda, *dlambda, *cost, *costnew, *alpha, *gnorm, *dV, *last_dcost, *ratio;
    int *aidx, *status, *iter, *acc, *diverge, *bpr;
    ddp_ilqg_trace* trace;     // (trace_cap, B) or nullptr
    int trace_cap;
    long long B;
    unsigned char *active, *need_bp, *bp_ok, *need_fwd, *accepted;
    int* counters;    // [0] bp retries, [1] still searching, [2] still active, [3] init pending
};

struct SolveOpts {
    int n_alpha;
    double alpha[16];
    double tol_fun, tol_grad, lam_factor, lam_max, lam_min, reduce_ratio_min;
    int max_iter;
};

// after a backward launch: apply iLQG.jl:244-249 to the trajectories that diverged
__global__ void bp_retry_kernel(long long B, SolveState s, SolveOpts o) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || !s.need_bp[b]) return;
    if (s.diverge[b] > 0) {
        const double lam = s.lambda[b], dl = s.dlambda[b];
        s.bpr[b] += 1;
        s.dlambda[b] = fmax(dl * o.lam_factor, o.lam_factor);     // tuple assignment: λ uses the OLD dλ (Q1)
        const double ln = fmax(lam * dl, o.lam_min);
        s.lambda[b] = ln;
        if (ln > o.lam_max) { s.need_bp[b] = 0; s.bp_ok[b] = 0; }
        else atomicAdd(&s.counters[0], 1);
    } else {
        s.need_bp[b] = 0;
        s.bp_ok[b] = 1;
    }
}

// g_norm = mean_t max_j |k|/(|u|+1)  (iLQG.jl:256) ; success test (:258).  One warp per trajectory.
__global__ void __launch_bounds__(128) gnorm_kernel(int m, int T, long long B, const double* __restrict__ k,
                                                    const double* __restrict__ u, SolveState s, SolveOpts o) {
    const int lane = threadIdx.x & 31;
    long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B || !s.active[b]) return;
    double acc = 0.0;
    for (int t = lane; t < T; t += 32) {
        double mx = 0.0;
        bool isnan_ = false;
        for (int a = 0; a < m; a++) {
            double v = fabs(k[(b * T + t) * m + a]) / (fabs(u[(b * T + t) * m + a]) + 1.0);
            if (v != v) isnan_ = true;
            mx = fmax(mx, v);
        }
        acc += isnan_ ? nan("") : mx;
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
        const double gn = acc / (double)T;
        s.gnorm[b] = gn;
        s.ratio[b] = 0.0;                                      // reduce_ratio = 0. at the top of every iteration (iLQG.jl:223)
        if (s.trace && s.iter[b] - 1 < s.trace_cap) s.trace[(long long)(s.iter[b] - 1) * s.B + b].grad_norm = gn;   // :257
        if (gn < o.tol_grad && s.lambda[b] < 1e-5) {          // SUCCESS: gradient norm < tol_grad
            s.status[b] = 0;
            s.active[b] = 0;
            s.need_fwd[b] = 0;
            s.accepted[b] = 0;
        } else {
            s.need_fwd[b] = s.bp_ok[b];
            s.aidx[b] = 0;
            s.alpha[b] = o.alpha[0];
            s.accepted[b] = 0;
        }
    }
}

// after a forward launch with alpha[b]: the serial backtracking test of iLQG.jl:267-281
__global__ void linesearch_kernel(long long B, SolveState s, SolveOpts o) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || !s.need_fwd[b]) return;
    const double a = s.alpha[b];
    const double dcost = s.cost[b] - s.costnew[b];
    const double expected = -a * (s.dV[2 * b] + a * s.dV[2 * b + 1]);
    double ratio;
    if (expected > 0) ratio = dcost / expected;
    else ratio = (dcost > 0) ? 1.0 : ((dcost < 0) ? -1.0 : dcost);     // sign(Δcost)
    s.last_dcost[b] = dcost;
    s.ratio[b] = ratio;
    if (ratio > o.reduce_ratio_min) {
        s.accepted[b] = 1;
        s.need_fwd[b] = 0;
    } else {
        const int ni = s.aidx[b] + 1;
        if (ni >= o.n_alpha) {
            s.need_fwd[b] = 0;                                  // line search exhausted
        } else {
            s.aidx[b] = ni;
            s.alpha[b] = o.alpha[ni];
            atomicAdd(&s.counters[1], 1);
        }
    }
}

// the same serial test over the costs of ALL remaining step sizes (one multi-alpha launch): the first alpha in
// order with ratio > reduce_ratio_min wins, exactly as the loop of iLQG.jl:267-281 would find it
__global__ void linesearch_multi_kernel(long long B, SolveState s, SolveOpts o, const double* __restrict__ costs /* (n_alpha-1, B) */) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || !s.need_fwd[b]) return;
    bool found = false;
    for (int ai = 1; ai < o.n_alpha; ai++) {
        const double a = o.alpha[ai];
        const double dcost = s.cost[b] - costs[(long long)(ai - 1) * B + b];
        const double expected = -a * (s.dV[2 * b] + a * s.dV[2 * b + 1]);
        double ratio;
        if (expected > 0) ratio = dcost / expected;
        else ratio = (dcost > 0) ? 1.0 : ((dcost < 0) ? -1.0 : dcost);
        s.last_dcost[b] = dcost;
        s.ratio[b] = ratio;
        s.aidx[b] = ai;
        s.alpha[b] = a;
        if (ratio > o.reduce_ratio_min) { found = true; break; }
    }
    if (found) { s.accepted[b] = 1; atomicAdd(&s.counters[1], 1); }     // need_fwd stays set: the accepted step is rolled out next
    else s.need_fwd[b] = 0;                                             // line search exhausted
}

__global__ void clear_flags_kernel(long long B, unsigned char* f) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) f[b] = 0;
}

// STEP 4 (iLQG.jl:293-323) scalars; the array copies are done by copy_accepted_kernel
__global__ void accept_kernel(long long B, SolveState s, SolveOpts o) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || !s.active[b]) return;
    const bool was_accepted = s.accepted[b] != 0;
    s.accepted[b] = 0;                                          // the flag never outlives its iteration
    if (was_accepted) {
        const double dl = fmin(s.dlambda[b] / o.lam_factor, 1.0 / o.lam_factor);   // :299
        s.dlambda[b] = dl;
        s.lambda[b] = fmax(s.lambda[b] * dl, o.lam_min);                           // :300 (new dλ)
        s.cost[b] = s.costnew[b];
        if (s.last_dcost[b] < o.tol_fun) {                      // SUCCESS: cost change < tol_fun (break before iter += 1)
            s.status[b] = 1;
            s.active[b] = 0;
            return;
        }
        s.acc[b] += 1;
    } else {
        const double lam = s.lambda[b], dl = s.dlambda[b];
        s.dlambda[b] = fmax(dl * o.lam_factor, o.lam_factor);
        const double ln = fmax(lam * dl, o.lam_min);
        s.lambda[b] = ln;
        if (ln > o.lam_max) {                                   // EXIT: λ > λmax
            s.status[b] = 2;
            s.active[b] = 0;
            return;
        }
    }
    if (s.trace && s.iter[b] - 1 < s.trace_cap) {               // "update trace" (iLQG.jl:324-330): not reached by the two breaks above
        ddp_ilqg_trace& r = s.trace[(long long)(s.iter[b] - 1) * s.B + b];
        r.lambda = s.lambda[b]; r.dlambda = s.dlambda[b]; r.cost = s.cost[b];
        r.alpha = was_accepted ? s.alpha[b] : nan("");          // αi = NaN on a rejected iteration (:312)
        r.improvement = s.last_dcost[b]; r.reduce_ratio = s.ratio[b];
        r.accepted = was_accepted ? 1 : 0; r.bp_retries = s.bpr[b];
    }
    s.bpr[b] = 0;
    s.iter[b] += 1;
    if (s.acc[b] > o.max_iter) {                                // while accepted_iter <= max_iter
        s.status[b] = 3;
        s.active[b] = 0;
        return;
    }
    s.need_bp[b] = 1;
    atomicAdd(&s.counters[2], 1);
}

// x,u <- xnew,unew ; traj_new.k <- u (Q11) for accepted trajectories.  One CTA per trajectory slice.
__global__ void __launch_bounds__(256) copy_accepted_kernel(int n, int m, int T, long long B, const unsigned char* __restrict__ accepted,
                                                            const double* __restrict__ xnew, const double* __restrict__ unew,
                                                            double* __restrict__ x, double* __restrict__ u, double* __restrict__ k) {
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        if (!accepted[b]) continue;
        const long long ox = b * T * n, ou = b * T * m;
        for (int e = threadIdx.x; e < T * n; e += blockDim.x) x[ox + e] = xnew[ox + e];
        for (int e = threadIdx.x; e < T * m; e += blockDim.x) { const double v = unew[ou + e]; u[ou + e] = v; k[ou + e] = v; }
    }
}

// initial rollout test all(|x| < 1e8) (iLQG.jl:187).  One warp per trajectory.
__global__ void __launch_bounds__(128) init_check_kernel(int n, int m, int T, long long B, const double* __restrict__ xnew,
                                                         const double* __restrict__ unew, const double* __restrict__ costnew,
                                                         double* __restrict__ x, double* __restrict__ u, SolveState s, SolveOpts o) {
    const int lane = threadIdx.x & 31;
    long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B || !s.need_fwd[b]) return;
    bool ok = true;
    for (int e = lane; e < T * n; e += 32) {
        const double v = fabs(xnew[b * T * n + e]);
        if (!(v < 1e8)) ok = false;
    }
    ok = __all_sync(0xffffffffu, ok);
    if (ok) {
        for (int e = lane; e < T * n; e += 32) x[b * T * n + e] = xnew[b * T * n + e];
        for (int e = lane; e < T * m; e += 32) u[b * T * m + e] = unew[b * T * m + e];
        if (lane == 0) { s.cost[b] = costnew[b]; s.need_fwd[b] = 0; s.need_bp[b] = 1; }
    } else if (lane == 0) {
        const int ni = s.aidx[b] + 1;
        if (ni >= o.n_alpha) {                                   // EXIT: initial control sequence caused divergence
            s.need_fwd[b] = 0;
            s.status[b] = 4;
            s.active[b] = 0;
        } else {
            s.aidx[b] = ni;
            s.alpha[b] = o.alpha[ni];
            atomicAdd(&s.counters[3], 1);
        }
    }
}

__global__ void init_state_kernel(long long B, SolveState s, double lam, double dlam, double alpha0) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    s.lambda[b] = lam; s.dlambda[b] = dlam; s.cost[b] = 0.0; s.costnew[b] = 0.0; s.alpha[b] = alpha0; s.gnorm[b] = nan("");
    s.dV[2 * b] = s.dV[2 * b + 1] = 0.0; s.last_dcost[b] = 0.0; s.ratio[b] = 0.0;
    s.aidx[b] = 0; s.status[b] = -1; s.iter[b] = 1; s.acc[b] = 1; s.diverge[b] = 0; s.bpr[b] = 0;
    s.active[b] = 1; s.need_bp[b] = 0; s.bp_ok[b] = 0; s.need_fwd[b] = 1; s.accepted[b] = 0;
}

// pre-rolled start (iLQG.jl:193-197): x = x0 (n,N), cost given, no initial rollout
__global__ void init_prerolled_kernel(long long B, SolveState s, const double* __restrict__ cost_init) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    s.cost[b] = cost_init[b];
    s.need_fwd[b] = 0;
    s.need_bp[b] = 1;
}

// the outer loop's safety cap was reached: whoever is still running is reported as such
__global__ void mark_incomplete_kernel(long long B, SolveState s) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B && s.active[b]) s.status[b] = 6;
}

__global__ void export_state_kernel(long long B, SolveState s, ddp_ilqg_state* out) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    ddp_ilqg_state r;
    r.lambda = s.lambda[b]; r.dlambda = s.dlambda[b]; r.cost = s.cost[b]; r.g_norm = s.gnorm[b];
    r.last_dcost = s.last_dcost[b]; r.last_alpha = s.alpha[b];
    r.iter = s.iter[b]; r.accepted_iter = s.acc[b]; r.status = s.status[b]; r.pad = 0;
    out[b] = r;
}

// Device scratch of one call.  With a handle, allocations are bump-allocated from the handle's workspace arena, which
// is kept between calls (cudaMalloc/cudaFree of two dozen buffers cost several hundred ms per solve otherwise); what
// does not fit falls back to cudaMalloc and the arena is grown for the next call.
struct DevBuf {
    std::vector<void*> ptrs;
    ddp_handle_s* h = nullptr;
    size_t off = 0, need = 0;
    DevBuf() = default;
    explicit DevBuf(ddp_handle_s* handle) : h(handle) {}
    cudaError_t reserve(size_t bytes) {                  // make the arena at least this large before the first alloc
        if (!h || bytes <= h->ws_cap) return cudaSuccess;
        if (h->ws) { cudaFree(h->ws); h->ws = nullptr; h->ws_cap = 0; }
        cudaError_t e = cudaMalloc(&h->ws, bytes);
        if (e == cudaSuccess) h->ws_cap = bytes;
        else cudaGetLastError();                         // not fatal: alloc() falls back to cudaMalloc
        return cudaSuccess;
    }
    template <typename T>
    cudaError_t alloc(T** p, size_t count) {
        const size_t bytes = (std::max<size_t>(count * sizeof(T), 16) + 255) & ~(size_t)255;
        need += bytes;
        if (h && h->ws && off + bytes <= h->ws_cap) {
            *p = (T*)((char*)h->ws + off);
            off += bytes;
            return cudaSuccess;
        }
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, bytes);
        if (e == cudaSuccess) { ptrs.push_back(q); *p = (T*)q; }
        return e;
    }
    ~DevBuf() {
        for (void* p : ptrs) cudaFree(p);
        if (h && need > h->ws_cap) reserve(need);
    }
};

// cx = Q (x - goal), cu = R u over the batch: shared-matrix kernel when Q, R are shared by the batch, else the general one
void launch_df_cost(ddp_handle_s* h, cudaStream_t st, int n, int m, int T, long long B, const double* x, const double* u, const TensorD& Q,
                    const TensorD& R, const double* goal, const unsigned char* mask, double* cx, double* cu, bool qdiag) {
    const long long wtot = B * T;
    if (Q.sb == 0 && R.sb == 0 && n <= 8 && m <= 8) {
        const unsigned grid = (unsigned)std::min<long long>((wtot * n + 255) / 256, (long long)h->sm_count * 16);
        df_cost_small_kernel<<<grid, 256, 0, st>>>(n, m, T, B, x, u, Q.p, R.p, goal, mask, cx, cu);
    } else if (Q.sb == 0 && R.sb == 0) {
        const int nw = 8;
        const size_t bytes = ((size_t)(qdiag ? n : n * n) + (size_t)m * m + n + (size_t)nw * (n + m)) * sizeof(double);
        const unsigned grid = (unsigned)std::min<long long>((wtot + nw - 1) / nw, (long long)h->sm_count * 8);
        if (qdiag) df_cost_shared_kernel<true><<<grid, nw * 32, bytes, st>>>(n, m, T, B, x, u, Q.p, R.p, goal, mask, cx, cu);
        else df_cost_shared_kernel<false><<<grid, nw * 32, bytes, st>>>(n, m, T, B, x, u, Q.p, R.p, goal, mask, cx, cu);
    } else {
        const unsigned grid = (unsigned)std::min<long long>((wtot + 3) / 4, (long long)h->sm_count * 16);
        df_cost_kernel<<<grid, 128, 0, st>>>(n, m, T, B, x, u, Q, R, goal, mask, cx, cu);
    }
    h->launches++;
}

int run_back(ddp_handle_s* h, const BackParams& P) {
    bool handled = false;
    int rc = 0;
    if (!(h->flags & 1u)) {
        rc = launch_back_pass_tile(h, P, false, &handled);
        if (!handled && rc == 0) rc = launch_back_pass_small(h, P, false, &handled);
    }
    if (!handled && rc == 0) rc = launch_back_pass_generic(h, P, false);
    return rc;
}

int run_fwd(ddp_handle_s* h, const FwdParams& P) {
    bool handled = false;
    int rc = 0;
    if (!(h->flags & 1u)) rc = launch_forward_fast(h, P, &handled);
    if (!handled && rc == 0) rc = launch_forward_generic(h, P);
    return rc;
}

}  // namespace

extern "C" {

// STEP 1 of iLQG.jl:225-229 for the built-in models: cx = Q(x - goal), cu = R u and, for the pendulum
// on a cart, the ZoH-discretised Jacobians fx, fu (system_pendcart.jl:137-154).  For the linear model
// the Jacobians are A and B themselves, so fx/fu must be NULL.
int ddp_model_derivs_f64(ddp_handle_t h, const ddp_model* model, const double* x, const double* u, double* fx, double* fu,
                         double* cx, double* cu) {
    if (!h) return DDP_ERR_INVALID;
    if (!model || !x || !u || !cx || !cu || !model->Q.ptr || !model->R.ptr) { h->err = "ddp_model_derivs_f64: missing argument"; return DDP_ERR_INVALID; }
    const int n = h->n, m = h->m, T = h->T;
    const long long B = h->B, wtot = B * T;
    if (cudaSetDevice(h->device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return DDP_ERR_CUDA; }
    launch_df_cost(h, h->stream, n, m, T, B, x, u, mk(model->Q), mk(model->R), model->goal, nullptr, cx, cu, (model->flags & DDP_MODEL_Q_DIAGONAL) != 0);
    if (model->kind == DDP_MODEL_PENDCART) {
        if (n != 4 || m != 1 || !fx || !fu) { h->err = "ddp_model_derivs_f64: pendcart needs n == 4, m == 1 and fx, fu outputs"; return DDP_ERR_INVALID; }
        df_pendcart_kernel<<<(unsigned)((wtot + 127) / 128), 128, 0, h->stream>>>(T, B, x, u, model->p[0], model->p[1], model->p[2], model->p[3],
                                                                                 nullptr, fx, fu);
        h->launches++;
    } else if (model->kind != DDP_MODEL_LINEAR) {
        h->err = "ddp_model_derivs_f64: unknown model kind (no CPU fallback for host callbacks)";
        return DDP_ERR_UNSUPPORTED;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { h->err = std::string("ddp_model_derivs_f64: ") + cudaGetErrorString(e); return DDP_ERR_CUDA; }
    return DDP_OK;
}

int ddp_ilqg_solve_f64(ddp_handle_t h, const ddp_model* model, const ddp_ilqg_opts* opts, const double* x0, const double* u0,
                       double* x, double* u, double* K, double* k, double* Vx, double* Vxx1, ddp_ilqg_state* state,
                       int32_t* n_outer) {
    if (!h) return DDP_ERR_INVALID;
    if (!model || !opts || (!x0 && !opts->x_init) || !u0 || !x || !u || !K || !k || !Vx || !state) { h->err = "ddp_ilqg_solve_f64: missing argument"; return DDP_ERR_INVALID; }
    if (model->kind != DDP_MODEL_LINEAR && model->kind != DDP_MODEL_PENDCART) {
        h->err = "ddp_ilqg_solve_f64: unknown model kind (arbitrary host callbacks cannot run on the device; no CPU fallback)";
        return DDP_ERR_UNSUPPORTED;
    }
    if (model->kind == DDP_MODEL_PENDCART && (h->n != 4 || h->m != 1)) { h->err = "pendcart model needs n == 4, m == 1"; return DDP_ERR_INVALID; }
    if (opts->n_alpha < 1 || opts->n_alpha > 16) { h->err = "ddp_ilqg_solve_f64: need 1 <= n_alpha <= 16"; return DDP_ERR_INVALID; }
    if (opts->reg_type != 1 && opts->reg_type != 2) { h->err = "ddp_ilqg_solve_f64: reg_type must be 1 or 2"; return DDP_ERR_INVALID; }
    if ((opts->x_init == nullptr) != (opts->cost_init == nullptr)) {
        h->err = "ddp_ilqg_solve_f64: Initial trajectory supplied, initial cost must also be supplied (x_init and cost_init go together)";
        return DDP_ERR_INVALID;
    }
    const int n = h->n, m = h->m, T = h->T;
    const long long B = h->B;
    cudaError_t err = cudaSuccess;
    DevBuf mem(h);
    SolveState s{};
    SolveOpts o{};
    double *cx = nullptr, *cu = nullptr, *xnew = nullptr, *unew = nullptr, *fxb = nullptr, *fub = nullptr, *zeros = nullptr;
    double* multi_costs = nullptr;
    // multi-alpha line search (headline shape only): after a rejected alpha[0] all remaining step sizes are evaluated in one launch
    const bool use_multi = ((model->kind == DDP_MODEL_LINEAR && h->n == 32 && h->m == 8 && model->A.stride_t == 0 && model->Bm.stride_t == 0) ||
                            (model->kind == DDP_MODEL_PENDCART && h->T % 2 == 0)) &&
                           !(h->flags & 1u) && !getenv("DDP_NO_MULTI_ALPHA");
    int hc[4] = {0, 0, 0, 0};
    int outer = 0;
    bool incomplete = false;
    const unsigned gB = (unsigned)((B + 255) / 256), gW = (unsigned)((B * 32 + 127) / 128);
    cudaStream_t st = h->stream;
    ModelD M;
    BackParams BP{};
    FwdParams FP{};

    o.n_alpha = opts->n_alpha;
    for (int i = 0; i < 16; i++) o.alpha[i] = opts->alpha[i];
    o.tol_fun = opts->tol_fun; o.tol_grad = opts->tol_grad; o.lam_factor = opts->lambda_factor; o.lam_max = opts->lambda_max;
    o.lam_min = opts->lambda_min; o.reduce_ratio_min = opts->reduce_ratio_min; o.max_iter = opts->max_iter;

    CUS(cudaSetDevice(h->device));
    CUS(mem.reserve((size_t)B * 8 * 40 + (size_t)B * T * (n + m) * 16 + (model->kind == DDP_MODEL_PENDCART ? (size_t)B * T * 160 : 0) + (size_t)B * 128 + (1 << 16)));
    CUS(mem.alloc(&s.lambda, B)); CUS(mem.alloc(&s.dlambda, B)); CUS(mem.alloc(&s.cost, B)); CUS(mem.alloc(&s.costnew, B));
    CUS(mem.alloc(&s.alpha, B)); CUS(mem.alloc(&s.gnorm, B)); CUS(mem.alloc(&s.dV, 2 * B)); CUS(mem.alloc(&s.last_dcost, B));
    CUS(mem.alloc(&s.aidx, B)); CUS(mem.alloc(&s.status, B)); CUS(mem.alloc(&s.iter, B)); CUS(mem.alloc(&s.acc, B));
    CUS(mem.alloc(&s.diverge, B)); CUS(mem.alloc(&s.bpr, B)); CUS(mem.alloc(&s.ratio, B));
    s.trace = (opts->trace && opts->trace_cap > 0) ? opts->trace : nullptr; s.trace_cap = s.trace ? opts->trace_cap : 0; s.B = B;
    CUS(mem.alloc(&s.active, B)); CUS(mem.alloc(&s.need_bp, B)); CUS(mem.alloc(&s.bp_ok, B)); CUS(mem.alloc(&s.need_fwd, B));
    CUS(mem.alloc(&s.accepted, B)); CUS(mem.alloc(&s.counters, 4));
    CUS(mem.alloc(&cx, (size_t)B * T * n)); CUS(mem.alloc(&cu, (size_t)B * T * m));
    CUS(mem.alloc(&xnew, (size_t)B * T * n)); CUS(mem.alloc(&unew, (size_t)B * T * m));
    CUS(mem.alloc(&zeros, (size_t)n * m));
    CUS(cudaMemsetAsync(zeros, 0, sizeof(double) * n * m, st));
    if (model->kind == DDP_MODEL_PENDCART) { CUS(mem.alloc(&fxb, (size_t)B * T * 16)); CUS(mem.alloc(&fub, (size_t)B * T * 4)); }
    if (use_multi) CUS(mem.alloc(&multi_costs, (size_t)B * 16));

    M.kind = model->kind; M.A = mk(model->A); M.Bm = mk(model->Bm); M.Q = mk(model->Q); M.R = mk(model->R); M.goal = model->goal;
    for (int i = 0; i < 8; i++) M.p[i] = model->p[i];
    M.terminal_cost = model->terminal_cost ? 1 : 0;
    M.flags = model->flags;

    init_state_kernel<<<gB, 256, 0, st>>>(B, s, opts->lambda, opts->dlambda, o.alpha[0]);
    h->launches++;

    // forward-pass parameter block (re-used for the initial rollout and the line search)
    FP.n = n; FP.m = m; FP.T = T; FP.B = B; FP.model = M;
    FP.x0 = TensorD{x0, n, 0};
    FP.lims = opts->lims; FP.active = s.need_fwd;
    FP.xnew = xnew; FP.unew = unew; FP.cost = s.costnew; FP.cost_t = nullptr; FP.cx = nullptr; FP.cu = nullptr;
    FP.alpha_scalar = 1.0;

    if (opts->x_init) {
        // ---- pre-rolled initial trajectory and its cost (iLQG.jl:193-197)
        CUS(cudaMemcpyAsync(x, opts->x_init, sizeof(double) * (size_t)B * T * n, cudaMemcpyDeviceToDevice, st));
        CUS(cudaMemcpyAsync(u, u0, sizeof(double) * (size_t)B * T * m, cudaMemcpyDeviceToDevice, st));
        init_prerolled_kernel<<<gB, 256, 0, st>>>(B, s, opts->cost_init);
        h->launches++;
        FP.x0 = TensorD{opts->x_init, (long long)T * n, 0};                // x0[:,1] of the pre-rolled trajectory (:268)
    }
    // ---- initial rollout over α with the open-loop controls αi*u0 (iLQG.jl:181-192)
    for (int ai = 0; ai < o.n_alpha && !opts->x_init; ai++) {
        FP.K = nullptr; FP.k = nullptr; FP.x = TensorD{nullptr, 0, 0};
        FP.u = TensorD{u0, (long long)T * m, m};
        FP.alpha = nullptr; FP.u_scale = o.alpha[ai];
        CUS(cudaMemsetAsync(s.counters, 0, 4 * sizeof(int), st));
        CUS((cudaError_t)run_fwd(h, FP));
        init_check_kernel<<<gW, 128, 0, st>>>(n, m, T, B, xnew, unew, s.costnew, x, u, s, o);
        h->launches++;
        CUS(cudaMemcpyAsync(hc, s.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
        CUS(cudaStreamSynchronize(st));
        if (hc[3] == 0) break;
    }

    // backward-pass parameter block
    BP.n = n; BP.m = m; BP.T = T; BP.B = B;
    BP.cx = TensorD{cx, (long long)T * n, n}; BP.cu = TensorD{cu, (long long)T * m, m};
    BP.cxx = M.Q; BP.cuu = M.R; BP.cxu = TensorD{zeros, 0, 0};
    if (model->kind == DDP_MODEL_LINEAR) { BP.fx = M.A; BP.fu = M.Bm; }
    else { BP.fx = TensorD{fxb, (long long)T * 16, 16}; BP.fu = TensorD{fub, (long long)T * 4, 4}; }
    BP.u = TensorD{u, (long long)T * m, m};
    BP.lambda = s.lambda; BP.reg_type = opts->reg_type; BP.lims = opts->lims; BP.active = s.need_bp;
    BP.Kp = TensorD{nullptr, 0, 0}; BP.kp = BP.Kp; BP.Sip = BP.Kp; BP.eta = nullptr; BP.Quui = nullptr;
    BP.diverge = s.diverge; BP.K = K; BP.k = k; BP.Vx = Vx; BP.Vxx = nullptr; BP.Vxx1 = Vxx1; BP.Quu = nullptr; BP.dV = s.dV;
    BP.qp = QPOpts{100, 1e-8, 1e-8, 0.6, 1e-22, 0.1};

    {
        // Every outer iteration either accepts a step (at most max_iter + 1 of those per trajectory) or raises λ, and λ
        // passes λmax after at most ~13 consecutive increases from λmin (λ grows by 1.6^k at the k-th): 16 (max_iter + 1)
        // + 256 outer iterations cannot be exceeded by the reference's own rules; the cap only guards the loop.
        const int outer_cap = (opts->max_iter + 1) * 16 + 256;
        for (outer = 0; outer < outer_cap; outer++) {
            // STEP 1: derivatives along the trajectories whose x,u changed (need_bp marks exactly those here)
            {
                long long wtot = B * T;
                launch_df_cost(h, st, n, m, T, B, x, u, M.Q, M.R, M.goal, s.need_bp, cx, cu, (M.flags & DDP_MODEL_Q_DIAGONAL) != 0);
                if (model->kind == DDP_MODEL_PENDCART) {
                    df_pendcart_kernel<<<(unsigned)((wtot + 127) / 128), 128, 0, st>>>(T, B, x, u, M.p[0], M.p[1], M.p[2], M.p[3],
                                                                                       s.need_bp, fxb, fub);
                    h->launches++;
                }
            }
            // STEP 2: backward pass, retried with larger λ where it diverged
            for (;;) {
                CUS(cudaMemsetAsync(s.counters, 0, 4 * sizeof(int), st));
                CUS((cudaError_t)run_back(h, BP));
                bp_retry_kernel<<<gB, 256, 0, st>>>(B, s, o);
                h->launches++;
                CUS(cudaMemcpyAsync(hc, s.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
                CUS(cudaStreamSynchronize(st));
                if (hc[0] == 0) break;
            }
            gnorm_kernel<<<gW, 128, 0, st>>>(m, T, B, k, u, s, o);
            h->launches++;
            // STEP 3: serial backtracking line search, one launch per α still needed by anyone
            FP.K = K; FP.k = k; FP.x = TensorD{x, (long long)T * n, n}; FP.u = TensorD{u, (long long)T * m, m};
            FP.alpha = s.alpha; FP.u_scale = 1.0;
            for (int ai = 0; ai < o.n_alpha; ai++) {
                CUS(cudaMemsetAsync(s.counters, 0, 4 * sizeof(int), st));
                CUS((cudaError_t)run_fwd(h, FP));
                linesearch_kernel<<<gB, 256, 0, st>>>(B, s, o);
                h->launches++;
                CUS(cudaMemcpyAsync(hc, s.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
                CUS(cudaStreamSynchronize(st));
                if (hc[1] == 0) break;
                if (ai == 0 && use_multi && o.n_alpha > 2) {
                    // someone rejected alpha[0]: evaluate ALL remaining step sizes in one pass over K (multi-alpha
                    // rollout, costs only), pick the first acceptable one per trajectory, then roll that one out
                    bool handled = false;
                    FwdParams FM = FP;
                    FM.alpha = nullptr;
                    CUS((cudaError_t)launch_forward_multi(h, FM, o.n_alpha - 1, o.alpha + 1, multi_costs, &handled));
                    if (handled) {
                        CUS(cudaMemsetAsync(s.counters, 0, 4 * sizeof(int), st));
                        linesearch_multi_kernel<<<gB, 256, 0, st>>>(B, s, o, multi_costs);
                        h->launches++;
                        CUS(cudaMemcpyAsync(hc, s.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
                        CUS(cudaStreamSynchronize(st));
                        if (hc[1] > 0) CUS((cudaError_t)run_fwd(h, FP));          // per-trajectory accepted alpha, masked by need_fwd
                        clear_flags_kernel<<<gB, 256, 0, st>>>(B, s.need_fwd);
                        h->launches++;
                        break;
                    }
                }
            }
            // STEP 4: accept / reject
            CUS(cudaMemsetAsync(s.counters, 0, 4 * sizeof(int), st));
            copy_accepted_kernel<<<(unsigned)std::min<long long>(B, (long long)h->sm_count * 8), 256, 0, st>>>(n, m, T, B, s.accepted, xnew,
                                                                                                          unew, x, u, k);
            accept_kernel<<<gB, 256, 0, st>>>(B, s, o);
            h->launches += 2;
            CUS(cudaMemcpyAsync(hc, s.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
            CUS(cudaStreamSynchronize(st));
            if (hc[2] == 0) { outer++; break; }
        }
        if (hc[2] != 0) { mark_incomplete_kernel<<<gB, 256, 0, st>>>(B, s); h->launches++; incomplete = true; }
    }
    export_state_kernel<<<gB, 256, 0, st>>>(B, s, state);
    h->launches++;
    CUS(cudaStreamSynchronize(st));
    if (n_outer) *n_outer = outer;
    if (incomplete) { h->err = "ddp_ilqg_solve_f64: outer-loop safety cap reached; trajectories with status 6 did not finish"; return DDP_ERR_INCOMPLETE; }
    return DDP_OK;
fail:
    h->err = std::string("ddp_ilqg_solve_f64: ") + cudaGetErrorString(err);
    return DDP_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

// ---------------------------------------------------------------------------------------------
// One chunk of one iteration: derivative step + backward sweep + forward rollout of `nb` trajectories, back to back on
// stream `st`.  Shared by the device-resident chunked iteration (ddp_ilqg_iter_f64) and the host-buffer pipeline.
struct ChunkIO {
    const double *x, *u, *lambda, *alpha;           // inputs of the chunk
    const unsigned char* active;
    double *cx, *cu;                                 // derivative scratch (nb trajectories) -- or uploaded by the host path
    double *fx, *fu;                                 // pendcart Jacobian scratch (nb) or nullptr
    double *K, *k, *Vx;                              // policy of the chunk
    double *xnew, *unew, *cost, *dV;
    int* diverge;
};

int run_chunk(ddp_handle_s* h, cudaStream_t st, const ModelD& M, const ChunkIO& c, long long nb, int reg_type, double alpha_scalar,
              const double* lims, const double* cxu_zero, bool form_derivs) {
    const int n = h->n, m = h->m, T = h->T;
    const long long Tn = (long long)T * n, Tm = (long long)T * m;
    cudaStream_t saved = h->stream;
    h->stream = st;
    int rc = 0;
    if (form_derivs) {                               // STEP 1 (iLQG.jl:225-229)
        launch_df_cost(h, st, n, m, T, nb, c.x, c.u, M.Q, M.R, M.goal, c.active, c.cx, c.cu, (M.flags & DDP_MODEL_Q_DIAGONAL) != 0);
        if (M.kind == DDP_MODEL_PENDCART) {
            df_pendcart_kernel<<<(unsigned)((nb * T + 127) / 128), 128, 0, st>>>(T, nb, c.x, c.u, M.p[0], M.p[1], M.p[2], M.p[3], c.active, c.fx, c.fu);
            h->launches++;
        }
    }
    BackParams BP{};
    BP.n = n; BP.m = m; BP.T = T; BP.B = nb;
    BP.cx = TensorD{c.cx, Tn, n}; BP.cu = TensorD{c.cu, Tm, m};
    BP.cxx = M.Q; BP.cuu = M.R; BP.cxu = TensorD{cxu_zero, 0, 0};
    if (M.kind == DDP_MODEL_PENDCART) { BP.fx = TensorD{c.fx, (long long)T * 16, 16}; BP.fu = TensorD{c.fu, (long long)T * 4, 4}; }
    else { BP.fx = M.A; BP.fu = M.Bm; }
    BP.u = TensorD{c.u, Tm, m};
    BP.lambda = c.lambda; BP.reg_type = reg_type; BP.lims = lims; BP.active = c.active;
    BP.diverge = c.diverge; BP.K = c.K; BP.k = c.k; BP.Vx = c.Vx; BP.dV = c.dV;
    BP.qp = QPOpts{100, 1e-8, 1e-8, 0.6, 1e-22, 0.1};
    rc = run_back(h, BP);
    if (rc == 0) {
        FwdParams FP{};
        FP.n = n; FP.m = m; FP.T = T; FP.B = nb; FP.model = M;
        FP.K = c.K; FP.k = c.k; FP.x0 = TensorD{c.x, Tn, 0}; FP.x = TensorD{c.x, Tn, n}; FP.u = TensorD{c.u, Tm, m};
        FP.alpha = c.alpha; FP.alpha_scalar = alpha_scalar; FP.u_scale = 1.0;
        FP.lims = lims; FP.active = c.active; FP.xnew = c.xnew; FP.unew = c.unew; FP.cost = c.cost;
        rc = run_fwd(h, FP);
    }
    h->stream = saved;
    return rc;
}

// model descriptor of a chunk: per-trajectory tensors advance by b0 trajectories, shared ones stay
ModelD model_at(const ModelD& M, long long b0) {
    ModelD r = M;
    r.A.p = M.A.p ? M.A.p + b0 * M.A.sb : nullptr;
    r.Bm.p = M.Bm.p ? M.Bm.p + b0 * M.Bm.sb : nullptr;
    r.Q.p = M.Q.p + b0 * M.Q.sb;
    r.R.p = M.R.p + b0 * M.R.sb;
    return r;
}

bool fill_model_d(const ddp_model* model, ModelD& M) {
    M.kind = model->kind; M.A = mk(model->A); M.Bm = mk(model->Bm); M.Q = mk(model->Q); M.R = mk(model->R); M.goal = model->goal;
    for (int i = 0; i < 8; i++) M.p[i] = model->p[i];
    M.terminal_cost = model->terminal_cost ? 1 : 0;
    M.flags = model->flags;
    return true;
}

// chunk-sized scratch of the device-resident chunked iteration, kept on the handle between calls
struct ChunkScratch {
    std::vector<void*> ptrs;
    long long cap = 0;
    bool policy = false, pend = false;
    double *cx = nullptr, *cu = nullptr, *fx = nullptr, *fu = nullptr, *K = nullptr, *k = nullptr, *Vx = nullptr, *zeros = nullptr;
    ~ChunkScratch() { for (void* p : ptrs) cudaFree(p); }
    template <typename T_>
    cudaError_t get(T_** p, size_t count) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, std::max<size_t>(count * sizeof(T_), 256));
        if (e == cudaSuccess) { ptrs.push_back(q); *p = (T_*)q; }
        return e;
    }
};
void free_chunk_scratch(void* p) { delete static_cast<ChunkScratch*>(p); }

cudaError_t make_chunk_scratch(ddp_handle_s* h, long long cap, bool policy, bool pend, ChunkScratch** out) {
    ChunkScratch* c = new ChunkScratch();
    cudaError_t err = cudaSuccess;
    const size_t Tn = (size_t)h->T * h->n, Tm = (size_t)h->T * h->m;
    CUS(c->get(&c->cx, cap * Tn)); CUS(c->get(&c->cu, cap * Tm));
    if (policy) { CUS(c->get(&c->K, cap * Tm * h->n)); CUS(c->get(&c->k, cap * Tm)); CUS(c->get(&c->Vx, cap * Tn)); }
    if (pend) { CUS(c->get(&c->fx, cap * h->T * 16)); CUS(c->get(&c->fu, cap * h->T * 4)); }
    CUS(c->get(&c->zeros, (size_t)h->n * h->m));
    CUS(cudaMemset(c->zeros, 0, sizeof(double) * h->n * h->m));
    c->cap = cap; c->policy = policy; c->pend = pend;
    *out = c;
    return cudaSuccess;
fail:
    delete c;
    return err;
}

// ---------------------------------------------------------------------------------------------
// host-buffer pipeline state: full-batch device mirrors of every host array (a chunk is a slice of them, so nothing is
// recycled and the policy of the whole batch can stay), three streams, a ring of events
constexpr int NEV = 64;
struct IterCache {
    std::vector<void*> ptrs;
    cudaStream_t s_in = nullptr, s_cp = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[NEV] = {}, ev_cp[NEV] = {}, ev_start = nullptr, ev_end = nullptr;
    double *fx = nullptr, *fu = nullptr, *x = nullptr, *u = nullptr, *lam = nullptr, *xnew = nullptr, *unew = nullptr, *cost = nullptr, *dV = nullptr, *cprev = nullptr;
    double *cx = nullptr, *cu = nullptr;          // full batch when the host supplies them, else one chunk
    double *K = nullptr, *k = nullptr, *Vx = nullptr;   // full batch (keep_policy) or one chunk
    int* div = nullptr;
    double *dQ = nullptr, *dR = nullptr, *dcxu = nullptr;
    long long chunk = 0;
    bool keep_policy = false, host_derivs = false, inputs_valid = false;
    template <typename T_>
    cudaError_t get(T_** p, size_t count) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, std::max<size_t>(count * sizeof(T_), 256));
        if (e == cudaSuccess) { ptrs.push_back(q); *p = (T_*)q; }
        return e;
    }
    ~IterCache() {
        for (int i = 0; i < NEV; i++) {
            if (ev_in[i]) cudaEventDestroy(ev_in[i]);
            if (ev_cp[i]) cudaEventDestroy(ev_cp[i]);
        }
        if (ev_start) cudaEventDestroy(ev_start);
        if (ev_end) cudaEventDestroy(ev_end);
        if (s_in) cudaStreamDestroy(s_in);
        if (s_cp) cudaStreamDestroy(s_cp);
        if (s_out) cudaStreamDestroy(s_out);
        for (void* p : ptrs) cudaFree(p);
    }
};
void free_iter_cache(void* p) { delete static_cast<IterCache*>(p); }

cudaError_t make_iter_cache(ddp_handle_s* h, long long chunk, bool keep_policy, bool host_derivs, IterCache** out) {
    IterCache* c = new IterCache();
    cudaError_t err = cudaSuccess;
    const int n = h->n, m = h->m, T = h->T;
    const size_t B = (size_t)h->B;
    const size_t nn = (size_t)n * n, nm = (size_t)n * m, Tn = (size_t)T * n, Tm = (size_t)T * m;
    const size_t pol = keep_policy ? B : (size_t)chunk, der = host_derivs ? B : (size_t)chunk;
    CUS(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    CUS(cudaStreamCreateWithFlags(&c->s_cp, cudaStreamNonBlocking));
    CUS(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < NEV; i++) {
        CUS(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
        CUS(cudaEventCreateWithFlags(&c->ev_cp[i], cudaEventDisableTiming));
    }
    CUS(cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
    CUS(cudaEventCreateWithFlags(&c->ev_end, cudaEventDisableTiming));
    CUS(c->get(&c->fx, B * nn)); CUS(c->get(&c->fu, B * nm)); CUS(c->get(&c->x, B * Tn)); CUS(c->get(&c->u, B * Tm)); CUS(c->get(&c->lam, B));
    CUS(c->get(&c->xnew, B * Tn)); CUS(c->get(&c->unew, B * Tm)); CUS(c->get(&c->cost, B)); CUS(c->get(&c->dV, 2 * B)); CUS(c->get(&c->div, B));
    CUS(c->get(&c->cx, der * Tn)); CUS(c->get(&c->cu, der * Tm));
    CUS(c->get(&c->K, pol * Tm * n)); CUS(c->get(&c->k, pol * Tm)); CUS(c->get(&c->Vx, pol * Tn));
    CUS(c->get(&c->dQ, nn)); CUS(c->get(&c->dR, (size_t)m * m)); CUS(c->get(&c->dcxu, nm)); CUS(c->get(&c->cprev, B));
    c->chunk = chunk; c->keep_policy = keep_policy; c->host_derivs = host_derivs;
    *out = c;
    return cudaSuccess;
fail:
    delete c;
    return err;
}

// x,u <- xnew,unew where the step was accepted: ratio = Δcost / expected > 0, expected = -α(dV1 + α dV2) (iLQG.jl:269-280, 302)
__global__ void __launch_bounds__(256) commit_accepted_kernel(int n, int m, int T, long long B, double alpha, const double* __restrict__ cost_old,
                                                              const double* __restrict__ cost_new, const double* __restrict__ dV,
                                                              const int* __restrict__ diverge, const double* __restrict__ xnew,
                                                              const double* __restrict__ unew, double* __restrict__ x, double* __restrict__ u) {
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        if (diverge[b] > 0) continue;
        const double dcost = cost_old[b] - cost_new[b];
        const double expected = -alpha * (dV[2 * b] + alpha * dV[2 * b + 1]);
        const double ratio = (expected > 0) ? dcost / expected : ((dcost > 0) ? 1.0 : ((dcost < 0) ? -1.0 : dcost));
        if (!(ratio > 0.0)) continue;
        const long long ox = b * T * n, ou = b * T * m;
        for (int e = threadIdx.x; e < T * n; e += blockDim.x) x[ox + e] = xnew[ox + e];
        for (int e = threadIdx.x; e < T * m; e += blockDim.x) u[ou + e] = unew[ou + e];
    }
}
}  // namespace

extern "C" {

int ddp_ilqg_iter_host_f64(ddp_handle_t h, ddp_iter_host_args* a) {
    if (!h) return DDP_ERR_INVALID;
    const bool resident = a && a->inputs_resident != 0;
    if (!a || (!resident && (!a->fx || !a->fu || !a->x || !a->u || !a->lambda)) || ((a->cx == nullptr) != (a->cu == nullptr)) || !a->Q || !a->R ||
        !a->cxu || !a->xnew || !a->unew || !a->cost || !a->dV || !a->diverge) {
        h->err = "ddp_ilqg_iter_host_f64: missing argument";
        return DDP_ERR_INVALID;
    }
    if (a->reg_type != 1 && a->reg_type != 2) { h->err = "ddp_ilqg_iter_host_f64: reg_type must be 1 or 2"; return DDP_ERR_INVALID; }
    if (a->commit_accepted && !a->cost_prev) { h->err = "ddp_ilqg_iter_host_f64: commit_accepted needs cost_prev (the cost of the current x,u)"; return DDP_ERR_INVALID; }
    const int n = h->n, m = h->m, T = h->T;
    const long long B = h->B;
    // Chunk schedule.  Default: chunks are whole "rounds" of the resident warp set of the sweep kernels
    // (sm_count x 8 warps, one trajectory each), so no launch ends with a partly filled round: two rounds per chunk (measured on a
    // B200 at 65 536 trajectories: 1 round 140.7 ms, 2 rounds 135.5 ms, 4 rounds 138.3 ms, 8 rounds 148.7 ms per step); the first and
    // the last chunk are one round to shorten the fill (first H2D) and the drain (last D2H) of the pipeline.
    std::vector<long long> sizes;
    long long chunk;
    if (a->chunk > 0) {
        chunk = std::min<long long>(a->chunk, B);
        for (long long b0 = 0; b0 < B; b0 += chunk) sizes.push_back(std::min<long long>(chunk, B - b0));
    } else {
        const long long unit = (long long)h->sm_count * 8;
        chunk = std::min<long long>(2 * unit, B);
        long long left = B;
        auto take = [&](long long rounds) { long long nb = std::min<long long>(rounds * unit, left); if (nb > 0) { sizes.push_back(nb); left -= nb; } };
        const long long R = (B + unit - 1) / unit;
        if (R >= 6) take(1);
        while (left > (R >= 6 ? 1 : 0) * unit) take(2);
        take(1);
    }
    cudaError_t err = cudaSuccess;
    const size_t nn = (size_t)n * n, nm = (size_t)n * m, Tn = (size_t)T * n, Tm = (size_t)T * m;
    long long h2d = 0, d2h = 0;
    ModelD M{};
    IterCache* C = nullptr;
    const bool host_derivs = (a->cx != nullptr);   // NULL: cx = Qx, cu = Ru are formed on the device (the reference's df step)
    const bool keep = a->keep_policy != 0;

    CUS(cudaSetDevice(h->device));
    if (h->cache && h->cache_free != free_iter_cache) { h->cache_free(h->cache); h->cache = nullptr; }     // another entry point's scratch
    if (h->cache) {
        IterCache* c0 = static_cast<IterCache*>(h->cache);
        if (c0->chunk != chunk || c0->keep_policy != keep || c0->host_derivs != host_derivs) {
            if (resident) { h->err = "ddp_ilqg_iter_host_f64: inputs_resident needs the same chunk / keep_policy / cx,cu configuration as the call that uploaded them"; return DDP_ERR_INVALID; }
            h->cache_free(h->cache); h->cache = nullptr;
        }
    }
    if (!h->cache) {
        if (resident) { h->err = "ddp_ilqg_iter_host_f64: inputs_resident without a previous call on this handle"; return DDP_ERR_INVALID; }
        err = make_iter_cache(h, chunk, keep, host_derivs, &C);
        if (err == cudaErrorMemoryAllocation) { cudaGetLastError(); h->err = "ddp_ilqg_iter_host_f64: out of device memory (try keep_policy = 0)"; return DDP_ERR_NOMEM; }
        CUS(err);
        h->cache = C;
        h->cache_free = free_iter_cache;
    }
    C = static_cast<IterCache*>(h->cache);
    if (resident && !C->inputs_valid) { h->err = "ddp_ilqg_iter_host_f64: inputs_resident without a completed previous call"; return DDP_ERR_INVALID; }
    {
    cudaStream_t saved = h->stream;
    cudaStream_t s_in = C->s_in, s_cp = C->s_cp, s_out = C->s_out;
    // everything below is ordered after work already queued on the handle's stream
    CUS(cudaEventRecord(C->ev_start, saved));
    CUS(cudaStreamWaitEvent(s_in, C->ev_start, 0));
    CUS(cudaStreamWaitEvent(s_cp, C->ev_start, 0));
    CUS(cudaStreamWaitEvent(s_out, C->ev_start, 0));
    CUS(cudaMemcpyAsync(C->dQ, a->Q, nn * 8, cudaMemcpyHostToDevice, s_in));
    CUS(cudaMemcpyAsync(C->dR, a->R, (size_t)m * m * 8, cudaMemcpyHostToDevice, s_in));
    CUS(cudaMemcpyAsync(C->dcxu, a->cxu, nm * 8, cudaMemcpyHostToDevice, s_in));
    h2d += (long long)(nn + (size_t)m * m + nm) * 8;
    M.kind = DDP_MODEL_LINEAR; M.goal = nullptr; M.terminal_cost = 0; M.flags = a->q_diagonal ? DDP_MODEL_Q_DIAGONAL : 0;
    M.Q = TensorD{C->dQ, 0, 0}; M.R = TensorD{C->dR, 0, 0};
    {
        long long b0 = 0;
        for (int c = 0; c < (int)sizes.size(); b0 += sizes[c], c++) {
            const long long nb = sizes[c];
            const int ei = c % NEV;
            const long long pb = keep ? b0 : 0, db = host_derivs ? b0 : 0;     // slice of the policy / derivative arrays
            if (!resident) {
                CUS(cudaMemcpyAsync(C->fx + b0 * nn, a->fx + b0 * nn, nb * nn * 8, cudaMemcpyHostToDevice, s_in));
                CUS(cudaMemcpyAsync(C->fu + b0 * nm, a->fu + b0 * nm, nb * nm * 8, cudaMemcpyHostToDevice, s_in));
                CUS(cudaMemcpyAsync(C->x + b0 * Tn, a->x + b0 * Tn, nb * Tn * 8, cudaMemcpyHostToDevice, s_in));
                CUS(cudaMemcpyAsync(C->u + b0 * Tm, a->u + b0 * Tm, nb * Tm * 8, cudaMemcpyHostToDevice, s_in));
                CUS(cudaMemcpyAsync(C->lam + b0, a->lambda + b0, nb * 8, cudaMemcpyHostToDevice, s_in));
                h2d += (long long)nb * (long long)(nn + nm + Tn + Tm + 1) * 8;
            }
            if (host_derivs) {
                CUS(cudaMemcpyAsync(C->cx + b0 * Tn, a->cx + b0 * Tn, nb * Tn * 8, cudaMemcpyHostToDevice, s_in));
                CUS(cudaMemcpyAsync(C->cu + b0 * Tm, a->cu + b0 * Tm, nb * Tm * 8, cudaMemcpyHostToDevice, s_in));
                h2d += (long long)nb * (long long)(Tn + Tm) * 8;
            }
            CUS(cudaEventRecord(C->ev_in[ei], s_in));
            CUS(cudaStreamWaitEvent(s_cp, C->ev_in[ei], 0));
            ChunkIO io{};
            io.x = C->x + b0 * Tn; io.u = C->u + b0 * Tm; io.lambda = C->lam + b0; io.alpha = nullptr; io.active = nullptr;
            io.cx = C->cx + db * Tn; io.cu = C->cu + db * Tm; io.fx = nullptr; io.fu = nullptr;
            io.K = C->K + pb * Tm * n; io.k = C->k + pb * Tm; io.Vx = C->Vx + pb * Tn;
            io.xnew = C->xnew + b0 * Tn; io.unew = C->unew + b0 * Tm; io.cost = C->cost + b0; io.dV = C->dV + 2 * b0; io.diverge = C->div + b0;
            ModelD Mc = M;
            Mc.A = TensorD{C->fx + b0 * nn, (long long)nn, 0}; Mc.Bm = TensorD{C->fu + b0 * nm, (long long)nm, 0};
            const int rc = run_chunk(h, s_cp, Mc, io, nb, a->reg_type, a->alpha, nullptr, C->dcxu, !host_derivs);
            CUS((cudaError_t)rc);
            CUS(cudaEventRecord(C->ev_cp[ei], s_cp));
            // results back to the host
            CUS(cudaStreamWaitEvent(s_out, C->ev_cp[ei], 0));
            CUS(cudaMemcpyAsync(a->xnew + b0 * Tn, io.xnew, nb * Tn * 8, cudaMemcpyDeviceToHost, s_out));
            CUS(cudaMemcpyAsync(a->unew + b0 * Tm, io.unew, nb * Tm * 8, cudaMemcpyDeviceToHost, s_out));
            CUS(cudaMemcpyAsync(a->cost + b0, io.cost, nb * 8, cudaMemcpyDeviceToHost, s_out));
            CUS(cudaMemcpyAsync(a->dV + 2 * b0, io.dV, nb * 16, cudaMemcpyDeviceToHost, s_out));
            CUS(cudaMemcpyAsync(a->diverge + b0, io.diverge, nb * 4, cudaMemcpyDeviceToHost, s_out));
            d2h += (long long)nb * (long long)((Tn + Tm + 3) * 8 + 4);
        }
    }
    if (a->commit_accepted) {          // x,u <- xnew,unew on the device where the step is accepted (iLQG.jl:275-303)
        CUS(cudaMemcpyAsync(C->cprev, a->cost_prev, (size_t)B * 8, cudaMemcpyHostToDevice, s_cp));
        h2d += B * 8;
        commit_accepted_kernel<<<(unsigned)std::min<long long>(B, (long long)h->sm_count * 8), 256, 0, s_cp>>>(n, m, T, B, a->alpha, C->cprev, C->cost, C->dV,
                                                                                                              C->div, C->xnew, C->unew, C->x, C->u);
        h->launches++;
    }
    CUS(cudaEventRecord(C->ev_end, s_out));
    CUS(cudaStreamWaitEvent(saved, C->ev_end, 0));       // later work on the handle's stream sees the results
    CUS(cudaStreamSynchronize(s_out));
    CUS(cudaStreamSynchronize(s_cp));
    CUS(cudaStreamSynchronize(s_in));
    C->inputs_valid = true;
    }
    a->h2d_bytes = h2d;
    a->d2h_bytes = d2h;
    return DDP_OK;
fail:
    h->err = std::string("ddp_ilqg_iter_host_f64: ") + cudaGetErrorString(err);
    cudaDeviceSynchronize();
    return DDP_ERR_CUDA;
}

int ddp_iter_host_policy(ddp_handle_t h, double** K, double** k, double** Vx, double** xnew, double** unew) {
    if (!h) return DDP_ERR_INVALID;
    IterCache* C = (h->cache && h->cache_free == free_iter_cache) ? static_cast<IterCache*>(h->cache) : nullptr;
    if (!C || !C->inputs_valid) { h->err = "ddp_iter_host_policy: no completed ddp_ilqg_iter_host_f64 call on this handle"; return DDP_ERR_INVALID; }
    if (K) *K = C->keep_policy ? C->K : nullptr;
    if (k) *k = C->keep_policy ? C->k : nullptr;
    if (Vx) *Vx = C->keep_policy ? C->Vx : nullptr;
    if (xnew) *xnew = C->xnew;
    if (unew) *unew = C->unew;
    return DDP_OK;
}

int ddp_ilqg_iter_f64(ddp_handle_t h, const ddp_model* model, ddp_iter_args* a) {
    if (!h) return DDP_ERR_INVALID;
    if (!model || !a || !a->x || !a->u || !a->lambda || !a->xnew || !a->unew || !a->cost || !a->dV || !a->diverge) {
        h->err = "ddp_ilqg_iter_f64: missing argument";
        return DDP_ERR_INVALID;
    }
    if (a->reg_type != 1 && a->reg_type != 2) { h->err = "ddp_ilqg_iter_f64: reg_type must be 1 or 2"; return DDP_ERR_INVALID; }
    const bool keep = (a->K != nullptr);
    if ((a->k != nullptr) != keep || (a->Vx != nullptr) != keep) { h->err = "ddp_ilqg_iter_f64: K, k, Vx must be all given or all NULL"; return DDP_ERR_INVALID; }
    if (model->kind != DDP_MODEL_LINEAR && model->kind != DDP_MODEL_PENDCART) { h->err = "ddp_ilqg_iter_f64: unknown model kind (no CPU fallback for host callbacks)"; return DDP_ERR_UNSUPPORTED; }
    if (model->kind == DDP_MODEL_PENDCART && (h->n != 4 || h->m != 1)) { h->err = "pendcart model needs n == 4, m == 1"; return DDP_ERR_INVALID; }
    if (model->kind == DDP_MODEL_LINEAR && (!model->A.ptr || !model->Bm.ptr)) { h->err = "linear model needs A and B"; return DDP_ERR_INVALID; }
    if (!model->Q.ptr || !model->R.ptr) { h->err = "model needs Q and R"; return DDP_ERR_INVALID; }
    const int n = h->n, m = h->m, T = h->T;
    const long long B = h->B;
    const size_t Tn = (size_t)T * n, Tm = (size_t)T * m;
    const bool pend = model->kind == DDP_MODEL_PENDCART;
    // default chunk: 55 rounds of the resident warp set (sm_count x 8 trajectories) ~ 65 120 trajectories on a B200
    long long chunk = a->chunk > 0 ? a->chunk : (long long)h->sm_count * 8 * 55;
    if (chunk > B) chunk = B;
    cudaError_t err = cudaSuccess;
    ModelD M{};
    fill_model_d(model, M);
    ChunkScratch* S = nullptr;
    CUS(cudaSetDevice(h->device));
    if (h->cache && h->cache_free != free_chunk_scratch) { h->cache_free(h->cache); h->cache = nullptr; }
    if (h->cache) {
        ChunkScratch* s0 = static_cast<ChunkScratch*>(h->cache);
        if (s0->cap < chunk || s0->policy != !keep || s0->pend != pend) { h->cache_free(h->cache); h->cache = nullptr; }
    }
    if (!h->cache) {
        err = make_chunk_scratch(h, chunk, !keep, pend, &S);
        if (err == cudaErrorMemoryAllocation) { cudaGetLastError(); h->err = "ddp_ilqg_iter_f64: out of device memory for the chunk scratch (use a smaller chunk)"; return DDP_ERR_NOMEM; }
        CUS(err);
        h->cache = S;
        h->cache_free = free_chunk_scratch;
    }
    S = static_cast<ChunkScratch*>(h->cache);
    {
        long long nchunks = 0;
        for (long long b0 = 0; b0 < B; b0 += chunk, nchunks++) {
            const long long nb = std::min<long long>(chunk, B - b0);
            ChunkIO io{};
            io.x = a->x + b0 * Tn; io.u = a->u + b0 * Tm; io.lambda = a->lambda + b0; io.alpha = a->alpha ? a->alpha + b0 : nullptr;
            io.active = a->active ? a->active + b0 : nullptr;
            io.cx = S->cx; io.cu = S->cu; io.fx = S->fx; io.fu = S->fu;
            io.K = keep ? a->K + b0 * Tm * n : S->K; io.k = keep ? a->k + b0 * Tm : S->k; io.Vx = keep ? a->Vx + b0 * Tn : S->Vx;
            io.xnew = a->xnew + b0 * Tn; io.unew = a->unew + b0 * Tm; io.cost = a->cost + b0; io.dV = a->dV + 2 * b0; io.diverge = a->diverge + b0;
            const int rc = run_chunk(h, h->stream, model_at(M, b0), io, nb, a->reg_type, a->alpha_scalar, a->lims, S->zeros, true);
            CUS((cudaError_t)rc);
        }
        a->n_chunks = nchunks;
    }
    return DDP_OK;
fail:
    h->err = std::string("ddp_ilqg_iter_f64: ") + cudaGetErrorString(err);
    return DDP_ERR_CUDA;
}

}  // extern "C"

// =============================================================================================
// ddp_ilqgkl_solve_f64: the iLQGkl outer loop (src/iLQGkl.jl:93-183) + calc_eta (src/klutils.jl:110-130)
// for a whole batch, device resident.  Per trajectory: eta bracket (3), del0, retry count, status.
namespace {

struct KlSolveState {
    double *eta3, *eta, *del0, *div, *dcost, *expected, *dV, *klmean;
    int *iter, *status, *retries, *diverge;
    unsigned char *active, *need_bp;
    int* counters;          // [0] back passes to retry, [1] trajectories still iterating
};

__global__ void kl_init_kernel(long long B, KlSolveState s, double e0, double e1, double e2, double del0) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    s.eta3[3 * b] = e0; s.eta3[3 * b + 1] = e1; s.eta3[3 * b + 2] = e2; s.eta[b] = e1; s.del0[b] = del0;
    s.div[b] = 0.0; s.dcost[b] = 0.0; s.expected[b] = 0.0; s.dV[2 * b] = s.dV[2 * b + 1] = 0.0; s.klmean[b] = 0.0;
    s.iter[b] = 0; s.status[b] = -1; s.retries[b] = 0; s.diverge[b] = 0; s.active[b] = 1; s.need_bp[b] = 1;
}

// after a KL-augmented backward launch: iLQGkl.jl:97-124 -- on failure eta += del0, del0 *= 2 and try again
__global__ void kl_retry_kernel(long long B, KlSolveState s, int max_retries) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || !s.need_bp[b]) return;
    if (s.diverge[b] > 0) {
        const double d0 = s.del0[b];
        const double e = s.eta3[3 * b + 1] + d0;          // :104
        s.eta3[3 * b + 1] = e;
        s.eta[b] = e;
        s.del0[b] = 2.0 * d0;                             // :105
        const int r = s.retries[b] + 1;
        s.retries[b] = r;
        if (r > max_retries) { s.need_bp[b] = 0; s.active[b] = 0; s.status[b] = 5; }
        else atomicAdd(&s.counters[0], 1);
    } else {
        s.need_bp[b] = 0;
    }
}

// after forward pass + KL evaluation: expected / actual improvement (iLQGkl.jl:137-139), calc_eta
// (klutils.jl:110-130) and the two exits (iLQGkl.jl:173-181)
__global__ void kl_eta_kernel(long long B, KlSolveState s, const double* __restrict__ cost, const double* __restrict__ costnew,
                              double kl_step, int it, int max_iter) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || !s.active[b]) return;
    s.iter[b] = it;
    s.dcost[b] = cost[b] - costnew[b];
    s.expected[b] = -(s.dV[2 * b] + s.dV[2 * b + 1]);     // :138
    double e0 = s.eta3[3 * b], e1 = s.eta3[3 * b + 1], e2 = s.eta3[3 * b + 2];
    bool satisfied;
    double divergence;
    if (!(kl_step > 0.0)) {                               // klutils.jl:111
        satisfied = true;
        divergence = 0.0;
    } else {
        divergence = s.klmean[b];
        const double violation = divergence - kl_step;
        satisfied = fabs(violation) < 0.1 * kl_step;      // :115
        if (!satisfied) {
            if (violation < 0.0) {                        // KL too small: eta is an upper bound (:119-122)
                e2 = e1;
                e1 = fmax(sqrt(e0 * e2), 0.1 * e2);
            } else {                                      // KL too big: eta is a lower bound (:123-126)
                e0 = e1;
                e1 = fmin(sqrt(e0 * e2), 10.0 * e0);
            }
        }
    }
    s.div[b] = divergence;
    s.eta3[3 * b] = e0; s.eta3[3 * b + 1] = e1; s.eta3[3 * b + 2] = e2; s.eta[b] = e1;
    if (satisfied) { s.status[b] = 0; s.active[b] = 0; }
    else if (e1 > 0.999 * e2) { s.status[b] = 1; s.active[b] = 0; }       // iLQGkl.jl:178
    else if (it >= max_iter) { s.status[b] = 3; s.active[b] = 0; }
    else { s.need_bp[b] = 1; atomicAdd(&s.counters[1], 1); }
}

__global__ void kl_export_kernel(long long B, KlSolveState s, const double* __restrict__ costnew, ddp_ilqgkl_state* out) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    ddp_ilqgkl_state r;
    r.eta_min = s.eta3[3 * b]; r.eta = s.eta3[3 * b + 1]; r.eta_max = s.eta3[3 * b + 2];
    r.del0 = s.del0[b]; r.divergence = s.div[b]; r.dcost = s.dcost[b]; r.expected = s.expected[b]; r.cost = costnew[b];
    r.iter = s.iter[b]; r.status = s.status[b]; r.retries = s.retries[b]; r.pad = 0;
    out[b] = r;
}

}  // namespace

extern "C" {

int ddp_ilqgkl_solve_f64(ddp_handle_t h, const ddp_model* model, const ddp_ilqgkl_opts* opts, const ddp_ilqgkl_args* a,
                         int32_t* n_outer) {
    if (!h) return DDP_ERR_INVALID;
    if (!model || !opts || !a || !a->x || !a->u || !a->cost || !a->K_prev.ptr || !a->Sig_prev.ptr || !a->Sigi_prev.ptr ||
        !a->fx_model.ptr || !a->R1.ptr || !a->xnew || !a->unew || !a->K || !a->k || !a->Sig || !a->Sigi || !a->Vx || !a->costnew ||
        !a->state) {
        h->err = "ddp_ilqgkl_solve_f64: missing argument";
        return DDP_ERR_INVALID;
    }
    if (model->kind != DDP_MODEL_LINEAR && model->kind != DDP_MODEL_PENDCART) {
        h->err = "ddp_ilqgkl_solve_f64: unknown model kind (arbitrary host callbacks cannot run on the device; no CPU fallback)";
        return DDP_ERR_UNSUPPORTED;
    }
    if (model->kind == DDP_MODEL_PENDCART && (h->n != 4 || h->m != 1)) { h->err = "pendcart model needs n == 4, m == 1"; return DDP_ERR_INVALID; }
    if (model->kind == DDP_MODEL_LINEAR && (!model->A.ptr || !model->Bm.ptr)) { h->err = "linear model needs A and B"; return DDP_ERR_INVALID; }
    if (!model->Q.ptr || !model->R.ptr) { h->err = "model needs Q and R"; return DDP_ERR_INVALID; }
    const int n = h->n, m = h->m, T = h->T;
    const long long B = h->B;
    const int max_iter = opts->max_iter > 0 ? opts->max_iter : 50;
    const int max_retries = opts->max_eta_retries > 0 ? opts->max_eta_retries : 200;
    cudaError_t err = cudaSuccess;
    DevBuf mem(h);
    KlSolveState s{};
    double *cx = nullptr, *cu = nullptr, *fxb = nullptr, *fub = nullptr, *zeros = nullptr;
    int hc[2];
    int it = 0;
    const unsigned gB = (unsigned)((B + 255) / 256);
    cudaStream_t st = h->stream;
    ModelD M;
    BackParams BP{};
    FwdParams FP{};
    KlParams KP{};
    struct DevBuf { void* p = nullptr; ~DevBuf() { if (p) cudaFree(p); } } sx;      // cache of the state covariances (kl_tile.cu)
    bool sx_filled = false;

    CUS(cudaSetDevice(h->device));
    CUS(mem.reserve((size_t)B * 8 * 40 + (size_t)B * T * (n + m) * 8 + (model->kind == DDP_MODEL_PENDCART ? (size_t)B * T * 160 : 0) + (1 << 16)));
    CUS(mem.alloc(&s.eta3, 3 * B)); CUS(mem.alloc(&s.eta, B)); CUS(mem.alloc(&s.del0, B)); CUS(mem.alloc(&s.div, B));
    CUS(mem.alloc(&s.dcost, B)); CUS(mem.alloc(&s.expected, B)); CUS(mem.alloc(&s.dV, 2 * B)); CUS(mem.alloc(&s.klmean, B));
    CUS(mem.alloc(&s.iter, B)); CUS(mem.alloc(&s.status, B)); CUS(mem.alloc(&s.retries, B)); CUS(mem.alloc(&s.diverge, B));
    CUS(mem.alloc(&s.active, B)); CUS(mem.alloc(&s.need_bp, B)); CUS(mem.alloc(&s.counters, 2));
    CUS(mem.alloc(&cx, (size_t)B * T * n)); CUS(mem.alloc(&cu, (size_t)B * T * m));
    CUS(mem.alloc(&zeros, (size_t)n * m));
    CUS(cudaMemsetAsync(zeros, 0, sizeof(double) * n * m, st));
    if (model->kind == DDP_MODEL_PENDCART) { CUS(mem.alloc(&fxb, (size_t)B * T * 16)); CUS(mem.alloc(&fub, (size_t)B * T * 4)); }

    M.kind = model->kind; M.A = mk(model->A); M.Bm = mk(model->Bm); M.Q = mk(model->Q); M.R = mk(model->R); M.goal = model->goal;
    for (int i = 0; i < 8; i++) M.p[i] = model->p[i];
    M.terminal_cost = model->terminal_cost ? 1 : 0;
    M.flags = model->flags;

    kl_init_kernel<<<gB, 256, 0, st>>>(B, s, opts->eta_bracket[0], opts->eta_bracket[1], opts->eta_bracket[2], opts->del0);
    h->launches++;

    // derivatives once, before the loop (iLQGkl.jl:88, quirk Q9)
    {
        const long long wtot = B * T;
        launch_df_cost(h, st, n, m, T, B, a->x, a->u, M.Q, M.R, M.goal, nullptr, cx, cu, (M.flags & DDP_MODEL_Q_DIAGONAL) != 0);
        if (model->kind == DDP_MODEL_PENDCART) {
            df_pendcart_kernel<<<(unsigned)((wtot + 127) / 128), 128, 0, st>>>(T, B, a->x, a->u, M.p[0], M.p[1], M.p[2], M.p[3], nullptr, fxb, fub);
            h->launches++;
        }
    }
    BP.n = n; BP.m = m; BP.T = T; BP.B = B;
    BP.cx = TensorD{cx, (long long)T * n, n}; BP.cu = TensorD{cu, (long long)T * m, m};
    BP.cxx = M.Q; BP.cuu = M.R; BP.cxu = TensorD{zeros, 0, 0};
    if (model->kind == DDP_MODEL_PENDCART) {
        BP.fx = TensorD{fxb, (long long)T * 16, 16}; BP.fu = TensorD{fub, (long long)T * 4, 4};
    } else {
        BP.fx = M.A; BP.fu = M.Bm;
    }
    BP.u = TensorD{a->u, (long long)T * m, m};
    BP.lambda = nullptr; BP.reg_type = 0; BP.lims = opts->lims; BP.active = s.need_bp;
    BP.Kp = mk(a->K_prev); BP.kp = TensorD{nullptr, 0, 0};          // k_prev := 0 (iLQGkl.jl:52)
    BP.Sip = mk(a->Sigi_prev); BP.eta = s.eta; BP.Quui = a->Sig;
    BP.diverge = s.diverge; BP.K = a->K; BP.k = a->k; BP.Vx = a->Vx; BP.Vxx = nullptr; BP.Vxx1 = a->Vxx1; BP.Quu = a->Sigi; BP.dV = s.dV;
    BP.qp = QPOpts{100, 1e-8, 1e-8, 0.6, 1e-22, 0.1};

    FP.n = n; FP.m = m; FP.T = T; FP.B = B; FP.model = M;
    FP.K = a->K; FP.k = a->k;
    FP.x0 = TensorD{a->x, (long long)T * n, 0};
    FP.x = TensorD{a->x, (long long)T * n, n}; FP.u = TensorD{a->u, (long long)T * m, m};
    FP.alpha = nullptr; FP.alpha_scalar = 1.0; FP.u_scale = 1.0;                       // forward pass with alpha = 1 (:134)
    FP.lims = opts->lims; FP.active = s.active;
    FP.xnew = a->xnew; FP.unew = a->unew; FP.cost = a->costnew; FP.cost_t = nullptr; FP.cx = nullptr; FP.cu = nullptr;

    KP.n = n; KP.m = m; KP.T = T; KP.B = B;
    KP.fx = mk(a->fx_model); KP.R1 = mk(a->R1); KP.Kp = mk(a->K_prev); KP.kp = TensorD{nullptr, 0, 0};
    KP.Sp = mk(a->Sig_prev); KP.Sip = mk(a->Sigi_prev);
    KP.xnew = a->xnew; KP.xold = a->x; KP.Kn = a->K; KP.kn = a->k; KP.Sn = a->Sig;
    KP.kl_t = nullptr; KP.kl_mean = s.klmean; KP.active = s.active;
    // forward_covariance's state block depends on fx_model and R1 only: keep it from the first eta iteration (kl_tile.cu MODE 1 / 2)
    if (opts->kl_step > 0.0 && !opts->no_covariance_cache && n == 32 && m == 8 && !(h->flags & 1u) && max_iter > 1) {
        // as many leading trajectories as fit (2 GiB are left alone); the rest is propagated in every iteration
        size_t free_b = 0, total_b = 0;
        const size_t per = (size_t)T * 528 * sizeof(double), margin = (size_t)2 << 30;
        long long count = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b > margin) count = (long long)std::min<size_t>((size_t)B, (free_b - margin) / per);
        if (count >= std::max<long long>(1, B / 8) && cudaMalloc(&sx.p, (size_t)count * per) != cudaSuccess) { cudaGetLastError(); sx.p = nullptr; }
        KP.Sx_tri = static_cast<double*>(sx.p);
        KP.sx_count = (sx.p && count < B) ? count : 0;
    }

    for (it = 1; it <= max_iter; it++) {
        // KL-augmented backward sweep with the eta-retry loop (iLQGkl.jl:97-124)
        for (;;) {
            CUS(cudaMemsetAsync(s.counters, 0, 2 * sizeof(int), st));
            {
                bool handled = false;
                int rc = 0;
                if (!(h->flags & 1u)) rc = launch_back_pass_tile(h, BP, true, &handled);
                if (!handled && rc == 0) rc = launch_back_pass_generic(h, BP, true);
                CUS((cudaError_t)rc);
            }
            kl_retry_kernel<<<gB, 256, 0, st>>>(B, s, max_retries);
            h->launches++;
            CUS(cudaMemcpyAsync(hc, s.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
            CUS(cudaStreamSynchronize(st));
            if (hc[0] == 0) break;
        }
        CUS((cudaError_t)run_fwd(h, FP));
        if (opts->kl_step > 0.0) {
            KP.sx_mode = KP.Sx_tri ? (sx_filled ? 2 : 1) : 0;
            CUS((cudaError_t)launch_kl_div(h, KP));
            sx_filled = true;
        }
        CUS(cudaMemsetAsync(s.counters, 0, 2 * sizeof(int), st));
        kl_eta_kernel<<<gB, 256, 0, st>>>(B, s, a->cost, a->costnew, opts->kl_step, it, max_iter);
        h->launches++;
        CUS(cudaMemcpyAsync(hc, s.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
        CUS(cudaStreamSynchronize(st));
        if (hc[1] == 0) break;
    }
    // traj_new.k = copy(u) (iLQGkl.jl:240-241, quirk Q11)
    CUS(cudaMemcpyAsync(a->k, a->unew, sizeof(double) * (size_t)B * T * m, cudaMemcpyDeviceToDevice, st));
    kl_export_kernel<<<gB, 256, 0, st>>>(B, s, a->costnew, a->state);
    h->launches++;
    CUS(cudaStreamSynchronize(st));
    if (n_outer) *n_outer = std::min(it, max_iter);
    return DDP_OK;
fail:
    h->err = std::string("ddp_ilqgkl_solve_f64: ") + cudaGetErrorString(err);
    return DDP_ERR_CUDA;
}

}  // extern "C"
