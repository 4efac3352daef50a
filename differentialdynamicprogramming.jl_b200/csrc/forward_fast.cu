// Specialised forward line-search rollouts.
//
//  fwd_lin32x8_kernel : n = 32, m = 8, linear model with time-invariant per-trajectory A, B
//                       (the headline workload).  One warp per trajectory; lane i owns state i and
//                       keeps row i of A, B and Q in registers for all T steps.  The step's K (2 KB),
//                       x, u, k are read with fully coalesced 16-byte loads, prefetched two steps
//                       ahead so ~5 KB per warp are always in flight (the kernel is HBM-bound: it
//                       streams K once).  cx = Q(x-goal), cu = R u of the NEW trajectory can be
//                       written in the same pass (the reference's df for this model,
//                       demo_linear.jl:38-39), which saves the next iteration a whole pass.
//  fwd_pend_kernel    : n = 4, m = 1 pendulum-on-cart (Euler step), ONE THREAD per trajectory.
//
// Replaces forward_pass of src/forward_pass.jl:9-33 (+ f/costfun of demo_linear.jl:35-50 and
// system_pendcart.jl:83-106).  Anything else dispatches to forward_generic.cu.
#include <cstdlib>
#include "ddp_common.cuh"

namespace {

__device__ __forceinline__ double2 ldg2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void stg2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

constexpr int FW_WPB = 4;
constexpr int FW_WARP_DOUBLES = 32 + 32 + 32 + 8;   // sx, sdx, sd, su  (per warp)
constexpr int FW_KD = 4;                            // depth of the cp.async ring of K_t (2 KB per step per warp)

struct Pre {
    double2 K0, K1, K2, K3;
    double2 u, k;
    double xo;
};

template <bool POLICY, bool WITHK = true>
__device__ __forceinline__ void prefetch(Pre& p, const FwdParams& P, long long b, int t, int lane, int q) {
    const int N = P.T;
    if (POLICY) {
        if (WITHK) {
            const double* Kt = P.K + (b * N + t) * 256 + 2 * lane;
            p.K0 = ldg2(Kt);
            p.K1 = ldg2(Kt + 64);
            p.K2 = ldg2(Kt + 128);
            p.K3 = ldg2(Kt + 192);
        }
        p.k = ldg2(P.k + (b * N + t) * 8 + 2 * q);
        p.xo = tp(P.x, b, t)[lane];
    }
    p.u = ldg2(tp(P.u, b, t) + 2 * q);
}

// QMODE 0: Q diagonal (only the diagonal is read), 1: dense Q shared by the batch, staged once per CTA
// in shared memory (column-major, conflict-free for "lane = row").
template <bool POLICY, int QMODE>
__global__ void __launch_bounds__(FW_WPB * 32, 2) fwd_lin32x8_kernel(FwdParams P) {
    __shared__ double smem[FW_WPB * FW_WARP_DOUBLES];
    __shared__ double sQ[QMODE == 1 ? 1024 : 1];
    __shared__ __align__(16) double sKring[POLICY ? FW_WPB * FW_KD * 256 : 2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    double* kring = sKring + (POLICY ? w * FW_KD * 256 : 0) + 2 * lane;      // this lane's 16-byte cells: + 64 i + 256 slot
    double* sx = smem + w * FW_WARP_DOUBLES;
    double* sdx = sx + 32;
    double* sd = sdx + 32;
    double* su = sd + 32;
    const int N = P.T;
    const bool has_goal = (P.model.goal != nullptr);
    const bool has_lims = (P.lims != nullptr);
    const long long warps_total = (long long)gridDim.x * FW_WPB;
    if (QMODE == 1) {
        for (int e = threadIdx.x; e < 1024; e += FW_WPB * 32) sQ[e] = P.model.Q.p[e];
        __syncthreads();
    }

    for (long long b = (long long)blockIdx.x * FW_WPB + w; b < P.B; b += warps_total) {
        if (P.active && !P.active[b]) continue;
        const double alpha = P.alpha ? P.alpha[b] : P.alpha_scalar;
        // ---- per-trajectory constants in registers: row `lane` of A and B; row (lane & 7) of R
        double Ar[32], Br[8], Rr[8];
        double qdiag = 0.0;
        {
            const double* A = P.model.A.p + b * P.model.A.sb;        // column-major: A[i + 32 j]
            const double* Bm = P.model.Bm.p + b * P.model.Bm.sb;
            const double* R = P.model.R.p + b * P.model.R.sb;
#pragma unroll
            for (int j = 0; j < 32; j++) Ar[j] = A[lane + 32 * j];
#pragma unroll
            for (int c = 0; c < 8; c++) { Br[c] = Bm[lane + 32 * c]; Rr[c] = R[(lane & 7) + 8 * c]; }
            if (QMODE == 0) qdiag = (P.model.Q.p + b * P.model.Q.sb)[lane * 33];
        }
        const double goal = has_goal ? P.model.goal[lane] : 0.0;
        double lo0 = 0, lo1 = 0, hi0 = 0, hi1 = 0;
        if (has_lims) { lo0 = P.lims[2 * q]; lo1 = P.lims[2 * q + 1]; hi0 = P.lims[8 + 2 * q]; hi1 = P.lims[8 + 2 * q + 1]; }
        double x = (P.x0.p + b * P.x0.sb)[lane];
        double cpart = 0.0;
        double* xnb = P.xnew + b * (long long)N * 32;
        double* unb = P.unew + b * (long long)N * 8;

        // `p` holds the operands of step t; they are copied out and `p` is refilled for step t+2 at once,
        // so two steps' worth of loads (~5 KB per warp) are always in flight and stay in registers.
        // K_t (2 KB per step) goes through a cp.async ring FW_KD steps deep: no registers are held while it is in flight,
        // and every lane reads back exactly the 16-byte cells it copied itself (no cross-lane hand-over).
        auto issue_K = [&](int t) {
            if (POLICY) {
                if (t < N) {
                    const double* Kt = P.K + (b * N + t) * 256 + 2 * lane;
                    double* dst = kring + (t % FW_KD) * 256;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const unsigned d = (unsigned)__cvta_generic_to_shared(dst + 64 * i);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(Kt + 64 * i));
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");      // one group per step, empty past the horizon
            }
        };
        auto step = [&](int t, Pre& p) {
            const double2 uu = p.u, kk = p.k;
            const double xo = p.xo;
            if (t + 2 < N) prefetch<POLICY, false>(p, P, b, t + 2, lane, q);
            double2 K0 = make_double2(0.0, 0.0), K1 = K0, K2 = K0, K3 = K0;
            if (POLICY) {
                issue_K(t + FW_KD - 1);
                asm volatile("cp.async.wait_group %0;" ::"n"(FW_KD - 1) : "memory");   // K_t has landed
                const double* src = kring + (t % FW_KD) * 256;
                K0 = ldg2(src); K1 = ldg2(src + 64); K2 = ldg2(src + 128); K3 = ldg2(src + 192);
            }
            // 1. publish x (and dx, d) to the warp
            const double d = x - goal;
            sx[lane] = x;
            if (POLICY) sdx[lane] = x - xo;
            if (has_goal && QMODE == 1) sd[lane] = d;
            xnb[(long long)t * 32 + lane] = x;
            __syncwarp();
            // 2. controls: u + alpha k + K dx   (lane holds K[2q..2q+1][g + 8i], i = 0..3)
            double un0 = __dmul_rn(uu.x, P.u_scale), un1 = __dmul_rn(uu.y, P.u_scale);
            if (POLICY) {
                const double d0 = sdx[g], d1 = sdx[g + 8], d2 = sdx[g + 16], d3 = sdx[g + 24];
                double p0 = fma(K3.x, d3, fma(K2.x, d2, fma(K1.x, d1, K0.x * d0)));
                double p1 = fma(K3.y, d3, fma(K2.y, d2, fma(K1.y, d1, K0.y * d0)));
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    p0 += __shfl_xor_sync(0xffffffffu, p0, o);
                    p1 += __shfl_xor_sync(0xffffffffu, p1, o);
                }
                un0 = __dadd_rn(__dadd_rn(un0, __dmul_rn(kk.x, alpha)), p0);   // forward_pass.jl:18,20: separate roundings, no FMA
                un1 = __dadd_rn(__dadd_rn(un1, __dmul_rn(kk.y, alpha)), p1);
            }
            if (has_lims) { un0 = clamp_jl(un0, lo0, hi0); un1 = clamp_jl(un1, lo1, hi1); }
            if (un0 != un0) un0 = 0.0;
            if (un1 != un1) un1 = 0.0;
            if (g == 0) {
                stg2(&su[2 * q], un0, un1);
                stg2(unb + (long long)t * 8 + 2 * q, un0, un1);
            }
            __syncwarp();
            // 3. x+ = A x + B u ; Qd ; Ru
            double ax0 = 0.0, ax1 = 0.0, ax2 = 0.0, ax3 = 0.0, qd0 = 0.0, qd1 = 0.0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const double2 xa = *reinterpret_cast<const double2*>(&sx[j]);
                const double2 xb = *reinterpret_cast<const double2*>(&sx[j + 2]);
                ax0 = fma(Ar[j], xa.x, ax0);
                ax1 = fma(Ar[j + 1], xa.y, ax1);
                ax2 = fma(Ar[j + 2], xb.x, ax2);
                ax3 = fma(Ar[j + 3], xb.y, ax3);
                if (QMODE == 1) {
                    if (!has_goal) {
                        qd0 = fma(sQ[lane + 32 * j], xa.x, qd0);
                        qd1 = fma(sQ[lane + 32 * (j + 1)], xa.y, qd1);
                        qd0 = fma(sQ[lane + 32 * (j + 2)], xb.x, qd0);
                        qd1 = fma(sQ[lane + 32 * (j + 3)], xb.y, qd1);
                    } else {
                        const double2 da = *reinterpret_cast<const double2*>(&sd[j]);
                        const double2 db = *reinterpret_cast<const double2*>(&sd[j + 2]);
                        qd0 = fma(sQ[lane + 32 * j], da.x, qd0);
                        qd1 = fma(sQ[lane + 32 * (j + 1)], da.y, qd1);
                        qd0 = fma(sQ[lane + 32 * (j + 2)], db.x, qd0);
                        qd1 = fma(sQ[lane + 32 * (j + 3)], db.y, qd1);
                    }
                }
            }
            double bu = 0.0, ru = 0.0;
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                const double2 uv = *reinterpret_cast<const double2*>(&su[c]);
                bu = fma(Br[c], uv.x, bu);
                bu = fma(Br[c + 1], uv.y, bu);
                ru = fma(Rr[c], uv.x, ru);
                ru = fma(Rr[c + 1], uv.y, ru);
            }
            const double qd = (QMODE == 0) ? qdiag * d : (qd0 + qd1);
            double cstep = 0.5 * d * qd;
            if (lane < 8) cstep = fma(0.5 * su[lane], ru, cstep);
            if (P.cx) P.cx[(b * N + t) * 32 + lane] = qd;
            if (P.cu && lane < 8) P.cu[(b * N + t) * 8 + lane] = ru;
            if (P.cost_t) {
                double ct = cstep;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ct += __shfl_xor_sync(0xffffffffu, ct, o);
                if (lane == 0) P.cost_t[b * (N + P.model.terminal_cost) + t] = ct;
            }
            cpart = __dadd_rn(cpart, cstep);
            if (t < N - 1) x = ((ax0 + ax1) + (ax2 + ax3)) + bu;
            __syncwarp();
        };

        Pre pa, pb;
        prefetch<POLICY, false>(pa, P, b, 0, lane, q);
        if (N > 1) prefetch<POLICY, false>(pb, P, b, 1, lane, q);
#pragma unroll
        for (int t = 0; t < FW_KD - 1; t++) issue_K(t);
        for (int t = 0; t < N; t += 2) {
            step(t, pa);
            if (t + 1 < N) step(t + 1, pb);
        }
        if (POLICY) asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (P.model.terminal_cost) {
            // ½ d'Qd at the last state once more (system_pendcart.jl:104)
            double qd = 0.0;
            if (QMODE == 0) qd = qdiag * (x - goal);
            else {
                const double* sv = has_goal ? sd : sx;
#pragma unroll
                for (int j = 0; j < 32; j++) qd = fma(sQ[lane + 32 * j], sv[j], qd);
            }
            double cterm = __dmul_rn(0.5 * (x - goal), qd);
            if (P.cost_t) {
                double ct = cterm;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ct += __shfl_xor_sync(0xffffffffu, ct, o);
                if (lane == 0) P.cost_t[b * (N + 1) + N] = ct;
            }
            cpart = __dadd_rn(cpart, cterm);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cpart += __shfl_xor_sync(0xffffffffu, cpart, o);
        if (lane == 0) P.cost[b] = cpart;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Multi-alpha line-search rollout for the headline shape: the total cost of the rollout
// for up to FM_NA step sizes in ONE pass over the policy, i.e. K (512 KB per trajectory, the dominant read of
// a forward pass) is streamed once instead of once per backtracking step of iLQG.jl:267-281.  Only costs are
// produced; the accepted step size is then rolled out by fwd_lin32x8_kernel, whose per-lane arithmetic
// (operation order, FMA placement, reduction trees) this kernel repeats exactly, so the costs agree bit for bit.
constexpr int FM_NA = 10;
constexpr int FM_WARP_DOUBLES = FM_NA * (32 + 32 + 32 + 8);
struct MultiAlpha {
    int na;
    double a[16];
};

template <int QMODE>
__global__ void __launch_bounds__(FW_WPB * 32, 2) fwd_lin32x8_multi_kernel(FwdParams P, MultiAlpha MA, double* __restrict__ cost_out) {
    extern __shared__ double fm_smem[];
    __shared__ double sQ[QMODE == 1 ? 1024 : 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    double* sx = fm_smem + w * FM_WARP_DOUBLES;          // [a][32]
    double* sdx = sx + FM_NA * 32;
    double* sd = sdx + FM_NA * 32;
    double* su = sd + FM_NA * 32;                         // [a][8]
    const int N = P.T, na = MA.na;
    const bool has_goal = (P.model.goal != nullptr);
    const bool has_lims = (P.lims != nullptr);
    const long long warps_total = (long long)gridDim.x * FW_WPB;
    if (QMODE == 1) {
        for (int e = threadIdx.x; e < 1024; e += FW_WPB * 32) sQ[e] = P.model.Q.p[e];
        __syncthreads();
    }
    for (long long b = (long long)blockIdx.x * FW_WPB + w; b < P.B; b += warps_total) {
        if (P.active && !P.active[b]) continue;
        double Ar[32], Br[8], Rr[8];
        double qdiag = 0.0;
        {
            const double* A = P.model.A.p + b * P.model.A.sb;
            const double* Bm = P.model.Bm.p + b * P.model.Bm.sb;
            const double* R = P.model.R.p + b * P.model.R.sb;
#pragma unroll
            for (int j = 0; j < 32; j++) Ar[j] = A[lane + 32 * j];
#pragma unroll
            for (int c = 0; c < 8; c++) { Br[c] = Bm[lane + 32 * c]; Rr[c] = R[(lane & 7) + 8 * c]; }
            if (QMODE == 0) qdiag = (P.model.Q.p + b * P.model.Q.sb)[lane * 33];
        }
        const double goal = has_goal ? P.model.goal[lane] : 0.0;
        double lo0 = 0, lo1 = 0, hi0 = 0, hi1 = 0;
        if (has_lims) { lo0 = P.lims[2 * q]; lo1 = P.lims[2 * q + 1]; hi0 = P.lims[8 + 2 * q]; hi1 = P.lims[8 + 2 * q + 1]; }
        double x[FM_NA], cpart[FM_NA];
        {
            const double x0 = (P.x0.p + b * P.x0.sb)[lane];
#pragma unroll
            for (int a = 0; a < FM_NA; a++) { x[a] = x0; cpart[a] = 0.0; }
        }
        Pre pa, pb;
        prefetch<true>(pa, P, b, 0, lane, q);
        if (N > 1) prefetch<true>(pb, P, b, 1, lane, q);
        auto step = [&](int t, Pre& p) {
            const double2 K0 = p.K0, K1 = p.K1, K2 = p.K2, K3 = p.K3, uu = p.u, kk = p.k;
            const double xo = p.xo;
            if (t + 2 < N) prefetch<true>(p, P, b, t + 2, lane, q);
#pragma unroll
            for (int a = 0; a < FM_NA; a++)
                if (a < na) {
                    sx[a * 32 + lane] = x[a];
                    sdx[a * 32 + lane] = x[a] - xo;
                    if (has_goal && QMODE == 1) sd[a * 32 + lane] = x[a] - goal;
                }
            __syncwarp();
#pragma unroll
            for (int a = 0; a < FM_NA; a++)
                if (a < na) {
                    double un0 = __dmul_rn(uu.x, P.u_scale), un1 = __dmul_rn(uu.y, P.u_scale);
                    const double* dxa = sdx + a * 32;
                    const double d0 = dxa[g], d1 = dxa[g + 8], d2 = dxa[g + 16], d3 = dxa[g + 24];
                    double p0 = fma(K3.x, d3, fma(K2.x, d2, fma(K1.x, d1, K0.x * d0)));
                    double p1 = fma(K3.y, d3, fma(K2.y, d2, fma(K1.y, d1, K0.y * d0)));
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {
                        p0 += __shfl_xor_sync(0xffffffffu, p0, o);
                        p1 += __shfl_xor_sync(0xffffffffu, p1, o);
                    }
                    un0 = __dadd_rn(__dadd_rn(un0, __dmul_rn(kk.x, MA.a[a])), p0);
                    un1 = __dadd_rn(__dadd_rn(un1, __dmul_rn(kk.y, MA.a[a])), p1);
                    if (has_lims) { un0 = clamp_jl(un0, lo0, hi0); un1 = clamp_jl(un1, lo1, hi1); }
                    if (un0 != un0) un0 = 0.0;
                    if (un1 != un1) un1 = 0.0;
                    if (g == 0) stg2(&su[a * 8 + 2 * q], un0, un1);
                }
            __syncwarp();
#pragma unroll
            for (int a = 0; a < FM_NA; a++)
                if (a < na) {
                    const double* sxa = sx + a * 32;
                    const double* sda = sd + a * 32;
                    const double* sua = su + a * 8;
                    const double d = x[a] - goal;
                    double ax0 = 0.0, ax1 = 0.0, ax2 = 0.0, ax3 = 0.0, qd0 = 0.0, qd1 = 0.0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const double2 xa = *reinterpret_cast<const double2*>(&sxa[j]);
                        const double2 xb = *reinterpret_cast<const double2*>(&sxa[j + 2]);
                        ax0 = fma(Ar[j], xa.x, ax0);
                        ax1 = fma(Ar[j + 1], xa.y, ax1);
                        ax2 = fma(Ar[j + 2], xb.x, ax2);
                        ax3 = fma(Ar[j + 3], xb.y, ax3);
                        if (QMODE == 1) {
                            if (!has_goal) {
                                qd0 = fma(sQ[lane + 32 * j], xa.x, qd0);
                                qd1 = fma(sQ[lane + 32 * (j + 1)], xa.y, qd1);
                                qd0 = fma(sQ[lane + 32 * (j + 2)], xb.x, qd0);
                                qd1 = fma(sQ[lane + 32 * (j + 3)], xb.y, qd1);
                            } else {
                                const double2 da = *reinterpret_cast<const double2*>(&sda[j]);
                                const double2 db = *reinterpret_cast<const double2*>(&sda[j + 2]);
                                qd0 = fma(sQ[lane + 32 * j], da.x, qd0);
                                qd1 = fma(sQ[lane + 32 * (j + 1)], da.y, qd1);
                                qd0 = fma(sQ[lane + 32 * (j + 2)], db.x, qd0);
                                qd1 = fma(sQ[lane + 32 * (j + 3)], db.y, qd1);
                            }
                        }
                    }
                    double bu = 0.0, ru = 0.0;
#pragma unroll
                    for (int c = 0; c < 8; c += 2) {
                        const double2 uv = *reinterpret_cast<const double2*>(&sua[c]);
                        bu = fma(Br[c], uv.x, bu);
                        bu = fma(Br[c + 1], uv.y, bu);
                        ru = fma(Rr[c], uv.x, ru);
                        ru = fma(Rr[c + 1], uv.y, ru);
                    }
                    const double qd = (QMODE == 0) ? qdiag * d : (qd0 + qd1);
                    double cstep = 0.5 * d * qd;
                    if (lane < 8) cstep = fma(0.5 * sua[lane], ru, cstep);
                    cpart[a] = __dadd_rn(cpart[a], cstep);
                    if (t < N - 1) x[a] = ((ax0 + ax1) + (ax2 + ax3)) + bu;
                }
            __syncwarp();
        };
        for (int t = 0; t < N; t += 2) {
            step(t, pa);
            if (t + 1 < N) step(t + 1, pb);
        }
#pragma unroll
        for (int a = 0; a < FM_NA; a++)
            if (a < na) {
                double c = cpart[a];
                if (P.model.terminal_cost) {
                    double qd = 0.0;
                    if (QMODE == 0) qd = qdiag * (x[a] - goal);
                    else {
                        const double* sv = has_goal ? (sd + a * 32) : (sx + a * 32);
#pragma unroll
                        for (int j = 0; j < 32; j++) qd = fma(sQ[lane + 32 * j], sv[j], qd);
                    }
                    c = __dadd_rn(c, __dmul_rn(0.5 * (x[a] - goal), qd));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                if (lane == 0) cost_out[(long long)a * P.B + b] = c;
            }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// pendulum on a cart: one thread per trajectory, everything in registers.  The rollout is a serial
// chain per trajectory (load -> control -> sincos -> next state), so the operands of the next PEND_PF
// steps are kept in flight in a register ring: the kernel is HBM-bound only if enough loads are pending.
// The control law and the Euler step of the pendulum on a cart, shared by the three rollout kernels below.  Every rounding
// is explicit (no FMA contraction), as in the reference's Julia arithmetic (forward_pass.jl:18-20, system_pendcart.jl:51-54),
// so the kernels agree with each other bit for bit whatever the compiler hoists or fuses around them.
__device__ __forceinline__ double pend_policy_control(double u_scaled, double kk, double alpha, double2 K01, double2 K23, double2 xo01,
                                                      double2 xo23, const double* x) {
    double un = __dadd_rn(u_scaled, __dmul_rn(kk, alpha));
    double acc = __dmul_rn(K01.x, x[0] - xo01.x);
    acc = __dadd_rn(acc, __dmul_rn(K01.y, x[1] - xo01.y));
    acc = __dadd_rn(acc, __dmul_rn(K23.x, x[2] - xo23.x));
    acc = __dadd_rn(acc, __dmul_rn(K23.y, x[3] - xo23.y));
    return __dadd_rn(un, acc);
}
__device__ __forceinline__ void pend_euler_step(double* x, double un, double gg, double l, double h, double dd) {
    double sn, cs2;
    sincos(x[0], &sn, &cs2);
    const double acc = __dsub_rn(__dadd_rn(__dmul_rn(-gg / l, sn), __dmul_rn(un / l, cs2)), __dmul_rn(dd, x[1]));
    const double x0n = __dadd_rn(x[0], __dmul_rn(h, x[1]));
    const double x1n = __dadd_rn(x[1], __dmul_rn(h, acc));
    const double x2n = __dadd_rn(x[2], __dmul_rn(h, x[3]));
    const double x3n = __dadd_rn(x[3], __dmul_rn(h, un));
    x[0] = x0n; x[1] = x1n; x[2] = x2n; x[3] = x3n;
}
// stage cost 1/2 d'Qd (+ 1/2 R u^2): qd = Q d is also the cost gradient cx
__device__ __forceinline__ double pend_state_cost(const double* x, const double* goal, const double* Q, double* qd) {
    double dlt[4];
#pragma unroll
    for (int i = 0; i < 4; i++) dlt[i] = x[i] - goal[i];
    double cs = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) s = __dadd_rn(s, __dmul_rn(Q[i + 4 * j], dlt[j]));
        qd[i] = s;
        cs = __dadd_rn(cs, __dmul_rn(__dmul_rn(0.5, dlt[i]), s));
    }
    return cs;
}

constexpr int PEND_PF = 4;
struct PendIn {
    double2 K01, K23, x01, x23;
    double k, u;
};
template <bool POLICY>
__device__ __forceinline__ void pend_load(PendIn& r, const FwdParams& P, long long b, int t) {
    const int N = P.T;
    r.u = tp(P.u, b, t)[0];
    if (POLICY) {
        const double* Kt = P.K + (b * N + t) * 4;
        const double* xo = tp(P.x, b, t);
        r.K01 = ldg2(Kt);
        r.K23 = ldg2(Kt + 2);
        r.x01 = ldg2(xo);
        r.x23 = ldg2(xo + 2);
        r.k = P.k[b * N + t];
    }
}

template <bool POLICY>
__global__ void __launch_bounds__(128) fwd_pend_kernel(FwdParams P) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.B) return;
    if (P.active && !P.active[b]) return;
    const int N = P.T;
    const double alpha = P.alpha ? P.alpha[b] : P.alpha_scalar;
    const double gg = P.model.p[0], l = P.model.p[1], h = P.model.p[2], dd = P.model.p[3];
    const double* Qm = P.model.Q.p + b * P.model.Q.sb;
    const double Rv = (P.model.R.p + b * P.model.R.sb)[0];
    double Q[16], goal[4], x[4];
#pragma unroll
    for (int i = 0; i < 16; i++) Q[i] = Qm[i];
#pragma unroll
    for (int i = 0; i < 4; i++) { goal[i] = P.model.goal ? P.model.goal[i] : 0.0; x[i] = (P.x0.p + b * P.x0.sb)[i]; }
    const bool has_lims = P.lims != nullptr;
    const double lo = has_lims ? P.lims[0] : 0.0, hi = has_lims ? P.lims[1] : 0.0;
    double* xnb = P.xnew + b * (long long)N * 4;
    double* unb = P.unew + b * (long long)N;
    double ctot = 0.0, clast = 0.0;
    PendIn ring[PEND_PF];
#pragma unroll
    for (int d = 0; d < PEND_PF; d++)
        if (d < N) pend_load<POLICY>(ring[d], P, b, d);
    for (int t0 = 0; t0 < N; t0 += PEND_PF) {
#pragma unroll
        for (int d = 0; d < PEND_PF; d++) {
            const int t = t0 + d;
            if (t < N) {
                const PendIn c = ring[d];
                if (t + PEND_PF < N) pend_load<POLICY>(ring[d], P, b, t + PEND_PF);
                double un = __dmul_rn(c.u, P.u_scale);
                if (POLICY) un = pend_policy_control(un, c.k, alpha, c.K01, c.K23, c.x01, c.x23, x);
                if (has_lims) un = clamp_jl(un, lo, hi);
                if (un != un) un = 0.0;
                stg2(xnb + (long long)t * 4, x[0], x[1]);
                stg2(xnb + (long long)t * 4 + 2, x[2], x[3]);
                unb[t] = un;
                double qd[4];
                const double cs = pend_state_cost(x, goal, Q, qd);
                clast = cs;
                const double ru = __dmul_rn(Rv, un);
                const double cstep = __dadd_rn(cs, __dmul_rn(__dmul_rn(0.5, un), ru));
                if (P.cx) { stg2(P.cx + (b * N + t) * 4, qd[0], qd[1]); stg2(P.cx + (b * N + t) * 4 + 2, qd[2], qd[3]); }
                if (P.cu) P.cu[b * N + t] = ru;
                if (P.cost_t) P.cost_t[b * (N + P.model.terminal_cost) + t] = cstep;
                ctot = __dadd_rn(ctot, cstep);
                if (t < N - 1) pend_euler_step(x, un, gg, l, h, dd);
            }
        }
    }
    if (P.model.terminal_cost) {
        if (P.cost_t) P.cost_t[b * (N + 1) + N] = clast;
        ctot += clast;
    }
    P.cost[b] = ctot;
}

// ---------------------------------------------------------------------------------------------
// pendulum on a cart, time-blocked staging (the fast path).  A thread walking its own trajectory touches
// one 32-byte sector per tensor per step, so every load/store instruction of a warp hits 32 different
// lines (32 L1 wavefronts): the LSU, not HBM, bounds fwd_pend_kernel.  Here a warp moves whole 128-byte
// lines instead: for each block of 4 steps the next K / x blocks (32 trajectories x 128 B) and k / u blocks
// (32 x 32 B) are copied with coalesced cp.async into a double-buffered shared-memory ring (rows padded to
// 144 B => conflict-free 16-byte row reads), the threads read their own rows, write xnew / unew (and cx, cu)
// back into the rows, and the block is written out with coalesced 16-byte stores.
constexpr int PS_W = 2;                                   // warps per CTA
constexpr int PS_R = 18;                                  // padded row of 16 doubles
constexpr int PS_WARP_DOUBLES = 2 * 32 * PS_R * 2 + 2 * 32 * 4 * 2 + 32 * PS_R + 32 * 4;   // K,X ring; k,u ring; cx; cu

__device__ __forceinline__ void cp16(double* dst_smem, const double* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src));
}

template <bool POLICY>
__global__ void __launch_bounds__(PS_W * 32) fwd_pend_staged_kernel(FwdParams P) {
    extern __shared__ __align__(16) double ps_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long b0 = ((long long)blockIdx.x * PS_W + wid) * 32;
    if (b0 >= P.B) return;
    const long long b_raw = b0 + lane;
    const bool valid = (b_raw < P.B) && !(P.active && !P.active[b_raw]);
    const unsigned amask = __ballot_sync(0xffffffffu, valid);
    const long long b = (b_raw < P.B) ? b_raw : P.B - 1;
    double* sm = ps_smem + (size_t)wid * PS_WARP_DOUBLES;
    double* sK = sm;                          // [2][32][18]
    double* sX = sK + 2 * 32 * PS_R;          // [2][32][18]   x of the old trajectory in, xnew out (in place)
    double* sk = sX + 2 * 32 * PS_R;          // [2][32][4]
    double* su = sk + 2 * 32 * 4;             // [2][32][4]    u in, unew out (in place)
    double* scx = su + 2 * 32 * 4;            // [32][18]
    double* scu = scx + 32 * PS_R;            // [32][4]
    const int N = P.T;
    const double alpha = P.alpha ? P.alpha[b] : P.alpha_scalar;
    const double gg = P.model.p[0], l = P.model.p[1], h = P.model.p[2], dd = P.model.p[3];
    const double* Qm = P.model.Q.p + b * P.model.Q.sb;
    const double Rv = (P.model.R.p + b * P.model.R.sb)[0];
    double Q[16], goal[4], x[4];
#pragma unroll
    for (int i = 0; i < 16; i++) Q[i] = Qm[i];
#pragma unroll
    for (int i = 0; i < 4; i++) { goal[i] = P.model.goal ? P.model.goal[i] : 0.0; x[i] = (P.x0.p + b * P.x0.sb)[i]; }
    const bool has_lims = P.lims != nullptr;
    const double lo = has_lims ? P.lims[0] : 0.0, hi = has_lims ? P.lims[1] : 0.0;
    const bool want_c = (P.cx != nullptr), want_cu = (P.cu != nullptr);
    // lane -> (row, 16-byte chunk) maps of the coalesced copies
    const int r8 = lane >> 3, c8 = lane & 7;              // 128-byte rows: 4 rows per instruction, 8 instructions
    const int r2 = lane >> 1, c2 = lane & 1;              // 32-byte rows: 16 rows per instruction, 2 instructions
    auto stage = [&](int j, int p) {
        const int t0 = 4 * j;
        if (t0 + (c8 >> 1) < N) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int row = r8 + 4 * k;
                const long long bb = b0 + row;
                if (bb < P.B) {
                    if (POLICY) {
                        cp16(sK + (p * 32 + row) * PS_R + 2 * c8, P.K + (bb * N + t0) * 4 + 2 * c8);
                        cp16(sX + (p * 32 + row) * PS_R + 2 * c8, tp(P.x, bb, t0) + 2 * c8);
                    }
                }
            }
        }
        if (t0 + 2 * c2 < N) {
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int row = r2 + 16 * k;
                const long long bb = b0 + row;
                if (bb < P.B) {
                    if (POLICY) cp16(sk + (p * 32 + row) * 4 + 2 * c2, P.k + bb * N + t0 + 2 * c2);
                    cp16(su + (p * 32 + row) * 4 + 2 * c2, tp(P.u, bb, t0) + 2 * c2);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double ctot = 0.0, clast = 0.0;
    const int nblk = (N + 3) >> 2;
    stage(0, 0);
    for (int j = 0; j < nblk; j++) {
        const int p = j & 1, t0 = 4 * j;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (j + 1 < nblk) stage(j + 1, p ^ 1);
        double* rK = sK + (p * 32 + lane) * PS_R;
        double* rX = sX + (p * 32 + lane) * PS_R;
        double* rk = sk + (p * 32 + lane) * 4;
        double* ru_ = su + (p * 32 + lane) * 4;
#pragma unroll
        for (int s_ = 0; s_ < 4; s_++) {
            const int t = t0 + s_;
            if (t < N) {
                double un = __dmul_rn(ru_[s_], P.u_scale);
                if (POLICY) {
                    const double2 K01 = ldg2(rK + 4 * s_), K23 = ldg2(rK + 4 * s_ + 2);
                    const double2 xo01 = ldg2(rX + 4 * s_), xo23 = ldg2(rX + 4 * s_ + 2);
                    un = pend_policy_control(un, rk[s_], alpha, K01, K23, xo01, xo23, x);
                }
                if (has_lims) un = clamp_jl(un, lo, hi);
                if (un != un) un = 0.0;
                stg2(rX + 4 * s_, x[0], x[1]);
                stg2(rX + 4 * s_ + 2, x[2], x[3]);
                ru_[s_] = un;
                double qd[4];
                const double cs = pend_state_cost(x, goal, Q, qd);
                clast = cs;
                const double ru = __dmul_rn(Rv, un);
                const double cstep = __dadd_rn(cs, __dmul_rn(__dmul_rn(0.5, un), ru));
                if (want_c) { stg2(scx + lane * PS_R + 4 * s_, qd[0], qd[1]); stg2(scx + lane * PS_R + 4 * s_ + 2, qd[2], qd[3]); }
                if (want_cu) scu[lane * 4 + s_] = ru;
                if (P.cost_t && valid) P.cost_t[b * (N + P.model.terminal_cost) + t] = cstep;
                ctot = __dadd_rn(ctot, cstep);
                if (t < N - 1) pend_euler_step(x, un, gg, l, h, dd);
            }
        }
        __syncwarp();
        // coalesced write-out of the block
        if (t0 + (c8 >> 1) < N) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int row = r8 + 4 * k;
                const long long bb = b0 + row;
                if ((amask >> row) & 1u) {
                    const double2 v = ldg2(sX + (p * 32 + row) * PS_R + 2 * c8);
                    stg2(P.xnew + (bb * N + t0) * 4 + 2 * c8, v.x, v.y);
                    if (want_c) {
                        const double2 c = ldg2(scx + row * PS_R + 2 * c8);
                        stg2(P.cx + (bb * N + t0) * 4 + 2 * c8, c.x, c.y);
                    }
                }
            }
        }
        if (t0 + 2 * c2 < N) {
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int row = r2 + 16 * k;
                const long long bb = b0 + row;
                if ((amask >> row) & 1u) {
                    const double2 v = ldg2(su + (p * 32 + row) * 4 + 2 * c2);
                    stg2(P.unew + bb * N + t0 + 2 * c2, v.x, v.y);
                    if (want_cu) {
                        const double2 c = ldg2(scu + row * 4 + 2 * c2);
                        stg2(P.cu + bb * N + t0 + 2 * c2, c.x, c.y);
                    }
                }
            }
        }
        __syncwarp();
    }
    if (P.model.terminal_cost) {
        if (P.cost_t && valid) P.cost_t[b * (N + 1) + N] = clast;
        ctot += clast;
    }
    if (valid) P.cost[b] = ctot;
}

// ---------------------------------------------------------------------------------------------
// pendulum on a cart, multi-alpha line-search rollout: the total cost for up to PM_NA step sizes in one pass over the
// policy (K, k) and the old trajectory, staged like fwd_pend_staged_kernel (no outputs but the costs).  The per-step
// expressions are the ones of fwd_pend_kernel / fwd_pend_staged_kernel, so the costs agree bit for bit with a rollout of
// the single step size.
constexpr int PM_NA = 8;
constexpr int PM_WARP_DOUBLES = 2 * 32 * PS_R * 2 + 2 * 32 * 4 * 2;      // K, X ring; k, u ring

__global__ void __launch_bounds__(PS_W * 32) fwd_pend_multi_kernel(FwdParams P, MultiAlpha MA, double* __restrict__ cost_out) {
    extern __shared__ __align__(16) double pm_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long b0 = ((long long)blockIdx.x * PS_W + wid) * 32;
    if (b0 >= P.B) return;
    const long long b_raw = b0 + lane;
    const bool valid = (b_raw < P.B) && !(P.active && !P.active[b_raw]);
    const long long b = (b_raw < P.B) ? b_raw : P.B - 1;
    double* sm = pm_smem + (size_t)wid * PM_WARP_DOUBLES;
    double* sK = sm;
    double* sX = sK + 2 * 32 * PS_R;
    double* sk = sX + 2 * 32 * PS_R;
    double* su = sk + 2 * 32 * 4;
    const int N = P.T, na = MA.na;
    const double gg = P.model.p[0], l = P.model.p[1], h = P.model.p[2], dd = P.model.p[3];
    const double* Qm = P.model.Q.p + b * P.model.Q.sb;
    const double Rv = (P.model.R.p + b * P.model.R.sb)[0];
    double Q[16], goal[4], x[PM_NA][4], ctot[PM_NA], clast[PM_NA];
#pragma unroll
    for (int i = 0; i < 16; i++) Q[i] = Qm[i];
#pragma unroll
    for (int i = 0; i < 4; i++) goal[i] = P.model.goal ? P.model.goal[i] : 0.0;
#pragma unroll
    for (int a = 0; a < PM_NA; a++) {
        ctot[a] = 0.0; clast[a] = 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++) x[a][i] = (P.x0.p + b * P.x0.sb)[i];
    }
    const bool has_lims = P.lims != nullptr;
    const double lo = has_lims ? P.lims[0] : 0.0, hi = has_lims ? P.lims[1] : 0.0;
    const int r8 = lane >> 3, c8 = lane & 7, r2 = lane >> 1, c2 = lane & 1;
    auto stage = [&](int j, int p) {
        const int t0 = 4 * j;
        if (t0 + (c8 >> 1) < N) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int row = r8 + 4 * k;
                const long long bb = b0 + row;
                if (bb < P.B) {
                    cp16(sK + (p * 32 + row) * PS_R + 2 * c8, P.K + (bb * N + t0) * 4 + 2 * c8);
                    cp16(sX + (p * 32 + row) * PS_R + 2 * c8, tp(P.x, bb, t0) + 2 * c8);
                }
            }
        }
        if (t0 + 2 * c2 < N) {
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int row = r2 + 16 * k;
                const long long bb = b0 + row;
                if (bb < P.B) {
                    cp16(sk + (p * 32 + row) * 4 + 2 * c2, P.k + bb * N + t0 + 2 * c2);
                    cp16(su + (p * 32 + row) * 4 + 2 * c2, tp(P.u, bb, t0) + 2 * c2);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int nblk = (N + 3) >> 2;
    stage(0, 0);
    for (int j = 0; j < nblk; j++) {
        const int p = j & 1, t0 = 4 * j;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (j + 1 < nblk) stage(j + 1, p ^ 1);
        const double* rK = sK + (p * 32 + lane) * PS_R;
        const double* rX = sX + (p * 32 + lane) * PS_R;
        const double* rk = sk + (p * 32 + lane) * 4;
        const double* ru_ = su + (p * 32 + lane) * 4;
#pragma unroll
        for (int s_ = 0; s_ < 4; s_++) {
            const int t = t0 + s_;
            if (t < N) {
                const double2 K01 = ldg2(rK + 4 * s_), K23 = ldg2(rK + 4 * s_ + 2);
                const double2 xo01 = ldg2(rX + 4 * s_), xo23 = ldg2(rX + 4 * s_ + 2);
                const double kk = rk[s_], uu = ru_[s_];
#pragma unroll
                for (int a = 0; a < PM_NA; a++)
                    if (a < na) {
                        const double alpha = MA.a[a];
                        double un = pend_policy_control(__dmul_rn(uu, P.u_scale), kk, alpha, K01, K23, xo01, xo23, x[a]);
                        if (has_lims) un = clamp_jl(un, lo, hi);
                        if (un != un) un = 0.0;
                        double qd[4];
                        const double cs = pend_state_cost(x[a], goal, Q, qd);
                        clast[a] = cs;
                        const double ru = __dmul_rn(Rv, un);
                        const double cstep = __dadd_rn(cs, __dmul_rn(__dmul_rn(0.5, un), ru));
                        ctot[a] = __dadd_rn(ctot[a], cstep);
                        if (t < N - 1) pend_euler_step(x[a], un, gg, l, h, dd);
                    }
            }
        }
        __syncwarp();
    }
    if (valid) {
#pragma unroll
        for (int a = 0; a < PM_NA; a++)
            if (a < na) {
                double c = ctot[a];
                if (P.model.terminal_cost) c = __dadd_rn(c, clast[a]);
                cost_out[(long long)a * P.B + b] = c;
            }
    }
}

bool al16(const void* p) { return ((uintptr_t)p % 16) == 0; }

}  // namespace

int launch_forward_fast(ddp_handle_s* h, const FwdParams& P, bool* handled) {
    *handled = false;
    if (P.lims && P.lims_st != 0) return 0;               // time-varying limits: generic kernel
    const bool policy = (P.K != nullptr);
    if (P.model.kind == DDP_MODEL_LINEAR && P.n == 32 && P.m == 8 && P.model.A.st == 0 && P.model.Bm.st == 0) {
        if (!al16(P.u.p) || (P.u.sb % 2) || (P.u.st % 2)) return 0;
        if (policy && (!al16(P.K) || !al16(P.k))) return 0;
        const bool qdiag = (P.model.flags & 1) != 0;
        if (!qdiag && P.model.Q.sb != 0) return 0;        // per-trajectory dense Q: generic kernel
        long long grid = (long long)h->sm_count * 2;
        long long need = (P.B + FW_WPB - 1) / FW_WPB;
        if (grid > need) grid = need;
        if (policy) {
            if (qdiag) fwd_lin32x8_kernel<true, 0><<<(unsigned)grid, FW_WPB * 32, 0, h->stream>>>(P);
            else fwd_lin32x8_kernel<true, 1><<<(unsigned)grid, FW_WPB * 32, 0, h->stream>>>(P);
        } else {
            if (qdiag) fwd_lin32x8_kernel<false, 0><<<(unsigned)grid, FW_WPB * 32, 0, h->stream>>>(P);
            else fwd_lin32x8_kernel<false, 1><<<(unsigned)grid, FW_WPB * 32, 0, h->stream>>>(P);
        }
        h->launches++;
        *handled = true;
        return (int)cudaGetLastError();
    }
    if (P.model.kind == DDP_MODEL_PENDCART && P.n == 4 && P.m == 1) {
        if (!al16(P.xnew) || (policy && (!al16(P.K) || !al16(P.x.p) || (P.x.sb % 2) || (P.x.st % 2)))) return 0;
        if (P.cx && !al16(P.cx)) return 0;
        // staged path: whole-line traffic; needs time-contiguous, 16-byte aligned rows (T even for the scalar rows)
        const bool staged = (P.T % 2 == 0) && al16(P.unew) && al16(P.u.p) && (P.u.sb % 2 == 0) && P.u.st == 1 &&
                            (!policy || (P.x.st == 4 && al16(P.k))) && (!P.cu || al16(P.cu)) && !getenv("DDP_PEND_NOSTAGE");
        if (staged) {
            const size_t bytes = (size_t)PS_W * PS_WARP_DOUBLES * sizeof(double);
            const unsigned sgrid = (unsigned)((P.B + 32 * PS_W - 1) / (32 * PS_W));
            cudaError_t e;
            if (policy) {
                e = cudaFuncSetAttribute(fwd_pend_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
                if (e == cudaSuccess) fwd_pend_staged_kernel<true><<<sgrid, PS_W * 32, bytes, h->stream>>>(P);
            } else {
                e = cudaFuncSetAttribute(fwd_pend_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
                if (e == cudaSuccess) fwd_pend_staged_kernel<false><<<sgrid, PS_W * 32, bytes, h->stream>>>(P);
            }
            if (e != cudaSuccess) return (int)e;
            h->launches++;
            *handled = true;
            return (int)cudaGetLastError();
        }
        unsigned grid = (unsigned)((P.B + 127) / 128);
        if (policy) fwd_pend_kernel<true><<<grid, 128, 0, h->stream>>>(P);
        else fwd_pend_kernel<false><<<grid, 128, 0, h->stream>>>(P);
        h->launches++;
        *handled = true;
        return (int)cudaGetLastError();
    }
    return 0;
}

// costs of the rollouts for na step sizes, cost_out (na, B) (row a = alpha[a]); *handled = false when the shape is
// not the headline one (the caller then loops over ddp_forward_pass launches)
int launch_forward_multi(ddp_handle_s* h, const FwdParams& P, int na, const double* alpha, double* cost_out, bool* handled) {
    *handled = false;
    if (P.lims && P.lims_st != 0) return 0;
    if (P.model.kind == DDP_MODEL_PENDCART && P.n == 4 && P.m == 1 && P.K != nullptr && na >= 1) {
        // same view requirements as the staged single-alpha kernel
        if (!((P.T % 2 == 0) && al16(P.u.p) && (P.u.sb % 2 == 0) && P.u.st == 1 && P.x.st == 4 && al16(P.k) && al16(P.K) && al16(P.x.p) &&
              (P.x.sb % 2 == 0))) return 0;
        const size_t bytes = (size_t)PS_W * PM_WARP_DOUBLES * sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(fwd_pend_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
        const unsigned sgrid = (unsigned)((P.B + 32 * PS_W - 1) / (32 * PS_W));
        for (int a0 = 0; a0 < na; a0 += PM_NA) {
            MultiAlpha MA;
            MA.na = (na - a0 < PM_NA) ? (na - a0) : PM_NA;
            for (int i = 0; i < 16; i++) MA.a[i] = (i < MA.na) ? alpha[a0 + i] : 0.0;
            fwd_pend_multi_kernel<<<sgrid, PS_W * 32, bytes, h->stream>>>(P, MA, cost_out + (long long)a0 * P.B);
            h->launches++;
        }
        *handled = true;
        return (int)cudaGetLastError();
    }
    if (!(P.model.kind == DDP_MODEL_LINEAR && P.n == 32 && P.m == 8 && P.model.A.st == 0 && P.model.Bm.st == 0)) return 0;
    if (P.K == nullptr || na < 1) return 0;
    {   // FP64 tensor-tile rollout of up to 16 step sizes at once (forward_multi_tile.cu); costs to rounding, not bit for bit
        int done = 0;
        for (int a0 = 0; a0 < na; a0 += 16) {
            bool hd = false;
            const int rc = launch_forward_multi_tile(h, P, (na - a0 < 16) ? (na - a0) : 16, alpha + a0, cost_out + (long long)a0 * P.B, &hd);
            if (rc != 0) return rc;
            if (!hd) break;
            done += 16;
        }
        if (done >= na) { *handled = true; return 0; }
    }
    if (!al16(P.u.p) || (P.u.sb % 2) || (P.u.st % 2) || !al16(P.K) || !al16(P.k)) return 0;
    const bool qdiag = (P.model.flags & 1) != 0;
    if (!qdiag && P.model.Q.sb != 0) return 0;
    long long grid = (long long)h->sm_count * 2;
    long long need = (P.B + FW_WPB - 1) / FW_WPB;
    if (grid > need) grid = need;
    const size_t bytes = (size_t)FW_WPB * FM_WARP_DOUBLES * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(fwd_lin32x8_multi_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fwd_lin32x8_multi_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    for (int a0 = 0; a0 < na; a0 += FM_NA) {
        MultiAlpha MA;
        MA.na = (na - a0 < FM_NA) ? (na - a0) : FM_NA;
        for (int i = 0; i < 16; i++) MA.a[i] = (i < MA.na) ? alpha[a0 + i] : 0.0;
        double* out = cost_out + (long long)a0 * P.B;
        if (qdiag) fwd_lin32x8_multi_kernel<0><<<(unsigned)grid, FW_WPB * 32, bytes, h->stream>>>(P, MA, out);
        else fwd_lin32x8_multi_kernel<1><<<(unsigned)grid, FW_WPB * 32, bytes, h->stream>>>(P, MA, out);
        h->launches++;
    }
    *handled = true;
    return (int)cudaGetLastError();
}
