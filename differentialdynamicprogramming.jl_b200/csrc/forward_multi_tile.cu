// Multi-alpha line-search rollout for n = 32, m = 8 on FP64 tensor tiles.
//
// The serial backtracking of iLQG.jl:267-281 re-runs forward_pass for one step size after the other.  Here the rollouts of up to 16
// step sizes of ONE trajectory advance together as the columns of a 32 x 16 state matrix X, so that every step is three small
// matrix products on mma.sync.m8n8k4.f64 ("DMMA", the instruction of the backward tile kernel) instead of 16 matrix-vector products
// on the FMA pipe, and the policy gains K_t -- 2 KB per step, the dominant read of a forward pass -- are streamed once:
//
//     U  = (u_t + k_t a') + K_t (X - x_t 1')      8 x 16      16 DMMA      a = the step sizes
//     U  = clamp(U, lims), NaN -> 0
//     c += 1/2 diag(X' Q X) + 1/2 diag(U' R U)               4 DMMA  (R U) + element-wise products (Q diagonal)
//     X+ = A X + B U                             32 x 16      64 + 16 DMMA
//
// Everything is held TRANSPOSED (rows = step sizes): in the accumulator layout of m8n8k4 a lane (g,q) owns X'[8nt+g][8mt+2q..2q+1],
// which is exactly the A operand of the next product when the contraction index is enumerated as 2q, 2q+1 (instead of q, q+4) on
// both operands -- the fragments of A, B, K, R are simply loaded in that order -- so the state never leaves the registers and is
// never re-laid-out between steps.  A (8 KB) sits in shared memory in fragment order; K_t, x_t, u_t, k_t arrive through a
// cp.async ring.  One warp per trajectory, 100 DMMA per step (NT = 2) = 1600 pipe cycles.
//
// Only the total cost of each step size is produced (the accepted one is then rolled out by fwd_lin32x8_kernel).  The sums are
// formed in tensor-tile order, so the costs agree with ddp_forward_pass_f64's to rounding (~1e-15 relative), not bit for bit.
// Restrictions (else the FMA kernel fwd_lin32x8_multi_kernel runs): Q diagonal and shared, R shared, no goal, no terminal cost.
#include <cstdlib>
#include "ddp_common.cuh"

namespace {

constexpr int FT_WPB = 4;                            // warps (trajectories) per CTA
constexpr int FT_KD = 3;                             // ring depth (steps in flight)
constexpr int FT_STEP = 256 + 32 + 8 + 8;            // K_t | x_t | u_t | k_t
constexpr int FT_A = 1024;                           // A in fragment order
constexpr int FT_WARP = FT_A + FT_KD * FT_STEP;      // doubles per warp (15.3 KB)

__device__ __forceinline__ void dmma_t(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cpa16(double* dst_smem, const double* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src));
}

struct MultiAlphaT {
    int na;
    double a[16];
};

template <int NT>
__global__ void __launch_bounds__(FT_WPB * 32, 3) fwd_lin32x8_multi_tile_kernel(FwdParams P, MultiAlphaT MA, double* __restrict__ cost_out) {
    extern __shared__ __align__(16) double ft_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    double* sA = ft_smem + (size_t)w * FT_WARP;       // [(mt2*4 + mt)*32 + lane][h]  =  A[8 mt2 + g][8 mt + 2q + h]
    double* ring = sA + FT_A;
    const int N = P.T;
    const bool has_lims = (P.lims != nullptr);
    const long long warps_total = (long long)gridDim.x * FT_WPB;
    double qd[4][2], lo[2] = {0.0, 0.0}, hi[2] = {0.0, 0.0}, RF[2];
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int h = 0; h < 2; h++) qd[mt][h] = P.model.Q.p[(8 * mt + 2 * q + h) * 33];     // diagonal of the shared Q
#pragma unroll
    for (int h = 0; h < 2; h++) {
        RF[h] = P.model.R.p[(2 * q + h) + 8 * g];                                            // B operand of U'R: R[2q+h][g]
        if (has_lims) { lo[h] = P.lims[2 * q + h]; hi[h] = P.lims[8 + 2 * q + h]; }
    }
    double al[NT];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) al[nt] = MA.a[8 * nt + g];

    for (long long b = (long long)blockIdx.x * FT_WPB + w; b < P.B; b += warps_total) {
        if (P.active && !P.active[b]) continue;
        __syncwarp();
        const double* A = P.model.A.p + b * P.model.A.sb;            // column-major: A[i + 32 j]
        const double* Bm = P.model.Bm.p + b * P.model.Bm.sb;
        const double* Kb = P.K + b * (long long)N * 256;
        const double* kb = P.k + b * (long long)N * 8;
        // ---- per-trajectory constants: A in fragment order (shared memory), B fragments (registers)
#pragma unroll
        for (int mt2 = 0; mt2 < 4; mt2++)
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                const double a0 = A[(8 * mt2 + g) + 32 * (8 * mt + 2 * q)], a1 = A[(8 * mt2 + g) + 32 * (8 * mt + 2 * q + 1)];
                *reinterpret_cast<double2*>(&sA[((mt2 * 4 + mt) * 32 + lane) * 2]) = make_double2(a0, a1);
            }
        double BF[4][2];
#pragma unroll
        for (int mt2 = 0; mt2 < 4; mt2++)
#pragma unroll
            for (int h = 0; h < 2; h++) BF[mt2][h] = Bm[(8 * mt2 + g) + 32 * (2 * q + h)];
        // ---- state: X'[8nt+g][8mt+2q+h], the same x0 for every step size
        double Xt[NT][4][2], cacc[NT];
        {
            const double* x0 = P.x0.p + b * P.x0.sb;
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                const double v0 = x0[8 * mt + 2 * q], v1 = x0[8 * mt + 2 * q + 1];
#pragma unroll
                for (int nt = 0; nt < NT; nt++) { Xt[nt][mt][0] = v0; Xt[nt][mt][1] = v1; }
            }
#pragma unroll
            for (int nt = 0; nt < NT; nt++) cacc[nt] = 0.0;
        }
        auto stage = [&](int t) {                     // K_t, x_t, u_t, k_t -> ring slot t % FT_KD (one commit group per call)
            if (t < N) {
                double* s = ring + (t % FT_KD) * FT_STEP;
                const double* Kt = Kb + (long long)t * 256 + 2 * lane;
#pragma unroll
                for (int i = 0; i < 4; i++) cpa16(s + 2 * lane + 64 * i, Kt + 64 * i);
                if (lane < 16) cpa16(s + 256 + 2 * lane, tp(P.x, b, t) + 2 * lane);
                else if (lane < 20) cpa16(s + 288 + 2 * (lane - 16), tp(P.u, b, t) + 2 * (lane - 16));
                else if (lane < 24) cpa16(s + 296 + 2 * (lane - 20), kb + (long long)t * 8 + 2 * (lane - 20));
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int t = 0; t < FT_KD - 1; t++) stage(t);
        __syncwarp();                                  // sA is complete for every lane
        for (int t = 0; t < N; t++) {
            asm volatile("cp.async.wait_group %0;" ::"n"(FT_KD - 2) : "memory");      // step t has landed (for this lane)
            __syncwarp();                              // ... for the whole warp; slot (t-1) % FT_KD is free again
            stage(t + FT_KD - 1);
            const double* s = ring + (t % FT_KD) * FT_STEP;
            const double2 uu = *reinterpret_cast<const double2*>(s + 288 + 2 * q);
            const double2 kk = *reinterpret_cast<const double2*>(s + 296 + 2 * q);
            // ---- U' = (u + k a')' + (X - x 1')' K'   : accumulator (g,q) = U'[8nt+g][2q..2q+1]; two partial accumulators per tile
            double Ut[NT][2], Ub[NT][2];
            {
                const double u0 = __dmul_rn(uu.x, P.u_scale), u1 = __dmul_rn(uu.y, P.u_scale);
#pragma unroll
                for (int nt = 0; nt < NT; nt++) {
                    Ut[nt][0] = __dadd_rn(u0, __dmul_rn(kk.x, al[nt]));           // forward_pass.jl:18: u + k*alpha, then + K*dx (:20)
                    Ut[nt][1] = __dadd_rn(u1, __dmul_rn(kk.y, al[nt]));
                    Ub[nt][0] = Ub[nt][1] = 0.0;
                }
            }
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                const double2 xo = *reinterpret_cast<const double2*>(s + 256 + 8 * mt + 2 * q);
                const double kf0 = s[g + 8 * (8 * mt + 2 * q)], kf1 = s[g + 8 * (8 * mt + 2 * q + 1)];     // K[g][8mt+2q+h]
#pragma unroll
                for (int nt = 0; nt < NT; nt++) {
                    if (mt & 1) {
                        dmma_t(Ub[nt][0], Ub[nt][1], Xt[nt][mt][0] - xo.x, kf0);
                        dmma_t(Ub[nt][0], Ub[nt][1], Xt[nt][mt][1] - xo.y, kf1);
                    } else {
                        dmma_t(Ut[nt][0], Ut[nt][1], Xt[nt][mt][0] - xo.x, kf0);
                        dmma_t(Ut[nt][0], Ut[nt][1], Xt[nt][mt][1] - xo.y, kf1);
                    }
                }
            }
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    double un = Ut[nt][h] + Ub[nt][h];
                    if (has_lims) un = clamp_jl(un, lo[h], hi[h]);               // forward_pass.jl:23
                    if (un != un) un = 0.0;                                       // u[isnan.(u)] .= 0 in f (demo_linear.jl:42)
                    Ut[nt][h] = un;
                }
            // ---- running cost 1/2 x'Qx + 1/2 u'Ru of every step size (demo_linear.jl:44): per-lane partial sums, reduced at the end
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                double v0 = 0.0, v1 = 0.0;                                        // (U'R)[8nt+g][2q..2q+1]
                dmma_t(v0, v1, Ut[nt][0], RF[0]);
                dmma_t(v0, v1, Ut[nt][1], RF[1]);
                double c = 0.0;
#pragma unroll
                for (int mt = 0; mt < 4; mt++) {
                    c = fma(qd[mt][0] * Xt[nt][mt][0], Xt[nt][mt][0], c);
                    c = fma(qd[mt][1] * Xt[nt][mt][1], Xt[nt][mt][1], c);
                }
                c = fma(v0, Ut[nt][0], c);
                c = fma(v1, Ut[nt][1], c);
                cacc[nt] = fma(0.5, c, cacc[nt]);
            }
            // ---- X+' = X' A' + U' B'  : 8 (NT = 2) independent accumulator tiles, the contraction index outermost
            double Xn[NT][4][2];
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int mt2 = 0; mt2 < 4; mt2++) Xn[nt][mt2][0] = Xn[nt][mt2][1] = 0.0;
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                double2 af[4];
#pragma unroll
                for (int mt2 = 0; mt2 < 4; mt2++) af[mt2] = *reinterpret_cast<const double2*>(&sA[((mt2 * 4 + mt) * 32 + lane) * 2]);
#pragma unroll
                for (int nt = 0; nt < NT; nt++)
#pragma unroll
                    for (int mt2 = 0; mt2 < 4; mt2++) dmma_t(Xn[nt][mt2][0], Xn[nt][mt2][1], Xt[nt][mt][0], af[mt2].x);
#pragma unroll
                for (int nt = 0; nt < NT; nt++)
#pragma unroll
                    for (int mt2 = 0; mt2 < 4; mt2++) dmma_t(Xn[nt][mt2][0], Xn[nt][mt2][1], Xt[nt][mt][1], af[mt2].y);
            }
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int nt = 0; nt < NT; nt++)
#pragma unroll
                    for (int mt2 = 0; mt2 < 4; mt2++) dmma_t(Xn[nt][mt2][0], Xn[nt][mt2][1], Ut[nt][h], BF[mt2][h]);
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int mt = 0; mt < 4; mt++) { Xt[nt][mt][0] = Xn[nt][mt][0]; Xt[nt][mt][1] = Xn[nt][mt][1]; }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        // ---- total cost of step size 8nt+g: sum of the 4 lanes of the group
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            double c = cacc[nt];
            c += __shfl_xor_sync(0xffffffffu, c, 1);
            c += __shfl_xor_sync(0xffffffffu, c, 2);
            if (q == 0 && 8 * nt + g < MA.na) cost_out[(long long)(8 * nt + g) * P.B + b] = c;
        }
        __syncwarp();
    }
}

bool a16(const void* p) { return ((uintptr_t)p % 16) == 0; }

}  // namespace

// Returns 0 with *handled = false when the shape / options are outside this kernel (the caller then uses the FMA kernel).
int launch_forward_multi_tile(ddp_handle_s* h, const FwdParams& P, int na, const double* alpha, double* cost_out, bool* handled) {
    *handled = false;
    if (getenv("DDP_MULTI_NO_TILE")) return 0;
    if (!(P.model.kind == DDP_MODEL_LINEAR && P.n == 32 && P.m == 8 && P.model.A.st == 0 && P.model.Bm.st == 0)) return 0;
    if (P.K == nullptr || na < 1 || na > 16) return 0;
    if (!(P.model.flags & 1) || P.model.Q.sb != 0 || P.model.R.sb != 0 || P.model.goal != nullptr || P.model.terminal_cost) return 0;
    if (!a16(P.K) || !a16(P.k) || !a16(P.u.p) || (P.u.sb % 2) || (P.u.st % 2) || !a16(P.x.p) || (P.x.sb % 2) || (P.x.st % 2)) return 0;
    if (P.lims && P.lims_st != 0) return 0;
    long long grid = (long long)h->sm_count * 3;                 // 3 CTAs (12 warps) per SM: 162 registers, 61 KB per CTA
    const long long need = (P.B + FT_WPB - 1) / FT_WPB;
    if (grid > need) grid = need;
    const size_t bytes = (size_t)FT_WPB * FT_WARP * sizeof(double);
    MultiAlphaT MA;
    MA.na = na;
    for (int i = 0; i < 16; i++) MA.a[i] = (i < na) ? alpha[i] : 0.0;
    cudaError_t e;
    if (na <= 8) {
        e = cudaFuncSetAttribute(fwd_lin32x8_multi_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
        fwd_lin32x8_multi_tile_kernel<1><<<(unsigned)grid, FT_WPB * 32, bytes, h->stream>>>(P, MA, cost_out);
    } else {
        e = cudaFuncSetAttribute(fwd_lin32x8_multi_tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
        fwd_lin32x8_multi_tile_kernel<2><<<(unsigned)grid, FT_WPB * 32, bytes, h->stream>>>(P, MA, cost_out);
    }
    h->launches++;
    *handled = true;
    return (int)cudaGetLastError();
}
