// Specialised backward sweep for small systems (n <= 4, m <= 2; config 3 is n = 4, m = 1 with
// control limits): ONE THREAD PER TRAJECTORY, every block of the step in registers, the next step's
// inputs prefetched into registers while the current one is processed.
//
// Replaces back_pass of src/backward_pass.jl:162-252 with both branches of @end_backward_pass
// (:31-42 Cholesky, :43-62 boxQP).  The factorisation, the triangular solves and the QP are the
// same sequential, FMA-free device functions as the generic kernel and the oracle (boxqp.cuh), so
// the integer outcomes (diverge, clamped set, QP result) follow the oracle's branch decisions.
//
// This shape is HBM/latency bound (0.4 Mflop vs 245 KB per trajectory-iteration).  The n x n block of
// fx (128 bytes per trajectory-step at n = 4) is the dominant read: a thread fetching its own line with
// eight 16-byte loads makes every load instruction touch 32 different lines (32 L1 wavefronts each),
// which saturates the LSU long before HBM.  With STAGE the warp instead copies the 32 lines of the next
// step with coalesced cp.async (lane -> (trajectory, 16-byte chunk): 4 whole lines per instruction) into
// a padded, conflict-free shared-memory ring and every thread then reads its own row from there.
#include <cstdlib>
#include "boxqp.cuh"

namespace {

template <int N, int M>
struct StepIn {
    double fu[N * M], cx[N], cu[M], u[M];
};

__device__ __forceinline__ void cp_async16s(double* dst_smem, const double* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src));
}

template <int N>
__device__ __forceinline__ void load_fx_direct(double* f, const BackParams& P, long long b, int i) {
    const double* fx = tp(P.fx, b, i);
#pragma unroll
    for (int e = 0; e < N * N; e++) f[e] = fx[e];
}

// rows of the staging ring: N*N doubles + one 16-byte pad => a thread's 16-byte reads of its own row are
// conflict-free (row stride 144 B at N = 4)
template <int N>
struct FxStage {
    static constexpr int ROW = N * N + 2;
    static constexpr int CH = (N * N) / 2;            // 16-byte chunks per row
};

// Coalesced copy of the warp's 32 fx blocks of step i into the ring: lane -> (row = lane / CH + (32 / CH) k, chunk = lane % CH).
// The per-lane source pointer (row, chunk) and the row validity mask are formed once per trajectory set; per step the
// address work is one 64-bit multiply-add plus one add per instruction.
template <int N>
__device__ __forceinline__ void stage_fx(double* sdst_lane, const double* src_lane, long long rowstep, unsigned rowmask) {
    constexpr int CH = FxStage<N>::CH, ROW = FxStage<N>::ROW, RPI = 32 / CH;    // rows per instruction
#pragma unroll
    for (int k = 0; k < CH; k++) {
        if ((rowmask >> k) & 1u) cp_async16s(sdst_lane + k * RPI * ROW, src_lane);
        src_lane += rowstep;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int N, int M>
__device__ __forceinline__ void load_step(StepIn<N, M>& s, const BackParams& P, long long b, int i, bool use_qp) {
    const double* fu = tp(P.fu, b, i);
    const double* cx = tp(P.cx, b, i);
    const double* cu = tp(P.cu, b, i);
    if ((N * M) % 2 == 0 && ((uintptr_t)fu % 16) == 0) {
#pragma unroll
        for (int e = 0; e < N * M; e += 2) { double2 t = *reinterpret_cast<const double2*>(fu + e); s.fu[e] = t.x; s.fu[e + 1] = t.y; }
    } else {
#pragma unroll
        for (int e = 0; e < N * M; e++) s.fu[e] = fu[e];
    }
    if (N % 2 == 0 && ((uintptr_t)cx % 16) == 0) {
#pragma unroll
        for (int e = 0; e < N; e += 2) { double2 t = *reinterpret_cast<const double2*>(cx + e); s.cx[e] = t.x; s.cx[e + 1] = t.y; }
        goto cx_done;
    }
#pragma unroll
    for (int e = 0; e < N; e++) s.cx[e] = cx[e];
cx_done:
#pragma unroll
    for (int e = 0; e < M; e++) { s.cu[e] = cu[e]; s.u[e] = use_qp ? tp(P.u, b, i)[e] : 0.0; }
}

template <int N, int M, int MINB, bool STAGE>
__global__ void __launch_bounds__(128, MINB) bp_small_kernel(BackParams P) {
    __shared__ __align__(16) double s_fx[STAGE ? 4 * 2 * 32 * FxStage<N>::ROW : 2];
    // cost Hessians shared by the batch and constant in time (the usual case): one copy per CTA, read by broadcast
    __shared__ double s_cost[N * N + N * M + M * M];
    const bool cost_shared = (P.cxx.sb == 0 && P.cxx.st == 0 && P.cxu.sb == 0 && P.cxu.st == 0 && P.cuu.sb == 0 && P.cuu.st == 0);
    if (cost_shared) {
        for (int e = threadIdx.x; e < N * N + N * M + M * M; e += blockDim.x)
            s_cost[e] = (e < N * N) ? P.cxx.p[e] : (e < N * N + N * M) ? P.cxu.p[e - N * N] : P.cuu.p[e - N * N - N * M];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long b_raw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long b0 = b_raw - lane;                       // first trajectory of this warp
    if (b0 >= P.B) return;                                   // warp-uniform
    const bool valid = (b_raw < P.B) && !(P.active && !P.active[b_raw]);
    const long long b = (b_raw < P.B) ? b_raw : P.B - 1;     // out-of-range lanes shadow the last trajectory, store nothing
    double* sfx = s_fx + (STAGE ? wid * 2 * 32 * FxStage<N>::ROW : 0);
    const int T = P.T;
    // staging maps (see stage_fx)
    constexpr int S_CH = FxStage<N>::CH, S_RPI = 32 / (S_CH > 0 ? S_CH : 1);
    const int s_row = lane / S_CH, s_ch = lane % S_CH;
    const double* fx_lane = P.fx.p + (b0 + s_row) * P.fx.sb + 2 * s_ch;
    const long long fx_rowstep = (long long)S_RPI * P.fx.sb;
    unsigned fx_rowmask = 0;
#pragma unroll
    for (int k = 0; k < S_CH; k++)
        if (b0 + s_row + (long long)S_RPI * k < P.B) fx_rowmask |= 1u << k;
    double* sfx_lane = sfx + s_row * FxStage<N>::ROW + 2 * s_ch;
    const bool use_qp = (P.lims != nullptr) && !(P.lims[0] > P.lims[M]);     // backward_pass.jl:31
    const double lam = P.lambda[b];
    const bool reg2 = (P.reg_type == 2), reg1 = (P.reg_type == 1);     // any other value: no regularisation, as `regType == 1 ? λ : 0`
    double* Kb = P.K + b * (long long)T * N * M;
    double* kb = P.k + b * (long long)T * M;
    double* Vxb = P.Vx + b * (long long)T * N;
    double* Vxxb = P.Vxx ? P.Vxx + b * (long long)T * N * N : nullptr;
    double* Quub = P.Quu ? P.Quu + b * (long long)T * M * M : nullptr;
    const bool vec_out = ((uintptr_t)P.K % 16 == 0) && ((uintptr_t)P.Vx % 16 == 0);
    double lims_lo[M], lims_hi[M];
#pragma unroll
    for (int a = 0; a < M; a++) { lims_lo[a] = use_qp ? P.lims[a] : 0.0; lims_hi[a] = use_qp ? P.lims[M + a] : 0.0; }

    double V[N * N], Vx[N];            // V column-major: V[r + N c]
    {
        const double* cxN = tp(P.cx, b, T - 1);
        const double* cxxN = tp(P.cxx, b, T - 1);
        const double* cuuN = tp(P.cuu, b, T - 1);
#pragma unroll
        for (int e = 0; e < N; e++) { Vx[e] = cxN[e]; if (valid) Vxb[(long long)(T - 1) * N + e] = Vx[e]; }
#pragma unroll
        for (int e = 0; e < N * N; e++) { V[e] = cxxN[e]; if (valid && Vxxb) Vxxb[(long long)(T - 1) * N * N + e] = V[e]; }
        if (valid) {
#pragma unroll
            for (int e = 0; e < N * M; e++) Kb[(long long)(T - 1) * N * M + e] = 0.0;
#pragma unroll
            for (int e = 0; e < M; e++) kb[(long long)(T - 1) * M + e] = 0.0;
        }
        if (valid && Quub)
#pragma unroll
            for (int e = 0; e < M * M; e++) Quub[(long long)(T - 1) * M * M + e] = cuuN[e];
    }
    double kw[M];
#pragma unroll
    for (int a = 0; a < M; a++) kw[a] = 0.0;
    double dV0 = 0.0, dV1 = 0.0;
    int diverge = 0;
    StepIn<N, M> cur, nxt;
    double cfx[N * N];
    bool alive = valid;
    if (T >= 2) {
        load_step<N, M>(cur, P, b, T - 2, use_qp);
        if (STAGE) stage_fx<N>(sfx_lane + ((T - 2) & 1) * 32 * FxStage<N>::ROW, fx_lane + (long long)(T - 2) * P.fx.st, fx_rowstep, fx_rowmask);
    }
    for (int i = T - 2; i >= 0; i--) {
        if (STAGE) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();                                            // the warp's copies of step i have landed
            const double* row = sfx + (i & 1) * 32 * FxStage<N>::ROW + lane * FxStage<N>::ROW;
#pragma unroll
            for (int e = 0; e < N * N; e += 2) { double2 t = *reinterpret_cast<const double2*>(row + e); cfx[e] = t.x; cfx[e + 1] = t.y; }
            // the other buffer was read in step i+1, before the __syncwarp above: refill it with step i-1
            if (i > 0) stage_fx<N>(sfx_lane + ((i - 1) & 1) * 32 * FxStage<N>::ROW, fx_lane + (long long)(i - 1) * P.fx.st, fx_rowstep, fx_rowmask);
        } else {
            load_fx_direct<N>(cfx, P, b, i);
        }
        if (i > 0) load_step<N, M>(nxt, P, b, i - 1, use_qp);        // in flight during this step's arithmetic
        if (alive) {
        const double* cxxi = cost_shared ? s_cost : tp(P.cxx, b, i);
        const double* cxui = cost_shared ? s_cost + N * N : tp(P.cxu, b, i);
        const double* cuui = cost_shared ? s_cost + N * N + N * M : tp(P.cuu, b, i);
        // ---- W = V fx, Z = V fu
        double W[N * N], Z[N * M];
#pragma unroll
        for (int c = 0; c < N; c++)
#pragma unroll
            for (int r = 0; r < N; r++) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < N; q++) acc = fma(V[r + N * q], cfx[q + N * c], acc);
                W[r + N * c] = acc;
            }
#pragma unroll
        for (int c = 0; c < M; c++)
#pragma unroll
            for (int r = 0; r < N; r++) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < N; q++) acc = fma(V[r + N * q], cur.fu[q + N * c], acc);
                Z[r + N * c] = acc;
            }
        // ---- Q expansion (backward_pass.jl:240-247)
        double Qxx[N * N], Qux[M * N], Quxr[M * N], Quu[M * M], QuuF[M * M], Qx[N], Qu[M];
#pragma unroll
        for (int c = 0; c < N; c++)
#pragma unroll
            for (int r = 0; r < N; r++) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < N; q++) acc = fma(cfx[q + N * r], W[q + N * c], acc);
                Qxx[r + N * c] = cxxi[r + N * c] + acc;
            }
#pragma unroll
        for (int j = 0; j < N; j++)
#pragma unroll
            for (int a = 0; a < M; a++) {
                double acc = 0.0, ff = 0.0;
#pragma unroll
                for (int q = 0; q < N; q++) { acc = fma(cur.fu[q + N * a], W[q + N * j], acc); ff = fma(cur.fu[q + N * a], cfx[q + N * j], ff); }
                const double v = cxui[j + N * a] + acc;
                Qux[a + M * j] = v;
                Quxr[a + M * j] = reg2 ? v + lam * ff : v;
            }
#pragma unroll
        for (int c = 0; c < M; c++)
#pragma unroll
            for (int a = 0; a < M; a++) {
                double acc = 0.0, ff = 0.0;
#pragma unroll
                for (int q = 0; q < N; q++) { acc = fma(cur.fu[q + N * a], Z[q + N * c], acc); ff = fma(cur.fu[q + N * a], cur.fu[q + N * c], ff); }
                const double v = cuui[a + M * c] + acc;
                Quu[a + M * c] = v;
                QuuF[a + M * c] = reg2 ? v + lam * ff : ((reg1 && a == c) ? v + lam : v);
            }
#pragma unroll
        for (int r = 0; r < N; r++) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < N; q++) acc = fma(cfx[q + N * r], Vx[q], acc);
            Qx[r] = cur.cx[r] + acc;
        }
#pragma unroll
        for (int a = 0; a < M; a++) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < N; q++) acc = fma(cur.fu[q + N * a], Vx[q], acc);
            Qu[a] = cur.cu[a] + acc;
        }
        // ---- gains: Cholesky or box QP, in the oracle's arithmetic order
        double ki[M], R[M * M], Ki[M * N];
        unsigned fm = (1u << M) - 1u;
        int nf = M;
        bool failed = false;
        if (!use_qp) {
            int idx[M];
#pragma unroll
            for (int a = 0; a < M; a++) idx[a] = a;
            if (!chol_upper_sub<M>(QuuF, M, idx, M, R, M)) failed = true;
            else {
#pragma unroll
                for (int a = 0; a < M; a++) ki[a] = Qu[a];
                chol_solve<M>(R, M, M, ki);
#pragma unroll
                for (int a = 0; a < M; a++) ki[a] = -ki[a];
            }
        } else {
            double lo[M], up[M];
#pragma unroll
            for (int a = 0; a < M; a++) { lo[a] = lims_lo[a] - cur.u[a]; up[a] = lims_hi[a] - cur.u[a]; }      // :45-46
            int nfac = 0;
            const int res = boxqp_seq<M>(M, QuuF, M, Qu, lo, up, kw, P.qp, ki, R, M, &fm, &nfac);
            if (res < 1) failed = true;                                                                  // :50-56
            nf = __popc(fm);
        }
        if (failed) { diverge = i + 1; alive = false; }
        else {
        if (M == 1) {
            // scalar control: K = -Qux_reg / (R'R); one reciprocal of the factor instead of 2 divisions per column
            const double rinv = (nf > 0) ? 1.0 / R[0] : 0.0;
#pragma unroll
            for (int j = 0; j < N; j++) Ki[j] = -((Quxr[j] * rinv) * rinv);
        } else {
#pragma unroll
        for (int j = 0; j < N; j++) {
            double v[M];
            int p = 0;
#pragma unroll
            for (int a = 0; a < M; a++)
                if ((fm >> a) & 1u) v[p++] = Quxr[a + M * j];
            if (nf > 0) chol_solve<M>(R, M, nf, v);
            p = 0;
#pragma unroll
            for (int a = 0; a < M; a++) Ki[a + M * j] = ((fm >> a) & 1u) ? -v[p++] : 0.0;
        }
        }
        // ---- value backup (:64-72)
        double Quuk[M], QK[M * N];
#pragma unroll
        for (int a = 0; a < M; a++) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < M; c++) acc = fma(Quu[a + M * c], ki[c], acc);
            Quuk[a] = acc;
        }
        {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int a = 0; a < M; a++) { a0 = fma(ki[a], Qu[a], a0); a1 = fma(ki[a], Quuk[a], a1); }
            dV0 += a0;
            dV1 += 0.5 * a1;
        }
#pragma unroll
        for (int j = 0; j < N; j++)
#pragma unroll
            for (int a = 0; a < M; a++) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < M; c++) acc = fma(Quu[a + M * c], Ki[c + M * j], acc);
                QK[a + M * j] = acc;
            }
#pragma unroll
        for (int r = 0; r < N; r++) {
            double t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
            for (int a = 0; a < M; a++) {
                t1 = fma(Ki[a + M * r], Quuk[a], t1);
                t2 = fma(Ki[a + M * r], Qu[a], t2);
                t3 = fma(Qux[a + M * r], ki[a], t3);
            }
            Vx[r] = ((Qx[r] + t1) + t2) + t3;
        }
#pragma unroll
        for (int c = 0; c < N; c++)
#pragma unroll
            for (int r = 0; r < N; r++) {
                double t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
                for (int a = 0; a < M; a++) {
                    t1 = fma(Ki[a + M * r], QK[a + M * c], t1);
                    t2 = fma(Ki[a + M * r], Qux[a + M * c], t2);
                    t3 = fma(Qux[a + M * r], Ki[a + M * c], t3);
                }
                W[r + N * c] = ((Qxx[r + N * c] + t1) + t2) + t3;
            }
#pragma unroll
        for (int c = 0; c < N; c++)
#pragma unroll
            for (int r = 0; r < N; r++) V[r + N * c] = 0.5 * (W[r + N * c] + W[c + N * r]);
        // ---- store
        if (vec_out && (N * M) % 2 == 0 && N % 2 == 0) {
#pragma unroll
            for (int e = 0; e < N * M; e += 2) *reinterpret_cast<double2*>(Kb + (long long)i * N * M + e) = make_double2(Ki[e], Ki[e + 1]);
#pragma unroll
            for (int r = 0; r < N; r += 2) *reinterpret_cast<double2*>(Vxb + (long long)i * N + r) = make_double2(Vx[r], Vx[r + 1]);
        } else {
#pragma unroll
            for (int e = 0; e < N * M; e++) Kb[(long long)i * N * M + e] = Ki[e];
#pragma unroll
            for (int r = 0; r < N; r++) Vxb[(long long)i * N + r] = Vx[r];
        }
#pragma unroll
        for (int a = 0; a < M; a++) { kb[(long long)i * M + a] = ki[a]; kw[a] = ki[a]; }
        if (Vxxb)
#pragma unroll
            for (int e = 0; e < N * N; e++) Vxxb[(long long)i * N * N + e] = V[e];
        if (Quub)
#pragma unroll
            for (int e = 0; e < M * M; e++) Quub[(long long)i * M * M + e] = Quu[e];
        }
        }
        cur = nxt;
    }
    if (!valid) return;
    if (diverge > 0) {                               // outputs below the failed step stay zero (quirk Q10)
        for (long long e = 0; e < (long long)diverge * N * M; e++) Kb[e] = 0.0;
        for (long long e = 0; e < (long long)diverge * M; e++) kb[e] = 0.0;
        for (long long e = 0; e < (long long)diverge * N; e++) Vxb[e] = 0.0;
        if (Vxxb)
            for (long long e = 0; e < (long long)diverge * N * N; e++) Vxxb[e] = 0.0;
    }
    if (P.Vxx1)
#pragma unroll
        for (int e = 0; e < N * N; e++) P.Vxx1[b * N * N + e] = (diverge > 0) ? 0.0 : V[e];
    P.diverge[b] = diverge;
    P.dV[2 * b] = dV0;
    P.dV[2 * b + 1] = dV1;
}

template <int N, int M>
int launch_small(ddp_handle_s* h, const BackParams& P) {
    const unsigned grid = (unsigned)((P.B + 127) / 128);
    const bool stage = ((N * N) % 2 == 0) && (32 % ((N * N) / 2) == 0) && ((uintptr_t)P.fx.p % 16 == 0) && (P.fx.sb % 2 == 0) && (P.fx.st % 2 == 0) &&
                       !(getenv("DDP_SMALL_NOSTAGE"));
    // 2 CTAs (8 warps) per SM: 168- and 128-register builds (3 / 4 CTAs) spill and measured 24-41 ms against 12 ms
    if (stage) bp_small_kernel<N, M, 2, true><<<grid, 128, 0, h->stream>>>(P);
    else bp_small_kernel<N, M, 2, false><<<grid, 128, 0, h->stream>>>(P);
    h->launches++;
    return (int)cudaGetLastError();
}

}  // namespace

int launch_back_pass_small(ddp_handle_s* h, const BackParams& P, bool gps, bool* handled) {
    *handled = false;
    if (gps || P.T < 2) return 0;
    if ((P.lims && P.lims_st != 0) || P.fxx.p || P.fxu.p || P.fuu.p || P.Vxx_tri || P.Quu_tri) return 0;   // generic kernel
    int rc = 0;
#define SMALL_CASE(NN, MM) if (P.n == NN && P.m == MM) { rc = launch_small<NN, MM>(h, P); *handled = true; return rc; }
    SMALL_CASE(4, 1)
    SMALL_CASE(2, 1)
    SMALL_CASE(3, 1)
    SMALL_CASE(4, 2)
#undef SMALL_CASE
    return rc;
}
