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

__device__ __forceinline__ void cp_async16s(double* dst_smem, const double* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async8s(double* dst_smem, const double* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src));
}

// One staged row = every per-step input of one trajectory: [fx N*N | fu N*M | cx N | cu M | u M], padded so that the row stride is
// 2 (mod 4) doubles: a thread's 16-byte reads of its own row are then bank-conflict free (quarter-warp strides cover all 8 bank groups).
template <int N, int M>
struct Row {
    static constexpr int FX = 0, FU = N * N, CX = FU + N * M, CU = CX + N, UU = CU + M, LEN = UU + M;
    static constexpr int ROW = LEN + ((2 - (LEN & 3)) & 3);
    // output block of one trajectory: OB steps of [K N*M | Vx N | k M], written out as whole lines every OB steps
    static constexpr int OB = 4;
    static constexpr int OK_ = 0, OVX = OB * N * M, OKK = OVX + OB * N, OLEN = OKK + OB * M;
    static constexpr int OROW = OLEN + ((2 - (OLEN & 3)) & 3);
    static constexpr int DEPTH = 2;      // ring buffers per warp: the inputs of step i are requested DEPTH-1 steps ahead (3 buffers were
                                         // measured slower: 11.3 vs 10.8 ms -- the larger ring takes the L1 capacity the row reads live on)
    static constexpr bool OK = (N % 2 == 0) && (32 % ((N * N) / 2) == 0) && (32 % ((N * M + 1) / 2) == 0) && ((N * M) % 2 == 0) &&
                               (32 % (OB * N * M / 2) == 0) && (32 % (OB * N / 2) == 0) && (32 % (OB * M / 2) == 0);
    static constexpr int WARP_DOUBLES = DEPTH * 32 * ROW + 32 * OROW;     // input ring + output block per warp
};

// symmetric n x n matrix kept as its upper triangle, column by column: (r,c), r <= c, at c(c+1)/2 + r
__host__ __device__ constexpr int tri(int r, int c) { return (r <= c) ? c * (c + 1) / 2 + r : r * (r + 1) / 2 + c; }

// Coalesced copy of one field of the warp's 32 trajectories at step i into the ring: CH 16-byte chunks per row, lane ->
// (row = lane / CH + (32 / CH) k, chunk = lane % CH): CH lanes read one contiguous run of one trajectory.
// `src` is the lane's source pointer for (row, chunk) AT THE STEP BEING STAGED (the caller walks it back by the time stride every
// step), `rowstep` the byte distance of RPI trajectories: per copy one 64-bit add, no multiplies.
template <int CH>
__device__ __forceinline__ void stage_field(double* dst, int ROWLEN, const char* src, long long rowstep, long long row_abs, long long B, bool full) {
    constexpr int RPI = 32 / CH;
    if (full) {                                            // warp-uniform: every row of the warp exists, no predicates
#pragma unroll
        for (int k = 0; k < CH; k++) {
            cp_async16s(dst + k * RPI * ROWLEN, reinterpret_cast<const double*>(src));
            src += rowstep;
        }
    } else {
#pragma unroll
        for (int k = 0; k < CH; k++) {
            if (row_abs + RPI * k < B) cp_async16s(dst + k * RPI * ROWLEN, reinterpret_cast<const double*>(src));
            src += rowstep;
        }
    }
}

// 1/d for d > 0 in the normal range, to the last bit or two: hardware seed + two Newton steps (the gains K are a 1e-8 quantity,
// not part of the bit-exact box-QP arithmetic)
__device__ __forceinline__ double rcp_fast(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// Coalesced write-out of one field of the warp's output block: CH 16-byte chunks per row (= OB steps of the field), lane ->
// (row = lane / CH + (32 / CH) k, chunk = lane % CH): CH lanes write one contiguous run of one trajectory; `nch` = chunks that exist
// (the topmost block of a horizon that is not a multiple of OB is partial), `vmask` = trajectories of the warp that store at all.
template <int CH>
__device__ __forceinline__ void flush_field(const double* srow0, int OROWLEN, double* g0, long long stride_b, int nch, unsigned vmask, int lane) {
    constexpr int RPI = 32 / CH;
    const int row = lane / CH, ch = lane % CH;
    if (ch >= nch) return;
#pragma unroll
    for (int k = 0; k < CH; k++) {
        const int r = row + RPI * k;
        if ((vmask >> r) & 1u)
            *reinterpret_cast<double2*>(g0 + (long long)r * stride_b + 2 * ch) = *reinterpret_cast<const double2*>(srow0 + r * OROWLEN + 2 * ch);
    }
}

// One thread per trajectory.  STAGE: every per-step input arrives through the warp's double-buffered cp.async ring (see Row);
// !STAGE (odd sizes, unaligned views, DDP_SMALL_NOSTAGE): direct loads with a one-step register prefetch.  Same arithmetic either way.
template <int N, int M, int MINB, bool STAGE, bool CSH>
__global__ void __launch_bounds__(128, MINB) bp_small_kernel(BackParams P) {
    using RW = Row<N, M>;
    constexpr int NT = N * (N + 1) / 2;
    extern __shared__ __align__(16) double s_ring[];         // STAGE: 4 warps x DEPTH buffers x 32 rows (dynamic: above the 48 KB static limit)
    // cost Hessians shared by the batch and constant in time (the usual case): one copy per CTA, read by broadcast; cxx symmetrised
    __shared__ double s_cost[NT + N * M + M * M];
    constexpr bool cost_shared = CSH;      // cxx, cxu, cuu shared by the batch and constant in time (decided by the launcher)
    if (cost_shared) {
        for (int e = threadIdx.x; e < NT + N * M + M * M; e += blockDim.x) {
            double v;
            if (e < NT) {
                int c = 0;
                while ((c + 1) * (c + 2) / 2 <= e) c++;
                const int r = e - c * (c + 1) / 2;
                v = 0.5 * (P.cxx.p[r + N * c] + P.cxx.p[c + N * r]);
            } else v = (e < NT + N * M) ? P.cxu.p[e - NT] : P.cuu.p[e - NT - N * M];
            s_cost[e] = v;
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long b_raw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long b0 = b_raw - lane;                       // first trajectory of this warp
    if (b0 >= P.B) return;                                   // warp-uniform
    bool valid = (b_raw < P.B) && !(P.active && !P.active[b_raw]);
    const long long b = (b_raw < P.B) ? b_raw : P.B - 1;     // out-of-range lanes shadow the last trajectory, store nothing
    const bool full = (b0 + 32 <= P.B);                      // warp-uniform: no row predicates in the staging copies
    double* ring = s_ring + (STAGE ? wid * RW::WARP_DOUBLES : 0);
    double* oblk = ring + RW::DEPTH * 32 * RW::ROW;         // STAGE: the warp's output block, row = trajectory
    double* orow = oblk + lane * RW::OROW;
    const int T = P.T;
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

    // staging: per field one lane pointer (row = lane / CH, chunk = lane % CH) positioned at step T-2 and walked back one time
    // stride per staged step; the shared-memory destinations are compile-time offsets from two lane bases
    constexpr int CHX = STAGE ? (N * N) / 2 : 1, CHU = STAGE ? (N * M) / 2 : 1, CHC = STAGE ? N / 2 : 1;
    const char *sp_fx = nullptr, *sp_fu = nullptr, *sp_cx = nullptr, *sp_cu = nullptr, *sp_u = nullptr;
    if (STAGE) {
        const long long t0 = T - 2;
        sp_fx = reinterpret_cast<const char*>(P.fx.p + (b0 + lane / CHX) * P.fx.sb + t0 * P.fx.st + 2 * (lane % CHX));
        sp_fu = reinterpret_cast<const char*>(P.fu.p + (b0 + lane / CHU) * P.fu.sb + t0 * P.fu.st + 2 * (lane % CHU));
        sp_cx = reinterpret_cast<const char*>(P.cx.p + (b0 + lane / CHC) * P.cx.sb + t0 * P.cx.st + 2 * (lane % CHC));
        sp_cu = reinterpret_cast<const char*>(P.cu.p + b * P.cu.sb + t0 * P.cu.st);
        sp_u = use_qp ? reinterpret_cast<const char*>(P.u.p + b * P.u.sb + t0 * P.u.st) : nullptr;
    }
    // byte strides of the staging walk, pinned in registers (otherwise re-read from the constant bank every step, with a scoreboard
    // wait in front of every address update)
    long long rs_fx = (32 / CHX) * 8 * P.fx.sb, rs_fu = (32 / CHU) * 8 * P.fu.sb, rs_cx = (32 / CHC) * 8 * P.cx.sb;
    long long ts_fx = 8 * P.fx.st, ts_fu = 8 * P.fu.st, ts_cx = 8 * P.cx.st, ts_cu = 8 * P.cu.st, ts_u = 8 * P.u.st;
    asm volatile("" : "+l"(rs_fx), "+l"(rs_fu), "+l"(rs_cx), "+l"(ts_fx), "+l"(ts_fu), "+l"(ts_cx), "+l"(ts_cu), "+l"(ts_u));
    auto stage = [&](int i) {                                // all inputs of step i -> ring buffer (i % DEPTH); steps are staged in order T-2, T-3, ...
        if (i < 0) { asm volatile("cp.async.commit_group;" ::: "memory"); return; }      // past the first step: an empty group keeps the wait count uniform
        double* r0 = ring + (i % RW::DEPTH) * 32 * RW::ROW;
        stage_field<CHX>(r0 + RW::FX + (lane / CHX) * RW::ROW + 2 * (lane % CHX), RW::ROW, sp_fx, rs_fx, b0 + lane / CHX, P.B, full);
        stage_field<CHU>(r0 + RW::FU + (lane / CHU) * RW::ROW + 2 * (lane % CHU), RW::ROW, sp_fu, rs_fu, b0 + lane / CHU, P.B, full);
        stage_field<CHC>(r0 + RW::CX + (lane / CHC) * RW::ROW + 2 * (lane % CHC), RW::ROW, sp_cx, rs_cx, b0 + lane / CHC, P.B, full);
        {
            double* rr = r0 + lane * RW::ROW;                // out-of-range lanes shadow the last trajectory (b): a valid address
            const double* cus = reinterpret_cast<const double*>(sp_cu);
            if (M % 2 == 0) {
#pragma unroll
                for (int a = 0; a < M; a += 2) cp_async16s(rr + RW::CU + a, cus + a);
            } else {
#pragma unroll
                for (int a = 0; a < M; a++) cp_async8s(rr + RW::CU + a, cus + a);
            }
            if (use_qp) {
                const double* us = reinterpret_cast<const double*>(sp_u);
                if (M % 2 == 0) {
#pragma unroll
                    for (int a = 0; a < M; a += 2) cp_async16s(rr + RW::UU + a, us + a);
                } else {
#pragma unroll
                    for (int a = 0; a < M; a++) cp_async8s(rr + RW::UU + a, us + a);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        sp_fx -= ts_fx; sp_fu -= ts_fu; sp_cx -= ts_cx; sp_cu -= ts_cu;
        if (use_qp) sp_u -= ts_u;
    };

    double Vs[NT], Vx[N];              // Vxx(i+1): upper triangle (exactly symmetric, backward_pass.jl:71-72)
    {
        const double* cxN = tp(P.cx, b, T - 1);
        const double* cxxN = tp(P.cxx, b, T - 1);
        const double* cuuN = tp(P.cuu, b, T - 1);
        // The products below assume Vxx = Vxx'.  A terminal cxx that is not exactly symmetric (the reference takes it as it is,
        // backward_pass.jl:22) is handed to the generic kernel (see BackParams::redo).
        bool asym = false;
#pragma unroll
        for (int c = 0; c < N; c++)
#pragma unroll
            for (int r = 0; r <= c; r++) {
                const double v = cxxN[r + N * c];
                Vs[tri(r, c)] = v;
                if (r < c && v != cxxN[c + N * r]) asym = true;
                if (v != v) asym = true;
            }
        if (valid) {
            P.redo[b] = asym ? 1 : 0;
            if (asym) atomicAdd(P.redo_count, 1);
        }
        if (asym) valid = false;
#pragma unroll
        for (int e = 0; e < N; e++) {
            Vx[e] = cxN[e];
            if (STAGE) orow[RW::OVX + ((T - 1) & (RW::OB - 1)) * N + e] = Vx[e];
            else if (valid) Vxb[(long long)(T - 1) * N + e] = Vx[e];
        }
        if (valid && Vxxb)
#pragma unroll
            for (int e = 0; e < N * N; e++) Vxxb[(long long)(T - 1) * N * N + e] = Vs[tri(e % N, e / N)];
        if (STAGE) {
#pragma unroll
            for (int e = 0; e < N * M; e++) orow[RW::OK_ + ((T - 1) & (RW::OB - 1)) * N * M + e] = 0.0;
#pragma unroll
            for (int e = 0; e < M; e++) orow[RW::OKK + ((T - 1) & (RW::OB - 1)) * M + e] = 0.0;
        } else if (valid) {
#pragma unroll
            for (int e = 0; e < N * M; e++) Kb[(long long)(T - 1) * N * M + e] = 0.0;
#pragma unroll
            for (int e = 0; e < M; e++) kb[(long long)(T - 1) * M + e] = 0.0;
        }
        if (valid && Quub)
#pragma unroll
            for (int e = 0; e < M * M; e++) Quub[(long long)(T - 1) * M * M + e] = cuuN[e];
    }
    // write-out of the output block that holds step i (steps OB*j .. OB*j+OB-1, j = i / OB), by the whole warp
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const bool kvec = (((long long)T * M) % 2 == 0) && ((uintptr_t)P.k % 16 == 0);     // 16-byte stores of k need even trajectory strides
    auto flush = [&](int i) {
        if (!STAGE) return;
        const int j = i / RW::OB;
        const int steps = (T - RW::OB * j < RW::OB) ? (T - RW::OB * j) : RW::OB;      // the topmost block may be partial
        __syncwarp();
        flush_field<RW::OB * N * M / 2>(oblk + RW::OK_, RW::OROW, P.K + b0 * (long long)T * N * M + (long long)RW::OB * j * N * M, (long long)T * N * M,
                                        steps * N * M / 2, vmask, lane);
        flush_field<RW::OB * N / 2>(oblk + RW::OVX, RW::OROW, P.Vx + b0 * (long long)T * N + (long long)RW::OB * j * N, (long long)T * N, steps * N / 2, vmask, lane);
        if ((RW::OB * M) % 2 == 0 && (steps * M) % 2 == 0 && kvec)
            flush_field<RW::OB * M / 2>(oblk + RW::OKK, RW::OROW, P.k + b0 * (long long)T * M + (long long)RW::OB * j * M, (long long)T * M, steps * M / 2, vmask, lane);
        else if (valid)                                   // an odd number of k entries in a partial block: the owner stores them
            for (int e = 0; e < steps * M; e++) kb[(long long)RW::OB * j * M + e] = orow[RW::OKK + e];
        __syncwarp();
    };
    if (STAGE && ((T - 1) & (RW::OB - 1)) == 0) flush(T - 1);       // the terminal step is alone in its block
    double kw[M];
#pragma unroll
    for (int a = 0; a < M; a++) kw[a] = 0.0;
    double dV0 = 0.0, dV1 = 0.0;
    int diverge = 0;
    bool alive = valid;
    // !STAGE: one-step register prefetch of the inputs
    double pf[STAGE ? 1 : RW::LEN];
    auto load_direct = [&](int i) {
        const double* fx = tp(P.fx, b, i);
        const double* fu = tp(P.fu, b, i);
        const double* cx = tp(P.cx, b, i);
        const double* cu = tp(P.cu, b, i);
#pragma unroll
        for (int e = 0; e < N * N; e++) pf[(STAGE ? 0 : RW::FX + e)] = fx[e];
#pragma unroll
        for (int e = 0; e < N * M; e++) pf[(STAGE ? 0 : RW::FU + e)] = fu[e];
#pragma unroll
        for (int e = 0; e < N; e++) pf[(STAGE ? 0 : RW::CX + e)] = cx[e];
#pragma unroll
        for (int e = 0; e < M; e++) { pf[(STAGE ? 0 : RW::CU + e)] = cu[e]; pf[(STAGE ? 0 : RW::UU + e)] = use_qp ? tp(P.u, b, i)[e] : 0.0; }
    };
    if (T >= 2) {
        if (STAGE) {
#pragma unroll
            for (int d = 0; d < RW::DEPTH - 1; d++) stage(T - 2 - d);
        }
        else load_direct(T - 2);
    }
    for (int i = T - 2; i >= 0; i--) {
        double in[RW::LEN];                                          // this step's [fx | fu | cx | cu | u]
        if (STAGE) {
            asm volatile("cp.async.wait_group %0;" ::"n"(RW::DEPTH - 2) : "memory");      // all but the newest DEPTH-2 groups: step i has landed
            __syncwarp();                                            // ... for every lane of the warp
            const double* row = ring + (i % RW::DEPTH) * 32 * RW::ROW + lane * RW::ROW;
#pragma unroll
            for (int e = 0; e + 1 < RW::LEN; e += 2) { const double2 t = *reinterpret_cast<const double2*>(row + e); in[e] = t.x; in[e + 1] = t.y; }
            if (RW::LEN & 1) in[RW::LEN - 1] = row[RW::LEN - 1];
            // the buffer of step i+1 was read before the __syncwarp above: refill it with step i-(DEPTH-1)
            stage(i - (RW::DEPTH - 1));
        } else {
#pragma unroll
            for (int e = 0; e < RW::LEN; e++) in[e] = pf[STAGE ? 0 : e];
            if (i > 0) load_direct(i - 1);                           // in flight during this step's arithmetic
        }
        if (alive) {
        const double* cfx = in + RW::FX;
        const double* cfu = in + RW::FU;
        // ---- W = V fx, Z = V fu   (V symmetric: VS(r,q))
        double W[N * N], Z[N * M];
#pragma unroll
        for (int c = 0; c < N; c++)
#pragma unroll
            for (int r = 0; r < N; r++) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < N; q++) acc = fma(Vs[tri(r, q)], cfx[q + N * c], acc);
                W[r + N * c] = acc;
            }
#pragma unroll
        for (int c = 0; c < M; c++)
#pragma unroll
            for (int r = 0; r < N; r++) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < N; q++) acc = fma(Vs[tri(r, q)], cfu[q + N * c], acc);
                Z[r + N * c] = acc;
            }
        // ---- Q expansion (backward_pass.jl:240-247); Qxx = cxx + fx'V fx is symmetric: upper triangle only
        double Qxx[NT], Qux[M * N], Quxr[M * N], Quu[M * M], QuuF[M * M], Qx[N], Qu[M];
        {
            const double* cxxi = cost_shared ? nullptr : tp(P.cxx, b, i);
#pragma unroll
            for (int c = 0; c < N; c++)
#pragma unroll
                for (int r = 0; r <= c; r++) {
                    double acc = 0.0;
#pragma unroll
                    for (int q = 0; q < N; q++) acc = fma(cfx[q + N * r], W[q + N * c], acc);
                    const double cs = cost_shared ? s_cost[tri(r, c)] : 0.5 * (cxxi[r + N * c] + cxxi[c + N * r]);
                    Qxx[tri(r, c)] = cs + acc;
                }
        }
        const double* cxui = cost_shared ? s_cost + NT : tp(P.cxu, b, i);
        const double* cuui = cost_shared ? s_cost + NT + N * M : tp(P.cuu, b, i);
#pragma unroll
        for (int j = 0; j < N; j++)
#pragma unroll
            for (int a = 0; a < M; a++) {
                double acc = 0.0, ff = 0.0;
#pragma unroll
                for (int q = 0; q < N; q++) { acc = fma(cfu[q + N * a], W[q + N * j], acc); ff = fma(cfu[q + N * a], cfx[q + N * j], ff); }
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
                for (int q = 0; q < N; q++) { acc = fma(cfu[q + N * a], Z[q + N * c], acc); ff = fma(cfu[q + N * a], cfu[q + N * c], ff); }
                const double v = cuui[a + M * c] + acc;
                Quu[a + M * c] = v;
                QuuF[a + M * c] = reg2 ? v + lam * ff : ((reg1 && a == c) ? v + lam : v);
            }
#pragma unroll
        for (int r = 0; r < N; r++) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < N; q++) acc = fma(cfx[q + N * r], Vx[q], acc);
            Qx[r] = in[RW::CX + r] + acc;
        }
#pragma unroll
        for (int a = 0; a < M; a++) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < N; q++) acc = fma(cfu[q + N * a], Vx[q], acc);
            Qu[a] = in[RW::CU + a] + acc;
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
            for (int a = 0; a < M; a++) { lo[a] = lims_lo[a] - in[RW::UU + a]; up[a] = lims_hi[a] - in[RW::UU + a]; }      // :45-46
            int nfac = 0;
            int res;
            if (M == 1) res = boxqp_scalar(QuuF[0], Qu[0], lo[0], up[0], kw[0], P.qp, &ki[0], &R[0], &fm, &nfac);     // same bits, scalar code
            else res = boxqp_seq<M>(M, QuuF, M, Qu, lo, up, kw, P.qp, ki, R, M, &fm, &nfac);
            if (res < 1) failed = true;                                                                  // :50-56
            nf = __popc(fm);
        }
        if (failed) { diverge = i + 1; alive = false; }
        else {
        if (M == 1) {
            // scalar control: K = -Qux_reg / QuuF (= R'R); one reciprocal that does not wait for the QP's sqrt/divide chain
            const double hinv = (nf > 0) ? -rcp_fast(QuuF[0]) : 0.0;       // QuuF > 0 here: its Cholesky factor exists
#pragma unroll
            for (int j = 0; j < N; j++) Ki[j] = Quxr[j] * hinv;
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
        // Vxx = Qxx + K'Quu K + K'Qux + Qux'K, symmetrised (:70-72): the upper triangle of the symmetric part, written
        // straight into Vs (V was last read by W and Z).  K'Qux + Qux'K is symmetric as it stands; K'Quu K is symmetrised
        // explicitly when Quu may be unsymmetric (M > 1).
#pragma unroll
        for (int c = 0; c < N; c++)
#pragma unroll
            for (int r = 0; r <= c; r++) {
                double t1 = 0.0, t1t = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
                for (int a = 0; a < M; a++) {
                    t1 = fma(Ki[a + M * r], QK[a + M * c], t1);
                    if (M > 1) t1t = fma(Ki[a + M * c], QK[a + M * r], t1t);
                    t2 = fma(Ki[a + M * r], Qux[a + M * c], t2);
                    t3 = fma(Qux[a + M * r], Ki[a + M * c], t3);
                }
                if (M > 1) t1 = 0.5 * (t1 + t1t);
                Vs[tri(r, c)] = ((Qxx[tri(r, c)] + t1) + t2) + t3;
            }
        // ---- store
        if (STAGE) {                                      // into the warp's output block; whole lines leave every OB steps (flush)
            const int sl = i & (RW::OB - 1);
#pragma unroll
            for (int e = 0; e < N * M; e += 2) *reinterpret_cast<double2*>(orow + RW::OK_ + sl * N * M + e) = make_double2(Ki[e], Ki[e + 1]);
#pragma unroll
            for (int r = 0; r < N; r += 2) *reinterpret_cast<double2*>(orow + RW::OVX + sl * N + r) = make_double2(Vx[r], Vx[r + 1]);
#pragma unroll
            for (int a = 0; a < M; a++) orow[RW::OKK + sl * M + a] = ki[a];
        } else if (vec_out && (N * M) % 2 == 0 && N % 2 == 0) {
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
        for (int a = 0; a < M; a++) { if (!STAGE) kb[(long long)i * M + a] = ki[a]; kw[a] = ki[a]; }
        if (Vxxb)
#pragma unroll
            for (int e = 0; e < N * N; e++) Vxxb[(long long)i * N * N + e] = Vs[tri(e % N, e / N)];
        if (Quub)
#pragma unroll
            for (int e = 0; e < M * M; e++) Quub[(long long)i * M * M + e] = Quu[e];
        }
        }
        if (STAGE && (i & (RW::OB - 1)) == 0) flush(i);            // block complete: steps i .. i+OB-1 leave as whole lines
    }
    if (STAGE) asm volatile("cp.async.wait_group 0;" ::: "memory");
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
        for (int e = 0; e < N * N; e++) P.Vxx1[b * N * N + e] = (diverge > 0) ? 0.0 : Vs[tri(e % N, e / N)];
    P.diverge[b] = diverge;
    P.dV[2 * b] = dV0;
    P.dV[2 * b + 1] = dV1;
}

#ifndef SMALL_MINB
#define SMALL_MINB 2
#endif
bool al16v(const TensorD& t) { return ((uintptr_t)t.p % 16 == 0) && (t.sb % 2 == 0) && (t.st % 2 == 0); }

template <int N, int M>
int launch_small(ddp_handle_s* h, const BackParams& P_in) {
    BackParams P = P_in;
    int rc = prepare_redo(h, P);                     // hand-over of trajectories with an unsymmetric terminal cxx to the generic kernel
    if (rc != 0) return rc;
    const unsigned grid = (unsigned)((P.B + 127) / 128);
    const bool u_ok = (P.lims == nullptr) || ((M % 2 == 0) ? al16v(P.u) : ((uintptr_t)P.u.p % 8 == 0));
    const bool cu_ok = (M % 2 == 0) ? al16v(P.cu) : true;
    const bool out_ok = ((uintptr_t)P.K % 16 == 0) && ((uintptr_t)P.Vx % 16 == 0);
    const bool stage = Row<N, M>::OK && al16v(P.fx) && al16v(P.fu) && al16v(P.cx) && cu_ok && u_ok && out_ok && !(getenv("DDP_SMALL_NOSTAGE"));
    // 2 CTAs (8 warps) per SM: 168- and 128-register builds (3 / 4 CTAs) spill
    const bool csh = (P.cxx.sb == 0 && P.cxx.st == 0 && P.cxu.sb == 0 && P.cxu.st == 0 && P.cuu.sb == 0 && P.cuu.st == 0);
#define LAUNCH_SMALL(MB, ST, CS, BYTES)                                                                                                   \
    do {                                                                                                                                  \
        cudaError_t ea = cudaFuncSetAttribute(bp_small_kernel<N, M, MB, ST, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES)); \
        if (ea != cudaSuccess) return (int)ea;                                                                                            \
        bp_small_kernel<N, M, MB, ST, CS><<<grid, 128, (BYTES), h->stream>>>(P);                                                          \
    } while (0)
    if (stage && Row<N, M>::OK) {
        const size_t bytes = sizeof(double) * 4 * Row<N, M>::WARP_DOUBLES;
        // residency: 2 CTAs (8 warps, <= 255 registers) or 3 CTAs (12 warps, 168 registers, a few spilled doubles) per SM;
        // DDP_SMALL_MINB selects for the A/B measurement; 2 is the faster one on B200 (12.8 vs 16.4 ms, profiles/README_r02.md)
        const char* mb = getenv("DDP_SMALL_MINB");
        const int minb = mb ? atoi(mb) : SMALL_MINB;
        if (minb == 3) { if (csh) LAUNCH_SMALL(3, (Row<N, M>::OK), true, bytes); else LAUNCH_SMALL(3, (Row<N, M>::OK), false, bytes); }
        else { if (csh) LAUNCH_SMALL(2, (Row<N, M>::OK), true, bytes); else LAUNCH_SMALL(2, (Row<N, M>::OK), false, bytes); }
    } else {
        if (csh) LAUNCH_SMALL(2, false, true, 16); else LAUNCH_SMALL(2, false, false, 16);
    }
#undef LAUNCH_SMALL
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    return launch_back_pass_generic(h, P, false);    // processes the handed-over trajectories; exits at once when there are none
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
