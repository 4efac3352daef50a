// Specialised backward sweep for n = 32, m = 8 (the headline shape): ONE WARP PER TRAJECTORY.
//
// Replaces back_pass (Cholesky branch) of src/backward_pass.jl:162-252 + :31-42, :64-76.
//
// The per-step dense algebra runs on FP64 tensor tiles (mma.sync.m8n8k4.f64, "DMMA"; measured
// 37.1 TFLOP/s on B200, and it shares the FP64 datapath with DFMA -- profiles/microbench).
// tcgen05 has no f64 kind, so this is the only matrix unit that keeps the 1e-8 FP64 contract.
// Everything between the two big products stays in registers, in DMMA fragment layouts:
//
//   F  = [fx fu]            32 x 40    shared memory (column-major, XOR-swizzled), loaded once (LTI)
//   W' = F' V               40 x 32    160 DMMA   V = Vxx(i+1), symmetric, shared memory
//   G  = W' F = F' V F      40 x 40    120 DMMA   upper 15 of 25 tiles; accumulators start from the
//                                                 cost terms; W' accumulators are re-used as the A
//                                                 operand by permuting the contraction index
//        G = [Qxx Qxu; . Quu].  In the accumulator layout a lane (g,q) holds Qux[2q..2q+1][8t+g]
//        and Quu[g][2q..2q+1] -- exactly the A/B fragments the next products need.
//   Minv = QuuF^-1          8 x 8      Gauss-Jordan in the accumulator layout, pivots broadcast by
//                                      shuffles; a pivot <= 0 is the reference's Cholesky failure
//   K'   = -Qux_reg' Minv   32 x 8     8 DMMA     (comes out as the fragment Vxx needs, and as
//                                                 coalesced 16-byte global stores)
//   M1'  = Qux' + K' Quu    32 x 8     8 DMMA
//   Vxx  = Qxx + K'M1 + Qux'K          40 DMMA onto the resident Qxx tiles, mirrored => exactly symmetric
//   k, Quu k, Vx, dV: a few FMAs per lane plus shuffles.
//
// 336 DMMA per step ~ 172 kflop, against the reference's 213 kflop formulation.
// Shared memory: 18.7 KB per warp (V, F, Vx); registers (255) limit residency to 8 warps per SM.
//
// Restrictions (anything else dispatches to the generic kernel): Cholesky branch only (lims ==
// NULL), 16-byte aligned fx/fu/cxx with even strides, symmetric cxx (it is a Hessian).
#include <algorithm>
#include <cstdlib>
#include "boxqp.cuh"

namespace {

constexpr int WPB_LTI = 4;                   // warps per CTA, time-invariant dynamics: 2 CTAs x 4 warps x 18.7 KB per SM
constexpr int WPB_LTV = 4;                   // time-varying: same residency; the single [fx fu] buffer is refilled for the next step right after its last read
constexpr int wpb(bool ltv) { return ltv ? WPB_LTV : WPB_LTI; }
constexpr int SV = 0;                        // Vxx            32 x 32 swizzled
constexpr int SF = SV + 1024;                // [fx fu]        32 x 40 swizzled
constexpr int SVX = SF + 1280;               // Vx (32)
constexpr int SCX = SVX + 32;                // this step's cx (32) and cu (8), landed by cp.async
constexpr int WARP_DOUBLES = SCX + 48;       // 2384 doubles = 19,072 B per warp
constexpr int WARP_DOUBLES_LTV = WARP_DOUBLES;  // (a second [fx fu] buffer would cost 10 KB per warp and halve the residency: measured 113-134 ms)
constexpr int COST_DOUBLES = 15 * 32 * 2;    // per-CTA table of the cost tiles in accumulator (fragment) order
// per-warp scratch of the box-QP branch (LIMS variants): H = QuuF (8 x 8), its reduced factor R, Qux_reg / K (8 x 32), vectors
constexpr int QH = 0, QR = 64, QQ = 128, QG = 384, QLO = 392, QUP = 400, QX0 = 408, QX = 416, QE = 424;     // QE: seven eight-double slots
constexpr int QP_DOUBLES = QE + 56;

// boxQP(QuuF, Qu, lims - u, k(i+1)) of backward_pass.jl:49 for m = 8 by the WHOLE WARP, in the oracle's arithmetic: every sum runs
// in the oracle's index order with separate multiply and add (boxqp.cuh), so result code, free set and every bit of k are what
// boxqp_seq<8> (generic kernel, oracle) produces -- but the O(m^2) sums are spread over the lanes and the code is kept small
// (the first versions were 80-140 KB of unrolled code and stalled on instruction fetch):
//   lane i (= lane & 7; the four groups of eight lanes compute the same thing) owns row i and column i of H, g_i, the bounds, x_i
//   and column i of the factor; vectors travel through eight-double slots of the per-warp scratch (one store, __syncwarp, four
//   LDS.128) and are then held by EVERY lane;
//   H x, x'H, H (x .* clamped)     one sequential 8-term sum per lane (row / column parallel)
//   x'g, (x'H) x                   products on the owning lanes, exchanged, then one sequential chain of adds on every lane
//   Cholesky of H[free,free]       row by row: element (r,i) on lane i subtracts R[p,r] R[p,i], p ascending; R[:,r] from the scratch
//   R'y = b, R z = y, search'grad  replicated: every lane runs the whole sequential solve on the factor in the scratch
// All lanes hold the same scalars, so control flow is warp-uniform.  Not inlined: the tile kernel keeps its registers.
// The factor lives in the scratch only (column j is written by lane j, row by row).
// Outputs: k -> sq[QX..], the factor (uncompacted: R[p][j] at sq[QR + p + 8 j], valid where p <= j are both free) and 1 / R[j][j] at
// sq[QG + j] (g is dead by then); returns the result code, *fm_out = free mask.
// The code is kept SMALL on purpose (loops over the rows of the factor are rolled, the value function is a call): the sweep loop and
// the QP must fit the instruction cache together -- the unrolled versions (41-81 KB) spent a third of their stalls on fetches.
__device__ __forceinline__ void qp_share8(double* sq, const int slot, const int i, const double v, double (&out)[8]) {
    sq[slot + i] = v;                                    // lanes with equal i write the same value
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const double2 t = *reinterpret_cast<const double2*>(&sq[slot + 2 * j]);
        out[2 * j] = t.x;
        out[2 * j + 1] = t.y;
    }
}

// x'g + ((0.5 x') H) x with the sums in index order (qp_value); leaves x (all eight entries) in slot QE+0
__device__ __noinline__ double qp_value_warp8(double* sq, const int i, const double xv, const double g) {
    double xs[8], pq[8];
    qp_share8(sq, QE + 0, i, xv, xs);
    double t = 0.0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const double2 h = *reinterpret_cast<const double2*>(&sq[QH + 2 * r + 8 * i]);      // H[2r..2r+1][i]
        t = DADD(t, DMUL(DMUL(0.5, xs[2 * r]), h.x));                                      // (0.5 x' H)_i
        t = DADD(t, DMUL(DMUL(0.5, xs[2 * r + 1]), h.y));
    }
    qp_share8(sq, QE + 8, i, DMUL(xv, g), pq);
    double s1 = 0.0;
#pragma unroll
    for (int j = 0; j < 8; j++) s1 = DADD(s1, pq[j]);
    qp_share8(sq, QE + 16, i, DMUL(t, xv), pq);
    double s2 = 0.0;
#pragma unroll
    for (int j = 0; j < 8; j++) s2 = DADD(s2, pq[j]);
    return DADD(s1, s2);
}

__device__ __noinline__ int boxqp_warp8(double* sq, QPOpts o, int lane, unsigned* fm_out) {
    const int i = lane & 7;
    double Hrow[8], xs[8];
#pragma unroll
    for (int j = 0; j < 8; j++) Hrow[j] = sq[QH + i + 8 * j];
    const double g = sq[QG + i], lower = sq[QLO + i], upper = sq[QUP + i];
    double x = clampd(sq[QX0 + i], lower, upper);                        // boxQP.jl:58
    double value = qp_value_warp8(sq, i, x, g);                          // :63
    double oldvalue = 0.0;
    unsigned clamped = 0u, free_mask = 0xffu;
    int result = 0, iter = 1;
    while (iter <= o.max_iter) {                                         // :71
        if (result != 0) break;
        if (iter > 1 && DSUB(oldvalue, value) < DMUL(o.min_rel_improve, fabs(oldvalue))) { result = 4; break; }     // :78
        oldvalue = value;
#pragma unroll
        for (int j = 0; j < 4; j++) {                                    // x, left in the slot by the value function
            const double2 t = *reinterpret_cast<const double2*>(&sq[QE + 2 * j]);
            xs[2 * j] = t.x;
            xs[2 * j + 1] = t.y;
        }
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++) s = DADD(s, DMUL(Hrow[j], xs[j]));
        const double grad = DADD(g, s);                                  // :85
        const unsigned old_clamped = clamped;
        clamped = __ballot_sync(0xffffffffu, (x == lower && grad > 0.0) || (x == upper && grad < 0.0)) & 0xffu;      // :92-94
        free_mask = 0xffu & ~clamped;
        if (clamped == 0xffu) { result = 6; break; }                     // :98
        if (iter == 1 || old_clamped != clamped) {                       // :104-117: upper factor of H[free,free], row by row
            bool fail = false;
#pragma unroll 1
            for (int r = 0; r < 8; r++) {
                if (!((free_mask >> r) & 1u)) continue;                  // warp-uniform
                double acc = sq[QH + r + 8 * i];                         // H[r][i] (upper triangle when r <= i)
#pragma unroll
                for (int p = 0; p < 7; p++)                              // - R[p][r] R[p][i], p < r ascending (rows p of both columns are final)
                    if (p < r && ((free_mask >> p) & 1u)) acc = DSUB(acc, DMUL(sq[QR + p + 8 * r], sq[QR + p + 8 * i]));
                const double d = __shfl_sync(0xffffffffu, acc, r);       // the pivot: lane r's element
                if (!(d > 0.0)) { fail = true; break; }
                const double rrr = __dsqrt_rn(d);
                sq[QR + r + 8 * i] = (i == r) ? rrr : DDIV(acc, rrr);    // meaningful on the free lanes i >= r; never read elsewhere
                __syncwarp();
            }
            if (fail) { *fm_out = free_mask; return -1; }                // PosDefException
        }
        double gv[8];
        qp_share8(sq, QE + 24, i, grad, gv);
        double gs = 0.0;                                                 // norm(grad[free]) :120
#pragma unroll
        for (int p = 0; p < 8; p++)
            if ((free_mask >> p) & 1u) gs = DADD(gs, DMUL(gv[p], gv[p]));
        if (__dsqrt_rn(gs) < o.min_grad) { result = 5; break; }
        double sc = 0.0;                                                 // grad_clamped = g + H (x .* clamped)  :127
#pragma unroll
        for (int j = 0; j < 8; j++) sc = DADD(sc, DMUL(Hrow[j], ((clamped >> j) & 1u) ? xs[j] : DMUL(xs[j], 0.0)));
        // R' y = b, then R z = y: the oracle's loops (chol_solve) on the uncompacted factor, by every lane (all lanes write the same
        // values); b, y and z have a slot each, so no element is read after it was overwritten
        double* vb = sq + QE + 32;
        double* vy = sq + QE + 40;
        double* vz = sq + QE + 48;
        vb[i] = DADD(g, sc);
        __syncwarp();
#pragma unroll 1
        for (int a = 0; a < 8; a++) {
            if (!((free_mask >> a) & 1u)) continue;
            double acc = vb[a];
#pragma unroll
            for (int p = 0; p < 7; p++)
                if (p < a && ((free_mask >> p) & 1u)) acc = DSUB(acc, DMUL(sq[QR + p + 8 * a], vy[p]));
            vy[a] = DDIV(acc, sq[QR + a + 8 * a]);
            __syncwarp();
        }
#pragma unroll 1
        for (int a = 7; a >= 0; a--) {
            if (!((free_mask >> a) & 1u)) continue;
            double acc = vy[a];
#pragma unroll
            for (int p = 1; p < 8; p++)
                if (p > a && ((free_mask >> p) & 1u)) acc = DSUB(acc, DMUL(sq[QR + a + 8 * p], vz[p]));
            vz[a] = DDIV(acc, sq[QR + a + 8 * a]);
            __syncwarp();
        }
        double search = 0.0, sdotg = 0.0;                                // :129, :132
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double sj = ((free_mask >> j) & 1u) ? DSUB(-vz[j], xs[j]) : 0.0;
            if (i == j) search = sj;
            sdotg = DADD(sdotg, DMUL(sj, gv[j]));
        }
        __syncwarp();                                                    // every lane is done with the solve slot
        if (sdotg >= 0.0) break;                                         // :133 leaves result == 0
        double step = 1.0;                                               // :138
        double xc = clampd(DADD(x, DMUL(step, search)), lower, upper);
        double vc = qp_value_warp8(sq, i, xc, g);
        while (DDIV(DSUB(vc, oldvalue), DMUL(step, sdotg)) < o.armijo) { // :142
            step = DMUL(step, o.step_dec);
            xc = clampd(DADD(x, DMUL(step, search)), lower, upper);
            vc = qp_value_warp8(sq, i, xc, g);
            if (step < o.min_step) { result = 2; break; }
        }
        x = xc;                                                          // :161
        value = vc;
        iter++;
    }
    if (iter == o.max_iter) result = 1;                                  // :167 (quirk Q4)
    __syncwarp();
    if (lane < 8) {
        sq[QX + i] = x;
        sq[QG + i] = 1.0 / sq[QR + i + 8 * i];                           // reciprocal pivots for the gain columns (garbage where clamped: unused)
    }
    __syncwarp();
    *fm_out = free_mask;
    return result;
}

// K[free, j] = -R \ (R' \ Qux_reg[free, j]) for column j = lane (backward_pass.jl:57-61), in place in sq[QQ..]; clamped rows = 0.
// Unrolled on registers; the factor and its reciprocal pivots are read from shared memory by broadcast.  (The gains carry the 1e-8
// contract, not the bit-exact one of the QP: products with 1/R[j][j] instead of divisions.)
__device__ __forceinline__ void qp8_gain_column(double* sq, int lane, unsigned fm) {
    double* col = sq + QQ + 8 * lane;
    double v[8];
#pragma unroll
    for (int a = 0; a < 8; a++) v[a] = col[a];
#pragma unroll
    for (int a = 0; a < 8; a++) {                                        // R' y = b
        if (!((fm >> a) & 1u)) continue;
        double acc = v[a];
#pragma unroll
        for (int p = 0; p < 8; p++)
            if (p < a && ((fm >> p) & 1u)) acc = fma(-sq[QR + p + 8 * a], v[p], acc);
        v[a] = acc * sq[QG + a];
    }
#pragma unroll
    for (int a = 7; a >= 0; a--) {                                       // R z = y
        if (!((fm >> a) & 1u)) continue;
        double acc = v[a];
#pragma unroll
        for (int p = 0; p < 8; p++)
            if (p > a && ((fm >> p) & 1u)) acc = fma(-sq[QR + a + 8 * p], v[p], acc);
        v[a] = acc * sq[QG + a];
    }
#pragma unroll
    for (int a = 0; a < 8; a++) col[a] = ((fm >> a) & 1u) ? -v[a] : 0.0;
}

// column-major 32-row matrices: element (i, c) lives at (i ^ s(c)) + 32 c with s(c) = ((c&1)<<3) | (((c>>1)&3)<<1).
// 16-byte row pairs stay together, DMMA fragment loads (LDS.128) are bank-conflict free.
__device__ __forceinline__ int swz(int i, int c) { return (i ^ (((c & 1) << 3) | (((c >> 1) & 3) << 1))) + 32 * c; }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }
__device__ __forceinline__ double shf(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

__device__ __forceinline__ void cp_async16(double* dst_smem, const double* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// stage [fx fu] of one step into the swizzled buffer (16-byte chunks = 2 consecutive rows of a column).
// Chunk c = lane + 32 k (k = 0..19) is column (lane >> 4) + 2k, rows i = 2 (lane & 15): per lane the row pair and the parity
// of the column are constants, so with k unrolled the swizzled destination is (i ^ s) + 32 col with s known up to a per-lane
// constant -- two integer instructions per chunk instead of a dozen.
template <bool ASYNC>
__device__ __forceinline__ void load_F(double* sF, const double* fx, const double* fu, int lane) {
    const int cb = lane >> 4, i = (lane & 15) << 1;
    const int i8 = i ^ (cb << 3);                       // (c & 1) << 3 with c & 1 == cb
    const double* sx = fx + cb * 32 + i;                // column cb + 2k of fx: + 64 k
    const double* su = fu + cb * 32 + i;                // column cb + 2 (k - 16) of fu
    double* d0 = sF + 32 * cb;
#pragma unroll
    for (int k = 0; k < 20; k++) {
        const double* src = (k < 16) ? (sx + 64 * k) : (su + 64 * (k - 16));
        double* dst = d0 + (i8 ^ ((k & 3) << 1)) + 64 * k;          // ((c >> 1) & 3) == (k & 3)
        if (ASYNC) cp_async16(dst, src);
        else st2(dst, src[0], src[1]);
    }
}

// 1/d with ONE Newton step on the hardware seed (relative error ~2^-44: enough for the 1e-8 contract of the gains; experimental)
__device__ __forceinline__ double rcp_nr1(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    return fma(y, e, y);
}

// 1/d for d > 0 in the normal range: hardware seed + two Newton steps
__device__ __forceinline__ double rcp_nr(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// In-place Gauss-Jordan inverse of an 8 x 8 matrix held in the DMMA accumulator layout (lane (g,q) holds
// A[g][2q], A[g][2q+1]); pivots and pivot rows/columns are broadcast by shuffles.  Returns false when a
// pivot is <= 0 (or NaN): for a symmetric matrix these are the Cholesky pivots squared, i.e. the
// reference's `cholesky` failure condition.
__device__ __forceinline__ bool gj_inverse8(double& I0, double& I1, int lane, int g, int q) {
    bool ok = true;
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const double own = (p & 1) ? I1 : I0;
        const double d = shf(own, 4 * p + (p >> 1));             // A[p][p]
        const double colp = shf(own, (lane & ~3) | (p >> 1));     // A[g][p]
        const double rp0 = shf(I0, 4 * p + q), rp1 = shf(I1, 4 * p + q);   // A[p][2q], A[p][2q+1]
        if (!(d > 0.0)) ok = false;
        const double r = rcp_nr(d);
        const double n0 = ((2 * q == p) ? 1.0 : rp0) * r, n1 = ((2 * q + 1 == p) ? 1.0 : rp1) * r;
        if (g == p) { I0 = n0; I1 = n1; }
        else {
            I0 = fma(-colp, n0, (2 * q == p) ? 0.0 : I0);
            I1 = fma(-colp, n1, (2 * q + 1 == p) ? 0.0 : I1);
        }
    }
    return ok;
}

constexpr int gidx(int at, int bt) { return at * 5 - (at * (at - 1)) / 2 + (bt - at); }   // upper-tile index, 15 tiles

// HIST: the optional Vxx histories (full and packed) are compiled in only when asked for, so the benchmarked variants
// carry neither their pointers nor their branches through the step loop
// LIMS: control limits given -> box-QP branch of @end_backward_pass (backward_pass.jl:43-62, :317-335) unless lims[1,1] > lims[1,2]
// EXP: experimental schedule of the Gauss-Jordan pivots (0: two per k-step of the fx'V block; 1: one per k-step of the fx'V block and one
//      per k-step of the W'F block, i.e. spread over 248 instead of 128 DMMAs; bit 1: one Newton step in the pivot reciprocal;
//      bit 3 (8): W' = F'V one row block at a time (16 instead of 64 registers of W); 4: the same with three CTAs per SM (168
//      registers, cost tiles read from global memory, no shared cost table)); selected by DDP_TILE_EXP for A/B measurements --
//      none beats the shipped schedule (profiles/tile_exp_r02.txt)
template <bool LTV, bool GPS, bool REG2, bool HIST, bool LIMS, int EXP = 0>
__global__ void __launch_bounds__(wpb(LTV) * 32, (EXP & 4) ? 3 : 2) bp_tile32x8_kernel(BackParams P) {
    constexpr int WPB = wpb(LTV);
    constexpr int WD = (LTV ? WARP_DOUBLES_LTV : WARP_DOUBLES) + (LIMS ? QP_DOUBLES : 0);
    extern __shared__ double smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    double* sm = smem_raw + (size_t)w * WD;
    double* sq = sm + (LTV ? WARP_DOUBLES_LTV : WARP_DOUBLES);          // box-QP scratch (LIMS)
    const bool use_qp = LIMS && !(P.lims[0] > P.lims[8]);               // backward_pass.jl:31
    const double lim_lo = LIMS ? P.lims[g] : 0.0, lim_hi = LIMS ? P.lims[8 + g] : 0.0;
    double* sV = sm + SV;
    double* sVx = sm + SVX;
    double* sCx = sm + SCX;
    const int N = P.T;
    const long long warps_total = (long long)gridDim.x * WPB;
    // Cost Hessians shared by the batch and constant in time (the reference's LTI/QTIC methods): stage the
    // symmetrised tiles once per CTA in accumulator order, so each step starts G with 15 conflict-free LDS.128.
    double* sCost = smem_raw + (size_t)WPB * WD;
    const bool cost_shared = !(EXP & 4) && (P.cxx.sb == 0 && P.cxx.st == 0 && P.cxu.sb == 0 && P.cxu.st == 0 && P.cuu.sb == 0 && P.cuu.st == 0);
    if (cost_shared) {
        if (w == 0) {
            const double* cxx0 = P.cxx.p;
            const double* cxu0 = P.cxu.p;
            const double* cuu0 = P.cuu.p;
#pragma unroll
            for (int at = 0; at < 4; at++) {
                const int a = 8 * at + g;
#pragma unroll
                for (int bt = at; bt < 4; bt++) {
                    const int b0 = 8 * bt + 2 * q;
                    st2(&sCost[(gidx(at, bt) * 32 + lane) * 2], 0.5 * (cxx0[a * 32 + b0] + cxx0[b0 * 32 + a]),
                        0.5 * (cxx0[a * 32 + b0 + 1] + cxx0[(b0 + 1) * 32 + a]));
                }
                st2(&sCost[(gidx(at, 4) * 32 + lane) * 2], cxu0[a + 32 * (2 * q)], cxu0[a + 32 * (2 * q + 1)]);
            }
            st2(&sCost[(gidx(4, 4) * 32 + lane) * 2], cuu0[g + 8 * (2 * q)], cuu0[g + 8 * (2 * q + 1)]);
        }
        __syncthreads();
    }
    // lane constants of the swizzled addressing: swz(8p + 2q, 8t + g) = (p even ? LAe : LAo) + 8p + 256t,
    // swz(8at + g, 8bt + 2q + h) = LM + 8(at ^ h) + 32h + 256bt
    const int gg = (g >> 1) & 3, par = g & 1;
    const int LA = 2 * (q ^ gg) + 32 * g;
    const int LAe = LA + 8 * par, LAo = LA - 8 * par;
    const int LM = (g ^ (2 * q)) + 64 * q;
#define FRAG(p, t) (((p) & 1 ? LAo : LAe) + 8 * (p) + 256 * (t))
#define MIRR(at, bt, h) (LM + 8 * ((at) ^ (h)) + 32 * (h) + 256 * (bt))
    const bool up0 = (2 * q >= g), st0 = (2 * q > g), up1 = (2 * q + 1 >= g), st1 = (2 * q + 1 > g);
    const int c0src = 8 * q, c1src = 8 * q + 4;        // a lane of the group that owns entry 2q / 2q+1

    for (long long b = (long long)blockIdx.x * WPB + w; b < P.B; b += warps_total) {
        if (P.active && !P.active[b]) continue;
        const double lam = GPS ? 0.0 : P.lambda[b];
        constexpr bool reg2 = REG2 && !GPS;
        const double eta = GPS ? P.eta[b] : 1.0, ieta = 1.0 / eta;
        double* Quuib = (GPS && P.Quui) ? P.Quui + b * (long long)N * 64 : nullptr;
        const bool has_kp = GPS && (P.kp.p != nullptr);
        double* Kb = P.K + b * (long long)N * 256;
        double* kb = P.k + b * (long long)N * 8;
        double* Vxb = P.Vx + b * (long long)N * 32;
        double* Vxxb = (HIST && P.Vxx) ? P.Vxx + b * (long long)N * 1024 : nullptr;
        double* Quub = P.Quu ? P.Quu + b * (long long)N * 64 : nullptr;
        // packed upper triangle of Vxx (528 per step), optional: the pointer is re-formed from the parameter block at each use
        // so that this rarely used output costs the step loop no register
        auto dump_tri = [&](long long step) {                                          // column c: rows 0..c are contiguous
            double* o = P.Vxx_tri + (b * (long long)N + step) * 528;
#pragma unroll 4
            for (int c = 0; c < 32; c++)
                if (lane <= c) o[c * (c + 1) / 2 + lane] = sV[swz(lane, c)];
        };
        const double* cxb = P.cx.p + b * P.cx.sb;
        const double* cub = P.cu.p + b * P.cu.sb;
        __syncwarp();
        // ---- terminal step
        {
            const double* cxN = cxb + (long long)(N - 1) * P.cx.st;
            const double* cxxN = tp(P.cxx, b, N - 1);
            const double* cuuN = tp(P.cuu, b, N - 1);
            double v = cxN[lane];
            sVx[lane] = v;
            Vxb[(long long)(N - 1) * 32 + lane] = v;
            for (int c = lane; c < 512; c += 32) {
                int col = c >> 4, i = (c & 15) << 1;
                double2 t = ld2(cxxN + col * 32 + i);
                st2(&sV[swz(i, col)], t.x, t.y);
                if (Vxxb) st2(Vxxb + (long long)(N - 1) * 1024 + col * 32 + i, t.x, t.y);
            }
            // the products below assume Vxx = Vxx': an inexactly symmetric (or NaN) terminal cxx goes to the generic kernel
            __syncwarp();
            {
                bool asym = false;
                for (int c = lane; c < 512; c += 32) {
                    const int col = c >> 4, i = (c & 15) << 1;
                    const double2 t = ld2(&sV[swz(i, col)]);
                    if (t.x != sV[swz(col, i)] || t.y != sV[swz(col, i + 1)]) asym = true;
                }
                asym = __any_sync(0xffffffffu, asym);
                if (lane == 0) {
                    P.redo[b] = asym ? 1 : 0;
                    if (asym) atomicAdd(P.redo_count, 1);
                }
                if (asym) continue;
            }
            if (HIST && P.Vxx_tri) dump_tri(N - 1);
            for (int c = lane; c < 128; c += 32) st2(Kb + (long long)(N - 1) * 256 + 2 * c, 0.0, 0.0);
            if (lane < 8) kb[(long long)(N - 1) * 8 + lane] = 0.0;
            if (GPS) {                              // Quu(N) = cuu/eta + Sigma_i_prev(N), Sigma(N) = inv  (backward_pass.jl:282-283)
                const double* SiN = tp(P.Sip, b, N - 1);
                double T0 = cuuN[g + 8 * (2 * q)] / eta + SiN[g + 8 * (2 * q)];
                double T1 = cuuN[g + 8 * (2 * q + 1)] / eta + SiN[g + 8 * (2 * q + 1)];
                if (Quub) { Quub[(long long)(N - 1) * 64 + g + 8 * (2 * q)] = T0; Quub[(long long)(N - 1) * 64 + g + 8 * (2 * q + 1)] = T1; }
                gj_inverse8(T0, T1, lane, g, q);
                if (Quuib) { Quuib[(long long)(N - 1) * 64 + g + 8 * (2 * q)] = T0; Quuib[(long long)(N - 1) * 64 + g + 8 * (2 * q + 1)] = T1; }
            } else if (Quub) st2(Quub + (long long)(N - 1) * 64 + 2 * lane, cuuN[2 * lane], cuuN[2 * lane + 1]);
        }
        if (LIMS && lane < 8) sq[QX0 + lane] = 0.0;          // warm start of the first QP: k[:, N-1] = 0 (backward_pass.jl:49)
        if (LTV) {
            if (N >= 2) load_F<true>(sm + SF, tp(P.fx, b, N - 2), tp(P.fu, b, N - 2), lane);
        } else {
            load_F<false>(sm + SF, tp(P.fx, b, 0), tp(P.fu, b, 0), lane);
        }
        __syncwarp();
        // regType 2 needs F'fu (40 x 8): constant over time for LTI
        double FF[5][2];
#pragma unroll
        for (int t = 0; t < 5; t++) FF[t][0] = FF[t][1] = 0.0;
        auto compute_FF = [&](const double* sF) {
#pragma unroll
            for (int t = 0; t < 5; t++) FF[t][0] = FF[t][1] = 0.0;
#pragma unroll
            for (int p = 0; p < 4; p++) {
                double2 fb = ld2(&sF[FRAG(p, 4)]);
#pragma unroll
                for (int t = 0; t < 5; t++) {
                    double2 fa = ld2(&sF[FRAG(p, t)]);
                    dmma(FF[t][0], FF[t][1], fa.x, fb.x);
                    dmma(FF[t][0], FF[t][1], fa.y, fb.y);
                }
            }
        };
        if (!LTV && reg2) compute_FF(sm + SF);

        double dV0 = 0.0, dV1 = 0.0;
        int diverge = 0;
        for (int i = N - 2; i >= 0; i--) {
            const double* sF = sm + SF;
            if (LTV) {
                cp_async_wait_all();
                __syncwarp();
                // the next step's [fx fu] (10 KB = 80 lines) is pulled into L2 now, a whole tensor phase before the cp.async refill
                // below asks for it: the refill has only the step's non-tensor tail to land in, too short for a DRAM round trip
                if (i > 0) {
                    const char* pfx = reinterpret_cast<const char*>(tp(P.fx, b, i - 1));
                    const char* pfu = reinterpret_cast<const char*>(tp(P.fu, b, i - 1));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pfx + 128 * lane));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pfx + 128 * (lane + 32)));
                    if (lane < 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(pfu + 128 * lane));
                }
            }
            // this step's cost gradients -> shared memory (group A; no registers are held across the tensor phase)
            if (lane < 16) cp_async16(&sCx[2 * lane], cxb + (long long)i * P.cx.st + 2 * lane);
            else if (lane < 20) cp_async16(&sCx[2 * lane], cub + (long long)i * P.cu.st + 2 * (lane - 16));
            cp_async_commit();
            if (LTV && reg2) compute_FF(sF);
            // this step's cost gradients: lanes of group g own Qx[8t+g] (t = 0..3) and Qu[g]
            // KL terms: fragments of the previous policy, K_prev[2q..2q+1][8t+g], Sigma_i_prev[g][2q..2q+1]
            double2 kpf[4];
            double Sg0 = 0.0, Sg1 = 0.0, kp0 = 0.0, kp1 = 0.0;
            if (GPS) {
                const double* Kpi = tp(P.Kp, b, i) + 8 * g + 2 * q;
#pragma unroll
                for (int t = 0; t < 4; t++) kpf[t] = ld2(Kpi + 64 * t);
                const double* Sii = tp(P.Sip, b, i);
                Sg0 = Sii[g + 8 * (2 * q)];
                Sg1 = Sii[g + 8 * (2 * q + 1)];
                if (has_kp) { const double* kpi = tp(P.kp, b, i); kp0 = kpi[2 * q]; kp1 = kpi[2 * q + 1]; }
            }
            // dump Vxx(i+1) history if requested (sV is stable here)
            if (Vxxb && i < N - 2) {
                for (int c = lane; c < 512; c += 32) {
                    int col = c >> 4, r = (c & 15) << 1;
                    double2 t = ld2(&sV[swz(r, col)]);
                    st2(Vxxb + (long long)(i + 1) * 1024 + col * 32 + r, t.x, t.y);
                }
            }
            if (HIST && P.Vxx_tri && i < N - 2) dump_tri(i + 1);
            // ---- G starts from the cost terms: their global loads are in flight during the tensor phase
            double G[15][2];
            if (cost_shared) {
#pragma unroll
                for (int t = 0; t < 15; t++) {
                    const double2 c = ld2(&sCost[(t * 32 + lane) * 2]);
                    G[t][0] = c.x;
                    G[t][1] = c.y;
                }
            } else {
                const double* cxxi = tp(P.cxx, b, i);
                const double* cxui = tp(P.cxu, b, i);
                const double* cuui = tp(P.cuu, b, i);
#pragma unroll
                for (int at = 0; at < 4; at++) {
                    const int a = 8 * at + g;
#pragma unroll
                    for (int bt = at; bt < 4; bt++) {
                        const int b0 = 8 * bt + 2 * q;
                        double2 lo = ld2(cxxi + a * 32 + b0);                    // cxx[b0..b0+1][a]
                        double u0 = cxxi[b0 * 32 + a], u1 = cxxi[(b0 + 1) * 32 + a];   // cxx[a][b0], cxx[a][b0+1]
                        G[gidx(at, bt)][0] = 0.5 * (lo.x + u0);
                        G[gidx(at, bt)][1] = 0.5 * (lo.y + u1);
                    }
                    G[gidx(at, 4)][0] = cxui[a + 32 * (2 * q)];                  // Qux[b'][a] = cxu[a][b'] + ...
                    G[gidx(at, 4)][1] = cxui[a + 32 * (2 * q + 1)];
                }
                G[gidx(4, 4)][0] = cuui[g + 8 * (2 * q)];
                G[gidx(4, 4)][1] = cuui[g + 8 * (2 * q + 1)];
            }
            // ---- tensor phase, by row blocks of W' = F'V so that only 1-2 row blocks (not all 5) are live:
            //      block 4 (fu'V) first => Quu = G(4,4) is complete after 40 DMMAs and its Gauss-Jordan inverse (a long
            //      latency-bound chain of shuffles and FMAs) is interleaved, pivot by pivot, with the DMMAs of blocks 0,1.
            double fv[5];
            double I0 = 0.0, I1 = 0.0, U0, U1;
            bool ok = true;
            auto gj_step = [&](const int p) {                 // one pivot of the in-place inverse (see gj_inverse8)
                const double own = (p & 1) ? I1 : I0;
                const double d = shf(own, 4 * p + (p >> 1));
                const double colp = shf(own, (lane & ~3) | (p >> 1));
                const double rp0 = shf(I0, 4 * p + q), rp1 = shf(I1, 4 * p + q);
                if (!(d > 0.0)) ok = false;
                const double r = (EXP & 2) ? rcp_nr1(d) : rcp_nr(d);
                const double n0 = ((2 * q == p) ? 1.0 : rp0) * r, n1 = ((2 * q + 1 == p) ? 1.0 : rp1) * r;
                if (g == p) { I0 = n0; I1 = n1; }              // (forming colp * rp before the reciprocal arrives was measured slower: 94.3 vs 92.9 ms)
                else {
                    I0 = fma(-colp, n0, (2 * q == p) ? 0.0 : I0);
                    I1 = fma(-colp, n1, (2 * q + 1 == p) ? 0.0 : I1);
                }
            };
            auto w_block1 = [&](const int a0, double (&Wa)[4][2], double& fva, const int piv = -1) {
#pragma unroll
                for (int jt = 0; jt < 4; jt++) Wa[jt][0] = Wa[jt][1] = 0.0;
                fva = 0.0;
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const double2 fa0 = ld2(&sF[FRAG(p, a0)]);
                    double2 fb[4];
#pragma unroll
                    for (int jt = 0; jt < 4; jt++) fb[jt] = ld2(&sV[FRAG(p, jt)]);
                    const double2 vx = ld2(&sVx[8 * p + 2 * q]);
#pragma unroll
                    for (int jt = 0; jt < 4; jt++) dmma(Wa[jt][0], Wa[jt][1], fa0.x, fb[jt].x);
                    if (piv >= 0 && p == 1) gj_step(piv);
                    fva = fma(fa0.y, vx.y, fma(fa0.x, vx.x, fva));
#pragma unroll
                    for (int jt = 0; jt < 4; jt++) dmma(Wa[jt][0], Wa[jt][1], fa0.y, fb[jt].y);
                }
            };
            auto w_block0123 = [&](double (&W)[4][4][2]) {      // rows 0..3 of W' (fx'V), the 8 Gauss-Jordan pivots in between
#pragma unroll
                for (int at = 0; at < 4; at++) {
                    fv[at] = 0.0;
#pragma unroll
                    for (int jt = 0; jt < 4; jt++) W[at][jt][0] = W[at][jt][1] = 0.0;
                }
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    double2 fa[4], fb[4];
#pragma unroll
                    for (int at = 0; at < 4; at++) fa[at] = ld2(&sF[FRAG(p, at)]);
#pragma unroll
                    for (int jt = 0; jt < 4; jt++) fb[jt] = ld2(&sV[FRAG(p, jt)]);
                    const double2 vx = ld2(&sVx[8 * p + 2 * q]);
#pragma unroll
                    for (int at = 0; at < 4; at++)
#pragma unroll
                        for (int jt = 0; jt < 4; jt++) dmma(W[at][jt][0], W[at][jt][1], fa[at].x, fb[jt].x);
                    if (!LIMS) gj_step((EXP & 1) ? p : 2 * p);
#pragma unroll
                    for (int at = 0; at < 4; at++) fv[at] = fma(fa[at].y, vx.y, fma(fa[at].x, vx.x, fv[at]));
#pragma unroll
                    for (int at = 0; at < 4; at++)
#pragma unroll
                        for (int jt = 0; jt < 4; jt++) dmma(W[at][jt][0], W[at][jt][1], fa[at].y, fb[jt].y);
                    if (!LIMS && !(EXP & 1)) gj_step(2 * p + 1);
                }
            };
            // G(a, a..4) += W'[a,:] F[:, a..4]
            auto g_block1 = [&](const int a0, double (&Wa)[4][2], const int piv = -1) {
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    double2 ff[5];
#pragma unroll
                    for (int bt = a0; bt < 5; bt++) ff[bt] = ld2(&sF[FRAG(p, bt)]);
#pragma unroll
                    for (int bt = a0; bt < 5; bt++) dmma(G[gidx(a0, bt)][0], G[gidx(a0, bt)][1], Wa[p][0], ff[bt].x);
                    if (piv >= 0 && p == 1) gj_step(piv);
#pragma unroll
                    for (int bt = a0; bt < 5; bt++) dmma(G[gidx(a0, bt)][0], G[gidx(a0, bt)][1], Wa[p][1], ff[bt].y);
                }
            };
            auto g_block0123 = [&](double (&W)[4][4][2]) {
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    double2 ff[5];
#pragma unroll
                    for (int bt = 0; bt < 5; bt++) ff[bt] = ld2(&sF[FRAG(p, bt)]);
#pragma unroll
                    for (int at = 0; at < 4; at++)
#pragma unroll
                        for (int bt = at; bt < 5; bt++) dmma(G[gidx(at, bt)][0], G[gidx(at, bt)][1], W[at][p][0], ff[bt].x);
                    if (!LIMS && (EXP & 1)) gj_step(4 + p);
#pragma unroll
                    for (int at = 0; at < 4; at++)
#pragma unroll
                        for (int bt = at; bt < 5; bt++) dmma(G[gidx(at, bt)][0], G[gidx(at, bt)][1], W[at][p][1], ff[bt].y);
                }
            };
            {
                double Wa[4][2];
                // block 4: fu'V, Quu
                w_block1(4, Wa, fv[4]);
                g_block1(4, Wa);
                // (U0,U1) = Quu[g][2q..2q+1]; QuuF in (I0,I1)
                U0 = G[gidx(4, 4)][0];
                U1 = G[gidx(4, 4)][1];
                if (GPS) {                                    // Quu/eta + Sigma_i, then 0.5 (Quu + Quu')  (:297, :301)
                    U0 = fma(U0, ieta, Sg0);
                    U1 = fma(U1, ieta, Sg1);
                    const int s0 = 4 * (2 * q) + (g >> 1), s1 = 4 * (2 * q + 1) + (g >> 1);
                    const double a00 = shf(U0, s0), a01 = shf(U1, s0), a10 = shf(U0, s1), a11 = shf(U1, s1);
                    U0 = 0.5 * (U0 + ((g & 1) ? a01 : a00));  // Quu[2q][g]
                    U1 = 0.5 * (U1 + ((g & 1) ? a11 : a10));  // Quu[2q+1][g]
                }
                if (reg2) { I0 = fma(lam, FF[4][0], U0); I1 = fma(lam, FF[4][1], U1); }
                else { I0 = U0 + ((g == 2 * q) ? lam : 0.0); I1 = U1 + ((g == 2 * q + 1) ? lam : 0.0); }
            }
            if (EXP & 12) {
                // one row block of W' at a time (16 instead of 64 registers of W: the three-CTAs-per-SM experiment)
#pragma unroll
                for (int at = 0; at < 4; at++) {
                    double Wa[4][2];
                    w_block1(at, Wa, fv[at], LIMS ? -1 : 2 * at);
                    g_block1(at, Wa, LIMS ? -1 : 2 * at + 1);
                }
            } else {
                // blocks 0..3 with the Gauss-Jordan pivots in between; pivot p <= 0  <=>  Cholesky fails
                double W[4][4][2];
                w_block0123(W);
                g_block0123(W);
            }
            if (LTV) {                                       // [fx fu] was read for the last time in this step: refill it for the
                __syncwarp();                                // next one (group B); it lands during the non-tensor tail of the step
                if (i > 0) load_F<true>(sm + SF, tp(P.fx, b, i - 1), tp(P.fu, b, i - 1), lane);
                cp_async_commit();
            }
#pragma unroll
            for (int at = 0; at < 5; at++) {
                fv[at] += shx(fv[at], 1);
                fv[at] += shx(fv[at], 2);
            }
            // this step's cost gradients have landed long ago: lanes of group g own Qx[8t+g] (t = 0..3) and Qu[g]
            if (LTV) cp_async_wait_group<1>(); else cp_async_wait_group<0>();      // group A has landed
            __syncwarp();
            double cxv[4];
#pragma unroll
            for (int t = 0; t < 4; t++) cxv[t] = sCx[8 * t + g];
            const double cuv = sCx[32 + g];
            // ---- KL augmentation (backward_pass.jl:295-301): Q/eta + KL terms, formed in fragment layouts
            double2 sf[4];
            double Sik_own = 0.0, Sik0 = 0.0, Sik1 = 0.0;
            if (GPS) {
#pragma unroll
                for (int t = 0; t < 14; t++) { G[t][0] *= ieta; G[t][1] *= ieta; }      // tile (4,4) was scaled into (U0,U1) above
#pragma unroll
                for (int t = 0; t < 4; t++) sf[t] = make_double2(0.0, 0.0);        // S = Sigma_i K_prev : sf[t] = S[2q..2q+1][8t+g]
#pragma unroll
                for (int t = 0; t < 4; t++) dmma(sf[t].x, sf[t].y, kpf[t].x, Sg0);
#pragma unroll
                for (int t = 0; t < 4; t++) dmma(sf[t].x, sf[t].y, kpf[t].y, Sg1);
#pragma unroll
                for (int at = 0; at < 4; at++)                                     // Qxx += K_prev' S
#pragma unroll
                    for (int bt = at; bt < 4; bt++) dmma(G[gidx(at, bt)][0], G[gidx(at, bt)][1], kpf[at].x, sf[bt].x);
#pragma unroll
                for (int at = 0; at < 4; at++)
#pragma unroll
                    for (int bt = at; bt < 4; bt++) dmma(G[gidx(at, bt)][0], G[gidx(at, bt)][1], kpf[at].y, sf[bt].y);
                if (has_kp) {                                                      // Sigma_i k_prev, entry g per group
                    double t = fma(Sg1, kp1, Sg0 * kp0);
                    t += shx(t, 1);
                    t += shx(t, 2);
                    Sik_own = t;
                    Sik0 = shf(Sik_own, c0src);
                    Sik1 = shf(Sik_own, c1src);
                }
            }
            // ---- fragments straight from the accumulators:
            //      qf[t]  = Qux[2q..2q+1][8t+g]   (unregularised), qr[t] = Qux_reg
            double2 qf[4], qr[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                qf[t] = make_double2(G[gidx(t, 4)][0], G[gidx(t, 4)][1]);
                if (GPS) { qf[t].x -= sf[t].x; qf[t].y -= sf[t].y; }              // cxukl = -Sigma_i K_prev
                qr[t] = reg2 ? make_double2(fma(lam, FF[t][0], qf[t].x), fma(lam, FF[t][1], qf[t].y)) : qf[t];
            }
            const double Qu_own = GPS ? fma(cuv + fv[4], ieta, -Sik_own) : (cuv + fv[4]);     // Qu/eta + cukl, cukl = -Sigma_i k_prev
            const double Qu0 = shf(Qu_own, c0src), Qu1 = shf(Qu_own, c1src);
            double2 kf[4];
            double k_own;
            double J0 = U0, J1 = U1;                          // GPS + LIMS: Sigma = inv(Quu) is still wanted (backward_pass.jl:346)
            if (LIMS && use_qp) {
                // ---- box-QP branch (backward_pass.jl:43-62): QP by the warp in the oracle's arithmetic order, gains by columns
                sq[QH + g + 8 * (2 * q)] = I0;               // H = QuuF, column-major (unsymmetrised, as the reference passes it)
                sq[QH + g + 8 * (2 * q + 1)] = I1;
                if (q == 0) {
                    const double ui = tp(P.u, b, i)[g];
                    sq[QG + g] = Qu_own;
                    sq[QLO + g] = lim_lo - ui;                // lower = lims[:,1] - u[:,i]   (:45-46)
                    sq[QUP + g] = lim_hi - ui;
                }
#pragma unroll
                for (int t = 0; t < 4; t++) st2(&sq[QQ + 2 * q + 8 * (8 * t + g)], qr[t].x, qr[t].y);     // Qux_reg (8 x 32), column-major
                __syncwarp();
                unsigned fm = 0u;
                const int res = boxqp_warp8(sq, P.qp, lane, &fm);
                if (res < 1) ok = false;                      // :50-56
                if (ok) {
                    qp8_gain_column(sq, lane, fm);
                    __syncwarp();
#pragma unroll
                    for (int t = 0; t < 4; t++) kf[t] = ld2(&sq[QQ + 2 * q + 8 * (8 * t + g)]);
                    k_own = sq[QX + g];
                    __syncwarp();
                    if (lane < 8) sq[QX0 + lane] = sq[QX + lane];      // next step's warm start
                    if (GPS) gj_inverse8(J0, J1, lane, g, q);
                } else {
#pragma unroll
                    for (int t = 0; t < 4; t++) kf[t] = make_double2(0.0, 0.0);
                    k_own = 0.0;
                }
            } else {
            if (LIMS) ok = gj_inverse8(I0, I1, lane, g, q);  // limits given but inverted (lims[1,1] > lims[1,2]): Cholesky branch, not interleaved
            // ---- K' = -Qux_reg' Minv : kf[t] = K[2q..2q+1][8t+g]
#pragma unroll
            for (int t = 0; t < 4; t++) kf[t] = make_double2(0.0, 0.0);
#pragma unroll
            for (int t = 0; t < 4; t++) dmma(kf[t].x, kf[t].y, qr[t].x, I0);
#pragma unroll
            for (int t = 0; t < 4; t++) dmma(kf[t].x, kf[t].y, qr[t].y, I1);
#pragma unroll
            for (int t = 0; t < 4; t++) { kf[t].x = -kf[t].x; kf[t].y = -kf[t].y; }
            // ---- k = -Minv Qu, Quu k   (group g owns entry g; entries 2q, 2q+1 are fetched by shuffle)
            double ks = fma(I1, Qu1, I0 * Qu0);
            ks += shx(ks, 1);
            ks += shx(ks, 2);
            k_own = -ks;
            J0 = I0; J1 = I1;
            }
            if (!ok) { diverge = i + 1; break; }
            const double k0 = shf(k_own, c0src), k1 = shf(k_own, c1src);
            double qs = fma(U1, k1, U0 * k0);
            qs += shx(qs, 1);
            qs += shx(qs, 2);
            const double Quuk_own = qs;
            const double z0 = shf(Quuk_own, c0src) + Qu0, z1 = shf(Quuk_own, c1src) + Qu1;   // (Quu k + Qu)[2q..2q+1]
            // ---- M1' = Qux' + K' Quu : mf[t] = M1[2q..2q+1][8t+g]
            double2 mf[4];
#pragma unroll
            for (int t = 0; t < 4; t++) mf[t] = qf[t];
#pragma unroll
            for (int t = 0; t < 4; t++) dmma(mf[t].x, mf[t].y, kf[t].x, U0);
#pragma unroll
            for (int t = 0; t < 4; t++) dmma(mf[t].x, mf[t].y, kf[t].y, U1);
            // ---- Vx(i) = Qx + K'(Quu k + Qu) + Qux'k ; dV   (backward_pass.jl:64-69)
            double vxn[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                double sacc = fma(kf[t].y, z1, kf[t].x * z0);
                sacc = fma(qf[t].x, k0, sacc);
                sacc = fma(qf[t].y, k1, sacc);
                sacc += shx(sacc, 1);
                sacc += shx(sacc, 2);
                double qxv = cxv[t] + fv[t];
                if (GPS) {                                    // Qx/eta + cxkl, cxkl = K_prev' Sigma_i k_prev
                    qxv *= ieta;
                    if (has_kp) {
                        double c = fma(kpf[t].y, Sik1, kpf[t].x * Sik0);
                        c += shx(c, 1);
                        c += shx(c, 2);
                        qxv += c;
                    }
                }
                vxn[t] = qxv + sacc;
            }
            {
                // one term per group g (the four lanes of a group hold the same one); the sum over the groups is taken once,
                // after the sweep: the per-step butterfly was a dependent shuffle chain in every tail
                dV0 = fma(k_own, Qu_own, dV0);
                dV1 = fma(0.5 * k_own, Quuk_own, dV1);
            }
            // ---- outputs of this step
            {
                double* Kg = Kb + (long long)i * 256 + 8 * g + 2 * q;       // K[2q..2q+1][8t+g]: 512 contiguous bytes per tile
#pragma unroll
                for (int t = 0; t < 4; t++) st2(Kg + 64 * t, kf[t].x, kf[t].y);
                if (q == 0) {
                    kb[(long long)i * 8 + g] = k_own;
#pragma unroll
                    for (int t = 0; t < 4; t++) Vxb[(long long)i * 32 + 8 * t + g] = vxn[t];
                }
                if (Quub) {
                    Quub[(long long)i * 64 + g + 8 * (2 * q)] = U0;
                    Quub[(long long)i * 64 + g + 8 * (2 * q + 1)] = U1;
                }
                if (Quuib) {                                  // Sigma = inv(Quu)  (backward_pass.jl:346)
                    Quuib[(long long)i * 64 + g + 8 * (2 * q)] = J0;
                    Quuib[(long long)i * 64 + g + 8 * (2 * q + 1)] = J1;
                }
            }
            // ---- step 4: Vxx = Qxx + K' M1 + Qux' K  (upper 10 tiles), all operands in registers
#pragma unroll
            for (int pass = 0; pass < 4; pass++)             // 10 independent tiles between revisits
#pragma unroll
                for (int at = 0; at < 4; at++)
#pragma unroll
                    for (int bt = at; bt < 4; bt++) {
                        double& c0 = G[gidx(at, bt)][0];
                        double& c1 = G[gidx(at, bt)][1];
                        if (pass == 0) dmma(c0, c1, kf[at].x, mf[bt].x);
                        if (pass == 1) dmma(c0, c1, kf[at].y, mf[bt].y);
                        if (pass == 2) dmma(c0, c1, qf[at].x, kf[bt].x);
                        if (pass == 3) dmma(c0, c1, qf[at].y, kf[bt].y);
                    }
            __syncwarp();          // every lane is past its reads of sV / sVx for this step
#pragma unroll
            for (int at = 0; at < 4; at++) {
#pragma unroll
                for (int bt = at; bt < 4; bt++) {
                    const double v0 = G[gidx(at, bt)][0], v1 = G[gidx(at, bt)][1];
                    if (bt > at) {
                        st2(&sV[FRAG(bt, at)], v0, v1);                  // V[8bt+2q..+1][8at+g]
                        sV[MIRR(at, bt, 0)] = v0;                        // V[8at+g][8bt+2q]
                        sV[MIRR(at, bt, 1)] = v1;
                    } else {                                             // diagonal tile: keep the upper part, mirror it
                        if (up0) sV[FRAG(at, at)] = v0;
                        if (st0) sV[MIRR(at, at, 0)] = v0;
                        if (up1) sV[FRAG(at, at) + 1] = v1;
                        if (st1) sV[MIRR(at, at, 1)] = v1;
                    }
                }
            }
            if (q == 0) {
#pragma unroll
                for (int t = 0; t < 4; t++) sVx[8 * t + g] = vxn[t];
            }
            __syncwarp();
        }
        if (LTV) { cp_async_wait_all(); }
        __syncwarp();
        // ---- epilogue
        if (diverge > 0) {                         // outputs below the failed step stay zero (quirk Q10)
            const int upto = diverge;              // steps 0 .. diverge-1 (0-based)
            for (long long e = lane; e < (long long)upto * 128; e += 32) st2(Kb + 2 * e, 0.0, 0.0);
            for (long long e = lane; e < (long long)upto * 8; e += 32) kb[e] = 0.0;
            for (long long e = lane; e < (long long)upto * 32; e += 32) Vxb[e] = 0.0;
            if (Vxxb) {
                for (long long e = lane; e < (long long)upto * 512; e += 32) st2(Vxxb + 2 * e, 0.0, 0.0);
                if (diverge < N - 1)               // Vxx(diverge) (0-based) is the last one computed
                    for (int c = lane; c < 512; c += 32) {
                        int col = c >> 4, r = (c & 15) << 1;
                        double2 t = ld2(&sV[swz(r, col)]);
                        st2(Vxxb + (long long)diverge * 1024 + col * 32 + r, t.x, t.y);
                    }
            }
        } else if (Vxxb && N >= 2) {
            for (int c = lane; c < 512; c += 32) {
                int col = c >> 4, r = (c & 15) << 1;
                double2 t = ld2(&sV[swz(r, col)]);
                st2(Vxxb + col * 32 + r, t.x, t.y);
            }
        }
        if (HIST && P.Vxx_tri) {
            if (diverge > 0) {
                for (long long e = lane; e < (long long)diverge * 528; e += 32) P.Vxx_tri[b * (long long)N * 528 + e] = 0.0;
                if (diverge < N - 1) dump_tri(diverge);
            } else if (N >= 2) dump_tri(0);
        }
        if (P.Vxx1) {
            for (int c = lane; c < 512; c += 32) {
                int col = c >> 4, r = (c & 15) << 1;
                double2 t = (diverge > 0) ? make_double2(0.0, 0.0) : ld2(&sV[swz(r, col)]);
                st2(P.Vxx1 + b * 1024 + col * 32 + r, t.x, t.y);
            }
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { dV0 += shx(dV0, o); dV1 += shx(dV1, o); }      // sum over the eight groups
        if (lane == 0) {
            P.diverge[b] = diverge;
            P.dV[2 * b] = dV0;
            P.dV[2 * b + 1] = dV1;
        }
    }
#undef FRAG
#undef MIRR
}

bool aligned16(const TensorD& t) { return ((uintptr_t)t.p % 16 == 0) && (t.sb % 2 == 0) && (t.st % 2 == 0); }

}  // namespace

int launch_back_pass_tile(ddp_handle_s* h, const BackParams& P_in, bool gps, bool* handled) {
    *handled = false;
    BackParams P = P_in;
    if (P.n != 32 || P.m != 8 || P.T < 2) return 0;
    const bool lims = (P.lims != nullptr);
    if (lims && (P.lims_st != 0 || !P.u.p || getenv("DDP_TILE_NO_LIMS"))) return 0;       // time-varying limits: generic kernel
    if (P.fxx.p || P.fxu.p || P.fuu.p || P.Quu_tri || P.Quui_tri) return 0;          // second-order terms / packed Quu: generic kernel
    if (!aligned16(P.fx) || !aligned16(P.fu) || !aligned16(P.cxx)) return 0;
    if (((uintptr_t)P.K % 16) || (P.Vxx && ((uintptr_t)P.Vxx % 16)) || (P.Vxx1 && ((uintptr_t)P.Vxx1 % 16))) return 0;
    if (gps && !aligned16(P.Kp)) return 0;
    if (!aligned16(P.cx) || !aligned16(P.cu)) return 0;
    const bool ltv = (P.fx.st != 0 || P.fu.st != 0);
    const int WPB = wpb(ltv);
    const size_t bytes = ((size_t)((ltv ? WARP_DOUBLES_LTV : WARP_DOUBLES) + (lims ? QP_DOUBLES : 0)) * WPB + COST_DOUBLES) * sizeof(double);
    long long grid = (long long)h->sm_count * 2;
    long long need = (P.B + WPB - 1) / WPB;
    if (grid > need) grid = need;
    cudaError_t e = cudaSuccess;
    {
        const int rc0 = prepare_redo(h, P);    // hand-over mask for trajectories with an unsymmetric terminal cxx
        if (rc0 != 0) return rc0;
    }
#define LAUNCH_TILE1(L, G, R2, H, LM)                                                                                           \
    do {                                                                                                                        \
        e = cudaFuncSetAttribute(bp_tile32x8_kernel<L, G, R2, H, LM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); \
        if (e == cudaSuccess) bp_tile32x8_kernel<L, G, R2, H, LM><<<(unsigned)grid, WPB * 32, bytes, h->stream>>>(P);            \
    } while (0)
    // the history outputs are compiled into the box-QP variants unconditionally (fewer instantiations of a rarely timed path)
#define LAUNCH_TILE(L, G, R2)                                \
    do {                                                     \
        if (lims) LAUNCH_TILE1(L, G, R2, true, true);        \
        else if (hist) LAUNCH_TILE1(L, G, R2, true, false);  \
        else LAUNCH_TILE1(L, G, R2, false, false);           \
    } while (0)
    const bool hist = (P.Vxx != nullptr) || (P.Vxx_tri != nullptr);
    const bool r2 = !gps && (P.reg_type == 2);
    if (ltv && gps) LAUNCH_TILE(true, true, false);
    else if (ltv && r2) LAUNCH_TILE(true, false, true);
    else if (ltv) LAUNCH_TILE(true, false, false);
    else if (gps) LAUNCH_TILE(false, true, false);
    else if (r2) LAUNCH_TILE(false, false, true);
    else {
        const char* ex = getenv("DDP_TILE_EXP");
        const int exv = (ex && !lims && !hist) ? atoi(ex) : 0;
#define LAUNCH_EXP(V)                                                                                                                          \
    do {                                                                                                                                       \
        e = cudaFuncSetAttribute(bp_tile32x8_kernel<false, false, false, false, false, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); \
        if (e == cudaSuccess) bp_tile32x8_kernel<false, false, false, false, false, V><<<(unsigned)grid, WPB * 32, bytes, h->stream>>>(P);          \
    } while (0)
        if (exv == 1) LAUNCH_EXP(1);
        else if (exv == 2) LAUNCH_EXP(2);
        else if (exv == 3) LAUNCH_EXP(3);
        else if (exv == 4) {
            const size_t bytes3 = bytes - COST_DOUBLES * sizeof(double);
            long long grid3 = (long long)h->sm_count * 3;
            if (grid3 > need) grid3 = need;
            e = cudaFuncSetAttribute(bp_tile32x8_kernel<false, false, false, false, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes3);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(bp_tile32x8_kernel<false, false, false, false, false, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            if (getenv("DDP_TILE_DEBUG")) {
                int nb = -1;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bp_tile32x8_kernel<false, false, false, false, false, 4>, WPB * 32, bytes3);
                fprintf(stderr, "[ddp] tile EXP=4: %d CTAs/SM, %zu B smem per CTA, grid %lld\n", nb, bytes3, grid3);
            }
            if (e == cudaSuccess) bp_tile32x8_kernel<false, false, false, false, false, 4><<<(unsigned)grid3, WPB * 32, bytes3, h->stream>>>(P);
        }
        else if (exv == 8) LAUNCH_EXP(8);
        else LAUNCH_TILE(false, false, false);
#undef LAUNCH_EXP
    }
#undef LAUNCH_TILE
#undef LAUNCH_TILE1
    if (e != cudaSuccess) return (int)e;
    h->launches++;
    *handled = true;
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    return launch_back_pass_generic(h, P, gps);     // processes the handed-over trajectories; exits at once when there are none
}
