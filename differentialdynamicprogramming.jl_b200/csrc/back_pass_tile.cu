// Specialised backward sweep for n = 32, m = 8 (the headline shape): ONE WARP PER TRAJECTORY.
//
// Replaces back_pass (Cholesky branch) of src/backward_pass.jl:162-252 + :31-42, :64-76.
//
// The per-step dense algebra is mapped onto FP64 tensor tiles (mma.sync.m8n8k4.f64, "DMMA";
// measured 37.1 TFLOP/s on B200 vs 34.1 for the DFMA pipe, profiles/microbench) -- tcgen05 has
// no f64 kind, so this is the only matrix unit that keeps the 1e-8 FP64 parity contract:
//
//   F  = [fx fu]            32 x 40    (shared memory, column-major, XOR-swizzled)
//   W' = F' V               40 x 32    160 DMMA   (V = Vxx(i+1), symmetric, shared memory)
//   G  = W' F = F' V F      40 x 40    120 DMMA   (upper 15 of 25 tiles: G is symmetric)
//        G = [Qxx Qxu; . Quu] - cost terms; the W' accumulators are re-used directly as the
//        A operand of the second product by permuting the contraction index (no smem round trip)
//   Quu (8x8): every lane factors it redundantly in registers (reciprocal square roots), lane j
//        solves for column j of K; k and Quu*k are computed warp-uniformly
//   Vxx = Qxx + K'(Quu K + Qux) + Qux' K   40 DMMA onto the resident Qxx tiles, mirrored => exactly symmetric
//
// 328 DMMA per step ~ 168 kflop, against the reference's 213 kflop formulation.
// Shared memory: 26.6 KB per warp => 8 warps (trajectories) per SM, 2 per SM sub-partition, so
// one warp's serial Cholesky/solve phase overlaps the other's tensor phase.
//
// Restrictions (anything else dispatches to the generic kernel): Cholesky branch only (lims ==
// NULL), 16-byte aligned fx/fu with even strides, symmetric cxx (it is a Hessian).
#include "ddp_common.cuh"

namespace {

constexpr int WPB = 4;                       // warps per CTA
constexpr int SV = 0;                        // Vxx            32 x 32 swizzled
constexpr int SF = SV + 1024;                // [fx fu]        32 x 40 swizzled
constexpr int SQUX = SF + 1280;              // Qux  (8 x 32, column j at 8j)
constexpr int SK = SQUX + 256;               // K
constexpr int SM1 = SK + 256;                // Quu K + Qux   (aliases Qux_reg before the solve)
constexpr int SQUU = SM1 + 256;              // Quu  (col-major 8 x 8)
constexpr int SQUUF = SQUU + 64;             // regularised Quu
constexpr int SVX = SQUUF + 64;              // Vx (32)
constexpr int SQU = SVX + 32;                // Qu (8)
constexpr int SFV = SQU + 8;                 // F'Vx (32)
constexpr int WARP_DOUBLES = SFV + 32 + 8;   // = 3280 doubles -> 26,240 B per warp
constexpr int SF2 = WARP_DOUBLES;            // second F buffer (time-varying dynamics only)
constexpr int WARP_DOUBLES_LTV = WARP_DOUBLES + 1280;

__device__ __forceinline__ int swz(int i, int c) { return (i ^ (((c & 1) << 3) | (((c >> 1) & 3) << 1))) + 32 * c; }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

__device__ __forceinline__ void cp_async16(double* dst_smem, const double* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// stage [fx fu] of one step into the swizzled buffer (16-byte chunks = 2 consecutive rows of a column)
template <bool ASYNC>
__device__ __forceinline__ void load_F(double* sF, const double* fx, const double* fu, int lane) {
#pragma unroll 4
    for (int c = lane; c < 640; c += 32) {
        int col = c >> 4, i = (c & 15) << 1;
        const double* src = (col < 32) ? (fx + col * 32 + i) : (fu + (col - 32) * 32 + i);
        if (ASYNC) cp_async16(&sF[swz(i, col)], src);
        else st2(&sF[swz(i, col)], src[0], src[1]);
    }
}

// 8 x 32 buffers (Qux, K, M1): column j at 8j, its four 16-byte row pairs XOR-swizzled by (j >> 1) so that both
// "lane = column" accesses and the DMMA fragment loads are bank-conflict free
__device__ __forceinline__ int cidx(int r, int j) { return 8 * j + ((((r >> 1) ^ (j >> 1)) & 3) << 1) + (r & 1); }

// 1/sqrt(d) for d > 0, normal range: hardware seed + two Newton steps (the Cholesky pivots are O(1e-3..1e3);
// non-positive pivots are caught before the value is used)
__device__ __forceinline__ double rsqrt_nr(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double h = 0.5 * d;
    double e = fma(-h * y, y, 0.5);      // 0.5 - h y^2
    y = fma(y, e, y);
    e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    return y;
}

constexpr int gidx(int at, int bt) { return at * 5 - (at * (at - 1)) / 2 + (bt - at); }   // upper-tile index, 15 tiles

template <bool LTV>
__global__ void __launch_bounds__(WPB * 32, 2) bp_tile32x8_kernel(BackParams P) {
    extern __shared__ double smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    double* sm = smem_raw + (size_t)w * (LTV ? WARP_DOUBLES_LTV : WARP_DOUBLES);
    double* sV = sm + SV;
    double* sQux = sm + SQUX;
    double* sK = sm + SK;
    double* sM1 = sm + SM1;
    double* sQuu = sm + SQUU;
    double* sQuuF = sm + SQUUF;
    double* sVx = sm + SVX;
    double* sQu = sm + SQU;
    double* sFV = sm + SFV;
    const int N = P.T;
    const long long warps_total = (long long)gridDim.x * WPB;

    for (long long b = (long long)blockIdx.x * WPB + w; b < P.B; b += warps_total) {
        if (P.active && !P.active[b]) continue;
        const double lam = P.lambda[b];
        const bool reg2 = (P.reg_type == 2);
        double* Kb = P.K + b * (long long)N * 256;
        double* kb = P.k + b * (long long)N * 8;
        double* Vxb = P.Vx + b * (long long)N * 32;
        double* Vxxb = P.Vxx ? P.Vxx + b * (long long)N * 1024 : nullptr;
        double* Quub = P.Quu ? P.Quu + b * (long long)N * 64 : nullptr;
        __syncwarp();
        // ---- terminal step
        {
            const double* cxN = tp(P.cx, b, N - 1);
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
            for (int c = lane; c < 128; c += 32) st2(Kb + (long long)(N - 1) * 256 + 2 * c, 0.0, 0.0);
            if (lane < 8) kb[(long long)(N - 1) * 8 + lane] = 0.0;
            if (Quub) st2(Quub + (long long)(N - 1) * 64 + 2 * lane, cuuN[2 * lane], cuuN[2 * lane + 1]);
        }
        int buf = 0;
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
                double2 fb = ld2(&sF[swz(8 * p + 2 * q, 32 + g)]);
#pragma unroll
                for (int t = 0; t < 5; t++) {
                    double2 fa = ld2(&sF[swz(8 * p + 2 * q, 8 * t + g)]);
                    dmma(FF[t][0], FF[t][1], fa.x, fb.x);
                    dmma(FF[t][0], FF[t][1], fa.y, fb.y);
                }
            }
        };
        if (!LTV && reg2) compute_FF(sm + SF);

        double dV0 = 0.0, dV1 = 0.0;
        int diverge = 0;
        for (int i = N - 2; i >= 0; i--) {
            const double* sF = sm + (LTV ? (buf ? SF2 : SF) : SF);
            if (LTV) {
                cp_async_wait_all();
                __syncwarp();
                if (i > 0) load_F<true>(sm + (buf ? SF : SF2), tp(P.fx, b, i - 1), tp(P.fu, b, i - 1), lane);
                if (reg2) compute_FF(sF);
            }
            // prefetch this step's cost gradients
            const double cxv = tp(P.cx, b, i)[lane];
            const double cuv = tp(P.cu, b, i)[g];             // lanes (g, q == 0) publish Qu[g]
            const double* cxxi = tp(P.cxx, b, i);
            const double* cxui = tp(P.cxu, b, i);
            const double* cuui = tp(P.cuu, b, i);
            // dump Vxx(i+1) history if requested (sV is stable here)
            if (Vxxb && i < N - 2) {
                for (int c = lane; c < 512; c += 32) {
                    int col = c >> 4, r = (c & 15) << 1;
                    double2 t = ld2(&sV[swz(r, col)]);
                    st2(Vxxb + (long long)(i + 1) * 1024 + col * 32 + r, t.x, t.y);
                }
            }
            // ---- G starts from the cost terms: their global loads are in flight during the tensor phase
            double G[15][2];
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
            // ---- step 1: W' = F' V.  The A fragments also give F'Vx: each lane sums its 8 rows, two shuffles finish it
            double W[5][4][2], fv[5];
#pragma unroll
            for (int at = 0; at < 5; at++) {
                fv[at] = 0.0;
#pragma unroll
                for (int jt = 0; jt < 4; jt++) W[at][jt][0] = W[at][jt][1] = 0.0;
            }
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const int row = 8 * p + 2 * q;
                double2 fa[5], fb[4];
#pragma unroll
                for (int at = 0; at < 5; at++) fa[at] = ld2(&sF[swz(row, 8 * at + g)]);
#pragma unroll
                for (int jt = 0; jt < 4; jt++) fb[jt] = ld2(&sV[swz(row, 8 * jt + g)]);
                const double2 vx = ld2(&sVx[row]);
                // consecutive DMMAs go to different accumulator tiles (a tile is revisited 20 issues later)
#pragma unroll
                for (int at = 0; at < 5; at++)
#pragma unroll
                    for (int jt = 0; jt < 4; jt++) dmma(W[at][jt][0], W[at][jt][1], fa[at].x, fb[jt].x);
#pragma unroll
                for (int at = 0; at < 5; at++) fv[at] = fma(fa[at].y, vx.y, fma(fa[at].x, vx.x, fv[at]));
#pragma unroll
                for (int at = 0; at < 5; at++)
#pragma unroll
                    for (int jt = 0; jt < 4; jt++) dmma(W[at][jt][0], W[at][jt][1], fa[at].y, fb[jt].y);
            }
#pragma unroll
            for (int at = 0; at < 5; at++) {
                fv[at] += __shfl_xor_sync(0xffffffffu, fv[at], 1);
                fv[at] += __shfl_xor_sync(0xffffffffu, fv[at], 2);
            }
            if (q == 0) {                                   // (F'Vx)[8at+g]; Qu[g] = cu[g] + (fu'Vx)[g]
#pragma unroll
                for (int at = 0; at < 4; at++) sFV[8 * at + g] = fv[at];
                sQu[g] = cuv + fv[4];
            }
            // ---- step 2: G = W' F  (upper tiles)
#pragma unroll
            for (int p = 0; p < 4; p++) {
                double2 ff[5];
#pragma unroll
                for (int bt = 0; bt < 5; bt++) ff[bt] = ld2(&sF[swz(8 * p + 2 * q, 8 * bt + g)]);
#pragma unroll
                for (int at = 0; at < 5; at++)
#pragma unroll
                    for (int bt = at; bt < 5; bt++) dmma(G[gidx(at, bt)][0], G[gidx(at, bt)][1], W[at][p][0], ff[bt].x);
#pragma unroll
                for (int at = 0; at < 5; at++)
#pragma unroll
                    for (int bt = at; bt < 5; bt++) dmma(G[gidx(at, bt)][0], G[gidx(at, bt)][1], W[at][p][1], ff[bt].y);
            }
            // ---- spill Qux / Qux_reg / Quu / QuuF / Qu to shared memory
#pragma unroll
            for (int at = 0; at < 4; at++) {
                const int a = 8 * at + g;
                double x0 = G[gidx(at, 4)][0], x1 = G[gidx(at, 4)][1];
                st2(&sQux[cidx(2 * q, a)], x0, x1);
                if (reg2) st2(&sM1[cidx(2 * q, a)], fma(lam, FF[at][0], x0), fma(lam, FF[at][1], x1));
            }
            {
                double u0 = G[gidx(4, 4)][0], u1 = G[gidx(4, 4)][1];
                st2(&sQuu[8 * g + 2 * q], u0, u1);          // Quu and QuuF are kept ROW-major: [a][b] at 8a + b
                double f0, f1;
                if (reg2) { f0 = fma(lam, FF[4][0], u0); f1 = fma(lam, FF[4][1], u1); }
                else { f0 = u0 + ((g == 2 * q) ? lam : 0.0); f1 = u1 + ((g == 2 * q + 1) ? lam : 0.0); }
                st2(&sQuuF[8 * g + 2 * q], f0, f1);
            }
            __syncwarp();
            const double qx = cxv + sFV[lane];               // Qx = cx + fx'Vx, owned by lane = state index
            // ---- Cholesky of QuuF (upper triangle), redundantly on every lane
            double R[8][8], rinv[8];
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 8; j++) {
#pragma unroll
                for (int r = 0; r < j; r++) {
                    double s = sQuuF[8 * r + j];
#pragma unroll
                    for (int p = 0; p < r; p++) s = fma(-R[p][r], R[p][j], s);
                    R[r][j] = s * rinv[r];
                }
                double d = sQuuF[9 * j];
#pragma unroll
                for (int p = 0; p < j; p++) d = fma(-R[p][j], R[p][j], d);
                if (!(d > 0.0)) ok = false;
                rinv[j] = rsqrt_nr(d);
            }
            if (!ok) { diverge = i + 1; break; }
            auto solve8 = [&](double (&v)[8]) {
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    double s = v[r];
#pragma unroll
                    for (int p = 0; p < r; p++) s = fma(-R[p][r], v[p], s);
                    v[r] = s * rinv[r];
                }
#pragma unroll
                for (int r = 7; r >= 0; r--) {
                    double s = v[r];
#pragma unroll
                    for (int p = r + 1; p < 8; p++) s = fma(-R[r][p], v[p], s);
                    v[r] = s * rinv[r];
                }
            };
            // ---- gains: lane j owns column j of K; k is warp-uniform
            double Kc[8], Qc[8], kv[8], Quv[8];
            {
                const double* src = reg2 ? sM1 : sQux;
#pragma unroll
                for (int r = 0; r < 8; r += 2) { double2 t = ld2(src + cidx(r, lane)); Kc[r] = t.x; Kc[r + 1] = t.y; }
                if (reg2) {
#pragma unroll
                    for (int r = 0; r < 8; r += 2) { double2 t = ld2(sQux + cidx(r, lane)); Qc[r] = t.x; Qc[r + 1] = t.y; }
                } else {
#pragma unroll
                    for (int r = 0; r < 8; r++) Qc[r] = Kc[r];
                }
#pragma unroll
                for (int r = 0; r < 8; r += 2) { double2 t = ld2(sQu + r); Quv[r] = t.x; Quv[r + 1] = t.y; kv[r] = t.x; kv[r + 1] = t.y; }
            }
            solve8(Kc);
            solve8(kv);
#pragma unroll
            for (int r = 0; r < 8; r++) { Kc[r] = -Kc[r]; kv[r] = -kv[r]; }
            // ---- M1 = Quu K + Qux (column j), Quuk = Quu k (uniform)
            double M1c[8], Quuk[8];
#pragma unroll
            for (int r = 0; r < 8; r++) M1c[r] = Qc[r];
            double quuk_own = 0.0;                 // lane a (mod 8) forms (Quu k)[a]; the 8 values are then shuffled
#pragma unroll
            for (int r = 0; r < 8; r++) {
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    double2 t = ld2(&sQuu[8 * r + c]);              // Quu[r][c], Quu[r][c+1]
                    M1c[r] = fma(t.x, Kc[c], M1c[r]);
                    M1c[r] = fma(t.y, Kc[c + 1], M1c[r]);
                }
            }
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                double2 t = ld2(&sQuu[8 * (lane & 7) + c]);
                quuk_own = fma(t.x, kv[c], quuk_own);
                quuk_own = fma(t.y, kv[c + 1], quuk_own);
            }
#pragma unroll
            for (int r = 0; r < 8; r++) Quuk[r] = __shfl_sync(0xffffffffu, quuk_own, r);
            __syncwarp();     // everyone has read its Qux_reg column (sM1 aliases it)
#pragma unroll
            for (int r = 0; r < 8; r += 2) {
                st2(&sK[cidx(r, lane)], Kc[r], Kc[r + 1]);
                st2(&sM1[cidx(r, lane)], M1c[r], M1c[r + 1]);
            }
            // ---- Vx(i), dV  (backward_pass.jl:64-69)
            double vxn;
            {
                double t1 = 0.0, t2 = 0.0, t3 = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    t1 = fma(Kc[r], Quuk[r], t1);
                    t2 = fma(Kc[r], Quv[r], t2);
                    t3 = fma(Qc[r], kv[r], t3);
                    d0 = fma(kv[r], Quv[r], d0);
                    d1 = fma(kv[r], Quuk[r], d1);
                }
                vxn = ((qx + t1) + t2) + t3;
                dV0 += d0;
                dV1 += 0.5 * d1;
            }
            __syncwarp();
            // ---- outputs of this step that are complete now
            {
                double* Kg = Kb + (long long)i * 256;
#pragma unroll
                for (int c = lane; c < 128; c += 32) { double2 t = ld2(&sK[cidx(2 * (c & 3), c >> 2)]); st2(Kg + 2 * c, t.x, t.y); }
                double ksel = kv[0];
#pragma unroll
                for (int r = 1; r < 8; r++) ksel = (lane == r) ? kv[r] : ksel;
                if (lane < 8) kb[(long long)i * 8 + lane] = ksel;
                Vxb[(long long)i * 32 + lane] = vxn;
                if (Quub) {                                   // column-major out: rows 2l%8, 2l%8+1 of column l/4
                    const int r0 = (2 * lane) & 7, c0 = lane >> 2;
                    st2(Quub + (long long)i * 64 + 2 * lane, sQuu[8 * r0 + c0], sQuu[8 * (r0 + 1) + c0]);
                }
            }
            // ---- step 4: Vxx = Qxx + K' M1 + Qux' K  (upper 10 tiles), mirrored into sV
            {
                double2 kf[4], mf[4], qf[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int o = cidx(2 * q, 8 * t + g);
                    kf[t] = ld2(&sK[o]);
                    mf[t] = ld2(&sM1[o]);
                    qf[t] = ld2(&sQux[o]);
                }
#pragma unroll
                for (int pass = 0; pass < 4; pass++)        // 10 independent tiles between revisits
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
            }
            // all lanes are past their step-1 reads of sV (the two __syncwarp above order them)
#pragma unroll
            for (int at = 0; at < 4; at++) {
                const int a = 8 * at + g;
#pragma unroll
                for (int bt = at; bt < 4; bt++) {
                    const int b0 = 8 * bt + 2 * q;
                    const double v0 = G[gidx(at, bt)][0], v1 = G[gidx(at, bt)][1];
                    if (bt > at) {
                        st2(&sV[swz(b0, a)], v0, v1);
                        sV[swz(a, b0)] = v0;
                        sV[swz(a, b0 + 1)] = v1;
                    } else {                       // diagonal tile: keep the upper part, mirror it
                        if (b0 >= a) { sV[swz(b0, a)] = v0; if (b0 > a) sV[swz(a, b0)] = v0; }
                        if (b0 + 1 >= a) { sV[swz(b0 + 1, a)] = v1; if (b0 + 1 > a) sV[swz(a, b0 + 1)] = v1; }
                    }
                }
            }
            sVx[lane] = vxn;
            if (LTV) buf ^= 1;
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
                if (diverge < N - 1)               // Vxx(diverge) (0-based) was computed but not dumped yet
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
        if (P.Vxx1) {
            for (int c = lane; c < 512; c += 32) {
                int col = c >> 4, r = (c & 15) << 1;
                double2 t = (diverge > 0) ? make_double2(0.0, 0.0) : ld2(&sV[swz(r, col)]);
                st2(P.Vxx1 + b * 1024 + col * 32 + r, t.x, t.y);
            }
        }
        if (lane == 0) {
            P.diverge[b] = diverge;
            P.dV[2 * b] = dV0;
            P.dV[2 * b + 1] = dV1;
        }
    }
}

bool aligned16(const TensorD& t) { return ((uintptr_t)t.p % 16 == 0) && (t.sb % 2 == 0) && (t.st % 2 == 0); }

}  // namespace

int launch_back_pass_tile(ddp_handle_s* h, const BackParams& P, bool gps, bool* handled) {
    *handled = false;
    if (gps || P.n != 32 || P.m != 8 || P.lims != nullptr || P.T < 2) return 0;
    if (!aligned16(P.fx) || !aligned16(P.fu) || !aligned16(P.cxx)) return 0;
    if (P.Vxx && ((uintptr_t)P.Vxx % 16)) return 0;
    const bool ltv = (P.fx.st != 0 || P.fu.st != 0);
    const size_t bytes = (size_t)(ltv ? WARP_DOUBLES_LTV : WARP_DOUBLES) * sizeof(double) * WPB;
    cudaError_t e;
    if (ltv) e = cudaFuncSetAttribute(bp_tile32x8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    else e = cudaFuncSetAttribute(bp_tile32x8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    int per_sm = ltv ? 1 : 2;
    long long grid = (long long)h->sm_count * per_sm;
    long long need = (P.B + WPB - 1) / WPB;
    if (grid > need) grid = need;
    if (ltv) bp_tile32x8_kernel<true><<<(unsigned)grid, WPB * 32, bytes, h->stream>>>(P);
    else bp_tile32x8_kernel<false><<<(unsigned)grid, WPB * 32, bytes, h->stream>>>(P);
    h->launches++;
    *handled = true;
    return (int)cudaGetLastError();
}
