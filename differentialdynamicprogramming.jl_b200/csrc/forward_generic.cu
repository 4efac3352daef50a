// Generic forward line-search rollout: any n <= 64, m <= 16, linear or pendcart model.
// One warp walks one trajectory forward in time; 4 trajectories per CTA.
// Replaces forward_pass of src/forward_pass.jl:9-33 with the user callbacks f / costfun
// replaced by device model descriptors (demo_linear.jl:35-50, system_pendcart.jl:83-106).
#include "ddp_common.cuh"

namespace {

constexpr int WPB = 4;   // warps (trajectories) per block

__host__ __device__ inline size_t fwd_smem_doubles(int n, int m) {
    int ldn = n | 1;
    return (size_t)ldn * n + (size_t)ldn * m + 4 * (size_t)n + 2 * (size_t)m + 4;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(WPB * 32) fwd_generic_kernel(FwdParams P) {
    extern __shared__ double smem_raw[];
    const int n = P.n, m = P.m, N = P.T, ldn = n | 1;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* base = smem_raw + (size_t)w * fwd_smem_doubles(n, m);
    double* sA = base;
    double* sB = sA + ldn * n;
    double* sx = sB + ldn * m;
    double* sdx = sx + n;
    double* sd = sdx + n;
    double* sxn = sd + n;
    double* su = sxn + n;
    double* sru = su + m;
    const bool linear = (P.model.kind == DDP_MODEL_LINEAR);
    const bool policy = (P.K != nullptr);
    const bool lti = (P.model.A.st == 0 && P.model.Bm.st == 0);
    const long long mn = (long long)m * n;

    for (long long b = (long long)blockIdx.x * WPB + w; b < P.B; b += (long long)gridDim.x * WPB) {
        if (P.active && !P.active[b]) continue;
        const double alpha = P.alpha ? P.alpha[b] : P.alpha_scalar;
        const double* Qm = P.model.Q.p + b * P.model.Q.sb;
        const double* Rm = P.model.R.p + b * P.model.R.sb;
        const double* x0 = P.x0.p + b * P.x0.sb;
        double* xnb = P.xnew + b * (long long)N * n;
        double* unb = P.unew + b * (long long)N * m;
        __syncwarp();
        for (int i = lane; i < n; i += 32) sx[i] = x0[i];
        double cpart = 0.0;
        for (int t = 0; t < N; t++) {
            if (linear && (!lti || t == 0)) {
                const double* At = tp(P.model.A, b, t);
                const double* Bt = tp(P.model.Bm, b, t);
                for (int e = lane; e < n * n; e += 32) sA[(e % n) + ldn * (e / n)] = At[e];
                for (int e = lane; e < n * m; e += 32) sB[(e % n) + ldn * (e / n)] = Bt[e];
            }
            __syncwarp();
            const double* ut = tp(P.u, b, t);
            for (int i = lane; i < n; i += 32) {
                double xv = sx[i];
                xnb[(long long)t * n + i] = xv;
                if (policy) sdx[i] = xv - tp(P.x, b, t)[i];              // diff = -, forward_pass.jl:19
                sd[i] = xv - (P.model.goal ? P.model.goal[i] : 0.0);
            }
            __syncwarp();
            // ---- control update  u + αk + K dx, clamp   (forward_pass.jl:17-24)
            if (policy && (32 % m) == 0) {
                const double* Kt = P.K + (b * N + t) * mn;
                const int a = lane % m, per = 32 / m;
                double acc = 0.0;
                for (int j = lane / m; j < n; j += per) acc = fma(Kt[a + m * j], sdx[j], acc);
                for (int o = m; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane < m) {
                    double un = ut[a] * P.u_scale;
                    un = un + P.k[(b * N + t) * m + a] * alpha;
                    un = un + acc;
                    if (P.lims) { const double* li = P.lims + (long long)t * P.lims_st; un = clamp_jl(un, li[a], li[m + a]); }   // NaN survives the clamp, as in Julia
                    if (un != un) un = 0.0;                                // u[isnan.(u)] .= 0 in f
                    su[a] = un;
                    unb[(long long)t * m + a] = un;
                }
            } else {
                for (int a = lane; a < m; a += 32) {
                    double un = ut[a] * P.u_scale;
                    if (policy) {
                        const double* Kt = P.K + (b * N + t) * mn;
                        un = un + P.k[(b * N + t) * m + a] * alpha;
                        double acc = 0.0;
                        for (int j = 0; j < n; j++) acc = fma(Kt[a + m * j], sdx[j], acc);
                        un = un + acc;
                    }
                    if (P.lims) { const double* li = P.lims + (long long)t * P.lims_st; un = clamp_jl(un, li[a], li[m + a]); }   // NaN survives the clamp, as in Julia
                    if (un != un) un = 0.0;
                    su[a] = un;
                    unb[(long long)t * m + a] = un;
                }
            }
            __syncwarp();
            // ---- running cost ½ d'Qd + ½ u'Ru  and fused cx = Q d, cu = R u
            double cstep = 0.0;
            for (int i = lane; i < n; i += 32) {
                double qd = 0.0;
                for (int j = 0; j < n; j++) qd = fma(Qm[i + n * j], sd[j], qd);
                cstep = fma(0.5 * sd[i], qd, cstep);
                if (P.cx) P.cx[(b * N + t) * n + i] = qd;
            }
            for (int a = lane; a < m; a += 32) {
                double ru = 0.0;
                for (int c = 0; c < m; c++) ru = fma(Rm[a + m * c], su[c], ru);
                cstep = fma(0.5 * su[a], ru, cstep);
                if (P.cu) P.cu[(b * N + t) * m + a] = ru;
            }
            if (P.cost_t) {
                double ct = warp_sum(cstep);
                if (lane == 0) P.cost_t[b * (N + P.model.terminal_cost) + t] = ct;
            }
            cpart += cstep;
            // ---- dynamics (called at t = N-1 too in the reference, result discarded)
            if (t < N - 1) {
                if (linear) {
                    for (int i = lane; i < n; i += 32) {
                        double ax = 0.0, bu = 0.0;
                        for (int j = 0; j < n; j++) ax = fma(sA[i + ldn * j], sx[j], ax);
                        for (int a = 0; a < m; a++) bu = fma(sB[i + ldn * a], su[a], bu);
                        sxn[i] = ax + bu;
                    }
                } else if (lane == 0) {                                   // pendcart dfsys :83-89
                    const double g = P.model.p[0], l = P.model.p[1], h = P.model.p[2], d = P.model.p[3];
                    double sn, cs;
                    sincos(sx[0], &sn, &cs);
                    sxn[0] = sx[0] + h * sx[1];
                    sxn[1] = sx[1] + h * (-g / l * sn + su[0] / l * cs - d * sx[1]);
                    sxn[2] = sx[2] + h * sx[3];
                    sxn[3] = sx[3] + h * su[0];
                }
                __syncwarp();
                for (int i = lane; i < n; i += 32) sx[i] = sxn[i];
            }
            __syncwarp();
        }
        if (P.model.terminal_cost) {                                      // system_pendcart.jl:104
            double cterm = 0.0;
            for (int i = lane; i < n; i += 32) {
                double qd = 0.0;
                for (int j = 0; j < n; j++) qd = fma(Qm[i + n * j], sd[j], qd);
                cterm = fma(0.5 * sd[i], qd, cterm);
            }
            if (P.cost_t) {
                double ct = warp_sum(cterm);
                if (lane == 0) P.cost_t[b * (N + 1) + N] = ct;
            }
            cpart += cterm;
        }
        double ctot = warp_sum(cpart);
        if (lane == 0) P.cost[b] = ctot;
    }
}

}  // namespace

int launch_forward_generic(ddp_handle_s* h, const FwdParams& P) {
    size_t bytes = fwd_smem_doubles(P.n, P.m) * sizeof(double) * WPB;
    if ((long long)bytes > h->max_smem_optin) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(fwd_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    int per_sm = (int)((size_t)h->max_smem_optin / (bytes + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    long long grid = (long long)h->sm_count * per_sm;
    long long need = (P.B + WPB - 1) / WPB;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    fwd_generic_kernel<<<(unsigned)grid, WPB * 32, bytes, h->stream>>>(P);
    h->launches++;
    return (int)cudaGetLastError();
}
