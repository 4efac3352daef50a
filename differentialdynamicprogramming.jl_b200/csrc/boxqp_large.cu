// Box-constrained QP for LARGE problems (n up to 1024; the reference's demoQP is n = 500, src/boxQP.jl:190-199): the same
// projected-Newton iteration as boxQP(H,g,lower,upper,x0) (src/boxQP.jl:29-188), one CTA per problem, H in global memory.
//
// The m <= 16 kernels (boxqp.cuh) run one QP per thread in the oracle's sequential arithmetic order.  Here the work of ONE QP is
// spread over 256 threads:
//   H x, H (x .* clamped)   thread per row, columns walked in order (coalesced: H is column-major) -- the oracle's summation order
//   Cholesky of H[free,free] right-looking in place in the compact work matrix: element (p,q) receives its subtractions
//                           R[k,p] R[k,q] for k = 0,1,2,... in that order, separate multiply and subtract -- the operations and the
//                           order of the oracle's sequential factorisation, so the factor agrees with it bit for bit
//   R' y = b                column sweep (ascending), again the oracle's order;  R z = y: column sweep (descending: the order
//                           of the reference BLAS dtrsv, not of the oracle's row loop -- last-bit differences)
//   dot products / norms    fixed-shape tree reductions (deterministic, not the oracle's left-to-right sums)
// Branch decisions (clamped set, result code) therefore agree with the oracle except on exact ties; tests compare the result code
// and the free set exactly and x to 1e-9.
#include <algorithm>
#include "boxqp.cuh"

namespace {

constexpr int LT = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {      // deterministic tree over LT partial sums
    const int tid = threadIdx.x;
    __syncthreads();
    red[tid] = v;
    __syncthreads();
    for (int o = LT / 2; o > 0; o >>= 1) {
        if (tid < o) red[tid] = DADD(red[tid], red[tid + o]);
        __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
}

// y_i = sum_j H[i,j] * (mask ? (clamped_j ? v_j : v_j * 0) : v_j), columns in order (the oracle's _matvec_seq)
__device__ __forceinline__ void matvec(int n, const double* __restrict__ H, const double* v, const unsigned char* clamped, bool masked, double* y) {
    for (int i = threadIdx.x; i < n; i += LT) {
        double s = 0.0;
        for (int j = 0; j < n; j++) {
            const double vj = masked ? (clamped[j] ? v[j] : DMUL(v[j], 0.0)) : v[j];
            s = DADD(s, DMUL(H[i + (size_t)n * j], vj));
        }
        y[i] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(LT) boxqp_large_kernel(int n, long long B, const double* __restrict__ Hall, const double* __restrict__ gall,
                                                         const double* __restrict__ lall, const double* __restrict__ uall,
                                                         const double* __restrict__ x0all, QPOpts o, double* __restrict__ xall,
                                                         int* __restrict__ result, double* __restrict__ work_all, unsigned char* __restrict__ free_all,
                                                         int* __restrict__ nfactor_all) {
    extern __shared__ double sm[];
    double* x = sm;                 // current iterate
    double* hx = x + n;             // H x
    double* grad = hx + n;
    double* search = grad + n;
    double* xc = search + n;        // candidate
    double* hxc = xc + n;           // H xc
    double* s = hxc + n;            // right-hand side / solution of the triangular solves (compact)
    double* red = s + n;            // LT
    int* idx = reinterpret_cast<int*>(red + LT);                                  // free indices (compact -> full)
    unsigned char* clamped = reinterpret_cast<unsigned char*>(idx + n);
    __shared__ int sh_nf, sh_fail, sh_nclamped, sh_changed;
    const int tid = threadIdx.x;
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        const double* H = Hall + (size_t)b * n * n;
        const double* g = gall + (size_t)b * n;
        const double* lower = lall + (size_t)b * n;
        const double* upper = uall + (size_t)b * n;
        double* R = work_all + (size_t)b * n * n;            // compact factor, leading dimension n
        __syncthreads();
        for (int i = tid; i < n; i += LT) { x[i] = clampd(x0all[(size_t)b * n + i], lower[i], upper[i]); clamped[i] = 0; }   // boxQP.jl:58
        __syncthreads();
        matvec(n, H, x, nullptr, false, hx);
        double part = 0.0, part2 = 0.0;
        for (int i = tid; i < n; i += LT) { part = DADD(part, DMUL(x[i], g[i])); part2 = DADD(part2, DMUL(DMUL(0.5, x[i]), hx[i])); }
        double value = DADD(block_sum(part, red), block_sum(part2, red));           // :63
        double oldvalue = 0.0;
        int res = 0, nfactor = 0, iter = 1, nf = 0;
        while (iter <= o.max_iter) {                                               // :71
            if (res != 0) break;
            if (iter > 1 && DSUB(oldvalue, value) < DMUL(o.min_rel_improve, fabs(oldvalue))) { res = 4; break; }     // :78
            oldvalue = value;
            if (tid == 0) { sh_nclamped = 0; sh_changed = 0; sh_fail = 0; }
            __syncthreads();
            int my_cl = 0, my_ch = 0;
            for (int i = tid; i < n; i += LT) {
                const double gi = DADD(g[i], hx[i]);                                // :85
                grad[i] = gi;
                const unsigned char c = ((x[i] == lower[i] && gi > 0.0) || (x[i] == upper[i] && gi < 0.0)) ? 1 : 0;    // :92-94
                if (c != clamped[i]) my_ch = 1;
                clamped[i] = c;
                my_cl += c;
            }
            if (my_cl) atomicAdd(&sh_nclamped, my_cl);
            if (my_ch) atomicOr(&sh_changed, 1);
            __syncthreads();
            if (sh_nclamped == n) { res = 6; break; }                              // :98
            const bool factorize = (iter == 1) || (sh_changed != 0);               // :104-108
            __syncthreads();
            if (factorize) {
                if (tid == 0) {
                    int c = 0;
                    for (int i = 0; i < n; i++)
                        if (!clamped[i]) idx[c++] = i;
                    sh_nf = c;
                }
                __syncthreads();
                nf = sh_nf;
                for (int q = 0; q < nf; q++)                                        // gather the upper triangle of H[free,free]
                    for (int p = tid; p <= q; p += LT) R[p + (size_t)n * q] = H[idx[p] + (size_t)n * idx[q]];
                __syncthreads();
                for (int k = 0; k < nf; k++) {                                      // :111 cholesky(H[free,free]).U, right-looking
                    if (tid == 0) {
                        const double d = R[k + (size_t)n * k];
                        if (!(d > 0.0)) sh_fail = 1;
                        else R[k + (size_t)n * k] = __dsqrt_rn(d);
                    }
                    __syncthreads();
                    if (sh_fail) break;
                    const double rkk = R[k + (size_t)n * k];
                    for (int q = k + 1 + tid; q < nf; q += LT) {                     // row k of the factor, also staged in shared memory
                        const double v = DDIV(R[k + (size_t)n * q], rkk);
                        R[k + (size_t)n * q] = v;
                        s[q] = v;
                    }
                    __syncthreads();
                    for (int q = k + 1; q < nf; q++) {
                        const double rkq = s[q];
                        for (int p = k + 1 + tid; p <= q; p += LT)
                            R[p + (size_t)n * q] = DSUB(R[p + (size_t)n * q], DMUL(s[p], rkq));
                    }
                    __syncthreads();
                }
                if (sh_fail) { res = -1; break; }                                   // PosDefException
                nfactor++;
            }
            nf = sh_nf;
            part = 0.0;
            for (int p = tid; p < nf; p += LT) part = DADD(part, DMUL(grad[idx[p]], grad[idx[p]]));
            const double gnorm = __dsqrt_rn(block_sum(part, red));                 // :120
            if (gnorm < o.min_grad) { res = 5; break; }
            matvec(n, H, x, clamped, true, hxc);                                    // H (x .* clamped)   :127
            for (int p = tid; p < nf; p += LT) s[p] = DADD(g[idx[p]], hxc[idx[p]]);
            __syncthreads();
            for (int p = 0; p < nf; p++) {                                          // R' y = grad_clamped[free]
                if (tid == 0) s[p] = DDIV(s[p], R[p + (size_t)n * p]);
                __syncthreads();
                const double yp = s[p];
                for (int q = p + 1 + tid; q < nf; q += LT) s[q] = DSUB(s[q], DMUL(R[p + (size_t)n * q], yp));
                __syncthreads();
            }
            for (int p = nf - 1; p >= 0; p--) {                                     // R z = y
                if (tid == 0) s[p] = DDIV(s[p], R[p + (size_t)n * p]);
                __syncthreads();
                const double zp = s[p];
                for (int q = tid; q < p; q += LT) s[q] = DSUB(s[q], DMUL(R[q + (size_t)n * p], zp));
                __syncthreads();
            }
            for (int i = tid; i < n; i += LT) search[i] = 0.0;
            __syncthreads();
            for (int p = tid; p < nf; p += LT) search[idx[p]] = DSUB(-s[p], x[idx[p]]);      // :129
            __syncthreads();
            part = 0.0;
            for (int i = tid; i < n; i += LT) part = DADD(part, DMUL(search[i], grad[i]));
            const double sdotg = block_sum(part, red);                              // :132
            if (sdotg >= 0.0) break;                                                // :133 leaves result == 0
            double step = 1.0, vc = 0.0;                                            // :138
            for (;;) {
                for (int i = tid; i < n; i += LT) xc[i] = clampd(DADD(x[i], DMUL(step, search[i])), lower[i], upper[i]);
                __syncthreads();
                matvec(n, H, xc, nullptr, false, hxc);
                part = 0.0; part2 = 0.0;
                for (int i = tid; i < n; i += LT) { part = DADD(part, DMUL(xc[i], g[i])); part2 = DADD(part2, DMUL(DMUL(0.5, xc[i]), hxc[i])); }
                vc = DADD(block_sum(part, red), block_sum(part2, red));
                if (!(DDIV(DSUB(vc, oldvalue), DMUL(step, sdotg)) < o.armijo)) break;      // :142
                step = DMUL(step, o.step_dec);
                if (step < o.min_step) {                                            // :147 (the reference recomputes xc, vc with this step first)
                    for (int i = tid; i < n; i += LT) xc[i] = clampd(DADD(x[i], DMUL(step, search[i])), lower[i], upper[i]);
                    __syncthreads();
                    matvec(n, H, xc, nullptr, false, hxc);
                    part = 0.0; part2 = 0.0;
                    for (int i = tid; i < n; i += LT) { part = DADD(part, DMUL(xc[i], g[i])); part2 = DADD(part2, DMUL(DMUL(0.5, xc[i]), hxc[i])); }
                    vc = DADD(block_sum(part, red), block_sum(part2, red));
                    res = 2;
                    break;
                }
            }
            for (int i = tid; i < n; i += LT) { x[i] = xc[i]; hx[i] = hxc[i]; }     // :161
            __syncthreads();
            value = vc;
            iter++;
        }
        if (iter == o.max_iter) res = 1;                                            // :167 (quirk Q4)
        __syncthreads();
        for (int i = tid; i < n; i += LT) {
            xall[(size_t)b * n + i] = x[i];
            if (free_all) free_all[(size_t)b * n + i] = clamped[i] ? 0 : 1;
        }
        // rows/columns of the work matrix beyond the factor are cleared so that Hfree[1:nfree,1:nfree] is the factor, zeros elsewhere
        for (size_t e = tid; e < (size_t)n * n; e += LT) {
            const int p = (int)(e % n), q = (int)(e / n);
            const bool keep = (p <= q && q < nf && res != -1);
            if (!keep) R[e] = 0.0;
        }
        if (tid == 0) {
            result[b] = res;
            if (nfactor_all) nfactor_all[b] = nfactor;
        }
    }
}

}  // namespace

extern "C" int ddp_boxqp_large_f64(ddp_handle_t h, int32_t n, int64_t B, const double* H, const double* g, const double* lower, const double* upper,
                                   const double* x0, const ddp_boxqp_opts* opts, double* x, int32_t* result, double* Hfree, uint8_t* free_out,
                                   int32_t* nfactor) {
    if (!h) return DDP_ERR_INVALID;
    if (!H || !g || !lower || !upper || !x0 || !x || !result || !Hfree) {
        h->err = "ddp_boxqp_large_f64: H, g, lower, upper, x0, x, result, Hfree are required (Hfree doubles as the n x n work matrix)";
        return DDP_ERR_INVALID;
    }
    if (n < 1 || n > 1024 || B < 0) { h->err = "ddp_boxqp_large_f64: need 1 <= n <= 1024 and B >= 0"; return DDP_ERR_UNSUPPORTED; }
    QPOpts q{100, 1e-8, 1e-8, 0.6, 1e-22, 0.1};   // boxQP.jl:29-36
    if (opts && opts->max_iter > 0) { q.max_iter = opts->max_iter; q.min_grad = opts->min_grad; q.min_rel_improve = opts->min_rel_improve;
                                      q.step_dec = opts->step_dec; q.min_step = opts->min_step; q.armijo = opts->armijo; }
    if (B == 0) return DDP_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return DDP_ERR_CUDA; }
    const size_t bytes = sizeof(double) * (7 * (size_t)n + LT) + sizeof(int) * (size_t)n + (size_t)n + 16;
    cudaError_t e = cudaFuncSetAttribute(boxqp_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) {
        const unsigned grid = (unsigned)std::min<long long>(B, (long long)h->sm_count * 2);
        boxqp_large_kernel<<<grid, LT, bytes, h->stream>>>(n, B, H, g, lower, upper, x0, q, x, result, Hfree, free_out, nfactor);
        h->launches++;
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { h->err = std::string("ddp_boxqp_large_f64: ") + cudaGetErrorString(e); return DDP_ERR_CUDA; }
    return DDP_OK;
}
