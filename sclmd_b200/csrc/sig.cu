// Surface self-energies by Sancho-Rubio decimation and the transmission built from them
// (replaces sig.sgf / sig.selfenergy / sig.retargf / sig.tm / sig.getse / sig.gettm,
// sclmd/selfenergy.py:105-178).  One CTA per frequency; every m x m complex matrix lives in
// shared memory (m = 24 in examples/runsig.py -> 9 KB each).  Lead blocks too large for that (m > 35) use the same
// code on a per-CTA workspace in global memory (L2-resident for moderate m), up to m = 512.
#include <algorithm>

#include "common.cuh"

using namespace sclmd;

namespace {

constexpr int ST = 256;

struct cplx {
    double x, y;
};
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

// C = op(A) . op(B), all m x m row-major in shared memory; ta/tb = plain transpose
__device__ void mm(cplx *C, const cplx *A, const cplx *B, int m, bool ta, bool tb) {
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        const int i = e / m, j = e % m;
        double sr = 0, si = 0;
        for (int k = 0; k < m; ++k) {
            const cplx a = ta ? A[k * m + i] : A[i * m + k];
            const cplx b = tb ? B[j * m + k] : B[k * m + j];
            sr += a.x * b.x - a.y * b.y;
            si += a.x * b.y + a.y * b.x;
        }
        C[e] = {sr, si};
    }
    __syncthreads();
}

// Ainv = inverse(z I - S) by Gauss-Jordan with partial pivoting on [A | I] (aug: m x 2m)
__device__ void inv_shift(cplx *Ainv, const cplx *S, cplx z, int m, cplx *aug, double *red, int *ired, int *bad) {
    const int w = 2 * m;
    for (int e = threadIdx.x; e < m * w; e += blockDim.x) {
        const int i = e / w, j = e % w;
        cplx v;
        if (j < m) v = {(i == j ? z.x : 0.0) - S[i * m + j].x, (i == j ? z.y : 0.0) - S[i * m + j].y};
        else v = {(j - m == i) ? 1.0 : 0.0, 0.0};
        aug[e] = v;
    }
    __syncthreads();
    for (int k = 0; k < m; ++k) {
        if (threadIdx.x < 32) {   // pivot search by one warp
            double best = -1.0;
            int arg = k;
            for (int i = k + (int)threadIdx.x; i < m; i += 32) {
                const double v = fabs(aug[i * w + k].x) + fabs(aug[i * w + k].y);
                if (v > best) { best = v; arg = i; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
            }
            if (threadIdx.x == 0) {
                *ired = arg;
                if (!(best > 0.0)) *bad = 1;
            }
        }
        __syncthreads();
        const int r = *ired;
        if (r != k)
            for (int j = threadIdx.x; j < w; j += blockDim.x) {
                const cplx t = aug[k * w + j];
                aug[k * w + j] = aug[r * w + j];
                aug[r * w + j] = t;
            }
        __syncthreads();
        const cplx d = aug[k * w + k];
        const double dn = d.x * d.x + d.y * d.y;
        const cplx id = {d.x / dn, -d.y / dn};
        __syncthreads();
        for (int j = threadIdx.x; j < w; j += blockDim.x) aug[k * w + j] = cmulc(aug[k * w + j], id);
        __syncthreads();
        // eliminate column k from every other row; column k of the factors is saved first
        for (int i = threadIdx.x; i < m; i += blockDim.x) red[2 * i] = aug[i * w + k].x, red[2 * i + 1] = aug[i * w + k].y;
        __syncthreads();
        for (int e = threadIdx.x; e < m * w; e += blockDim.x) {
            const int i = e / w, j = e % w;
            if (i == k) continue;
            const cplx f = {red[2 * i], red[2 * i + 1]};
            const cplx p = aug[k * w + j];
            aug[e].x -= f.x * p.x - f.y * p.y;
            aug[e].y -= f.x * p.y + f.y * p.x;
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) Ainv[e] = aug[(e / m) * w + m + e % m];
    __syncthreads();
}

struct SigArgs {
    int m, nw, mode;             // mode 0: self-energy of one lead; 1: transmission; 2: device Green function; 3: surface Green function of one lead
    const double *K00, *K11, *K01, *K10;
    double eta;
    int dirR;                    // mode 0: 1 = 'R', 0 = 'L'
    const double *omegas;
    double *se_out;              // [nw][m][m][2]
    double *tm_out;              // [nw]
    int *iters;                  // [nw] (mode 0) or [nw][2]
    int *status;                 // [nw] 0 ok, 1 not converged, 2 singular
    double *gws;                 // per-CTA matrix workspace in global memory when the matrices do not fit shared memory, else NULL
    size_t gws_stride;           // doubles per CTA
};

// selfenergy.py:105-131: returns g = inv(z - s) in `g`; s,e,al,t1,t2,t3 are m x m work matrices
__device__ int sgf(const SigArgs &a, bool dirR, cplx z, cplx *s, cplx *e, cplx *al, cplx *g, cplx *t1, cplx *t2, cplx *t3, cplx *aug,
                   double *red, int *ired, int *bad, int *notconv) {
    const int m = a.m, mm2 = m * m;
    const double *S0 = dirR ? a.K00 : a.K11, *E0 = dirR ? a.K11 : a.K00, *A0 = dirR ? a.K01 : a.K10;
    for (int i = threadIdx.x; i < mm2; i += blockDim.x) {
        s[i] = {S0[i], 0.0};
        e[i] = {E0[i], 0.0};
        al[i] = {A0[i], 0.0};
    }
    __syncthreads();
    int it = 0;
    while (true) {
        double nr = 0.0;
        for (int i = threadIdx.x; i < mm2; i += blockDim.x) nr += al[i].x * al[i].x + al[i].y * al[i].y;
        nr = block_sum(nr, red);
        if (threadIdx.x == 0) red[40] = nr;
        __syncthreads();
        nr = red[40];
        __syncthreads();
        if (!(sqrt(nr) > 1e-8)) break;              // while np.linalg.norm(alpha) > 1e-8
        inv_shift(g, e, z, m, aug, red, ired, bad);  // g = inv(z - e)
        mm(t1, al, g, m, false, false);              // alpha g
        mm(t2, t1, al, m, false, true);              // alpha g beta,  beta = alpha^T
        mm(t3, al, g, m, true, false);               // beta g
        mm(g, t3, al, m, false, false);              // beta g alpha   (g is free again)
        mm(t3, t1, al, m, false, false);             // alpha g alpha
        for (int i = threadIdx.x; i < mm2; i += blockDim.x) {
            s[i].x += t2[i].x; s[i].y += t2[i].y;
            e[i].x += t2[i].x + g[i].x; e[i].y += t2[i].y + g[i].y;
            al[i] = t3[i];
        }
        __syncthreads();
        ++it;
        if (it >= 100) {                             // selfenergy.py:127-130
            if (threadIdx.x == 0) *notconv = 1;
            break;
        }
    }
    inv_shift(g, s, z, m, aug, red, ired, bad);
    return it;
}

__global__ void __launch_bounds__(ST) k_sig(const SigArgs a) {
    extern __shared__ double smraw[];
    const int m = a.m, mm2 = m * m;
    cplx *s = reinterpret_cast<cplx *>(a.gws ? a.gws + (size_t)blockIdx.x * a.gws_stride : smraw), *e = s + mm2, *al = e + mm2, *g = al + mm2, *t1 = g + mm2, *t2 = t1 + mm2, *t3 = t2 + mm2;
    cplx *sl = t3 + mm2, *sr = sl + mm2;
    cplx *aug = sr + mm2;                        // m x 2m
    double *red = a.gws ? smraw : reinterpret_cast<double *>(aug + 2 * mm2);   // >= max(64, 2m), always in shared memory
    __shared__ int ired, bad, notconv;
    for (int iw = blockIdx.x; iw < a.nw; iw += gridDim.x) {
        if (threadIdx.x == 0) { bad = 0; notconv = 0; }
        __syncthreads();
        const double w = a.omegas[iw];
        const cplx z = {w * w - a.eta * a.eta, 2.0 * w * a.eta};
        if (a.mode == 0 || a.mode == 3) {
            const int it = sgf(a, a.dirR, z, s, e, al, g, t1, t2, t3, aug, red, &ired, &bad, &notconv);
            if (a.mode == 3) {            // sig.sgf: the surface Green function itself
                for (int i = threadIdx.x; i < mm2; i += blockDim.x) {
                    a.se_out[((size_t)iw * mm2 + i) * 2] = g[i].x;
                    a.se_out[((size_t)iw * mm2 + i) * 2 + 1] = g[i].y;
                }
                if (threadIdx.x == 0) { a.iters[iw] = it; a.status[iw] = notconv ? 1 : (bad ? 2 : 0); }
                __syncthreads();
                continue;
            }
            // Sigma_R = K01 g K10 ; Sigma_L = K10 g K01 (selfenergy.py:133-140)
            const double *A = a.dirR ? a.K01 : a.K10, *B = a.dirR ? a.K10 : a.K01;
            for (int i = threadIdx.x; i < mm2; i += blockDim.x) { t1[i] = {A[i], 0.0}; t2[i] = {B[i], 0.0}; }
            __syncthreads();
            mm(t3, t1, g, m, false, false);
            mm(sl, t3, t2, m, false, false);
            for (int i = threadIdx.x; i < mm2; i += blockDim.x) {
                a.se_out[((size_t)iw * mm2 + i) * 2] = sl[i].x;
                a.se_out[((size_t)iw * mm2 + i) * 2 + 1] = sl[i].y;
            }
            if (threadIdx.x == 0) { a.iters[iw] = it; a.status[iw] = notconv ? 1 : (bad ? 2 : 0); }
        } else {
            for (int d = 0; d < 2; ++d) {    // d = 0: 'L', 1: 'R'
                const int it = sgf(a, d == 1, z, s, e, al, g, t1, t2, t3, aug, red, &ired, &bad, &notconv);
                const double *A = d ? a.K01 : a.K10, *B = d ? a.K10 : a.K01;
                for (int i = threadIdx.x; i < mm2; i += blockDim.x) { t1[i] = {A[i], 0.0}; t2[i] = {B[i], 0.0}; }
                __syncthreads();
                mm(t3, t1, g, m, false, false);
                mm(d ? sr : sl, t3, t2, m, false, false);
                if (threadIdx.x == 0 && a.iters) a.iters[2 * iw + d] = it;
            }
            // G = inv((w + 1e-8 i)^2 - K00 - Sigma_L - Sigma_R)  (selfenergy.py:145-147)
            for (int i = threadIdx.x; i < mm2; i += blockDim.x) s[i] = {a.K00[i] + sl[i].x + sr[i].x, sl[i].y + sr[i].y};
            __syncthreads();
            const cplx z2 = {w * w - 1e-16, 2.0 * w * 1e-8};
            inv_shift(g, s, z2, m, aug, red, &ired, &bad);
            if (a.mode == 2) {            // sig.retargf: the Green function itself
                for (int i = threadIdx.x; i < mm2; i += blockDim.x) {
                    a.se_out[((size_t)iw * mm2 + i) * 2] = g[i].x;
                    a.se_out[((size_t)iw * mm2 + i) * 2 + 1] = g[i].y;
                }
                if (threadIdx.x == 0) a.status[iw] = notconv ? 1 : (bad ? 2 : 0);
                __syncthreads();
                continue;
            }
            // Gamma = -i (Sigma - Sigma^dagger):  Gamma_ij = -i (S_ij - conj(S_ji))
            for (int i = threadIdx.x; i < mm2; i += blockDim.x) {
                const int r = i / m, c = i % m;
                const cplx dl = {sl[i].x - sl[c * m + r].x, sl[i].y + sl[c * m + r].y};
                const cplx dr = {sr[i].x - sr[c * m + r].x, sr[i].y + sr[c * m + r].y};
                t1[i] = {dl.y, -dl.x};
                t2[i] = {dr.y, -dr.x};
            }
            __syncthreads();
            mm(t3, g, t1, m, false, false);            // G Gamma_L
            // (G Gamma_L) G^dagger : B^T with conjugation -> build conj(G) first
            for (int i = threadIdx.x; i < mm2; i += blockDim.x) e[i] = {g[i].x, -g[i].y};
            __syncthreads();
            mm(al, t3, e, m, false, true);             // . G^dagger
            double tr = 0.0;                           // Re Tr[X Gamma_R]
            for (int i = threadIdx.x; i < mm2; i += blockDim.x) {
                const int r = i / m, c = i % m;
                const cplx x = al[r * m + c], y = t2[c * m + r];
                tr += x.x * y.x - x.y * y.y;
            }
            tr = block_sum(tr, red);
            if (threadIdx.x == 0) { a.tm_out[iw] = tr; a.status[iw] = notconv ? 1 : (bad ? 2 : 0); }
        }
        __syncthreads();
    }
}

int run_sig(int device, int m, const double *K00, const double *K11, const double *K01, const double *K10, double eta, int mode, int dirR,
            const double *omegas, int nw, double *se_out, double *tm_out, int32_t *iters_out) {
    SCLMD_REQUIRE(m > 0 && K00 && K11 && K01 && K10 && omegas && nw > 0, "sig: bad arguments");
    if (int e = select_device(device)) return e;
    const size_t mat_doubles = (size_t)11 * m * m * 2, red_doubles = std::max(64, 2 * m) + 64;
    const bool in_smem = (mat_doubles + red_doubles) * sizeof(double) <= 220 * 1024;
    SCLMD_REQUIRE(m <= 512, "sig: m=%d exceeds the supported lead block size (512)", m);
    const size_t smem = ((in_smem ? mat_doubles : 0) + red_doubles) * sizeof(double);
    int grid = std::min(nw, 8 * sm_count(device));
    if (!in_smem) grid = std::min(nw, std::max(1, std::min(2 * sm_count(device), (int)(((size_t)2 << 30) / (mat_doubles * sizeof(double))))));
    DevBuf<double> k00, k11, k01, k10, dom, dse, dtm, gws;
    if (!in_smem) SCLMD_CUDA(gws.alloc(mat_doubles * grid));
    DevBuf<int> dit, dst;
    const size_t mm2 = (size_t)m * m;
    SCLMD_CUDA(k00.alloc(mm2)); SCLMD_CUDA(k11.alloc(mm2)); SCLMD_CUDA(k01.alloc(mm2)); SCLMD_CUDA(k10.alloc(mm2));
    SCLMD_CUDA(dom.alloc(nw)); SCLMD_CUDA(dit.alloc((size_t)2 * nw)); SCLMD_CUDA(dst.alloc(nw));
    if (mode != 1) SCLMD_CUDA(dse.alloc((size_t)nw * mm2 * 2));
    else SCLMD_CUDA(dtm.alloc(nw));
    SCLMD_CUDA(cudaMemcpy(k00.p, K00, mm2 * 8, cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(k11.p, K11, mm2 * 8, cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(k01.p, K01, mm2 * 8, cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(k10.p, K10, mm2 * 8, cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(dom.p, omegas, nw * 8, cudaMemcpyHostToDevice));
    SigArgs a{};
    a.m = m; a.nw = nw; a.mode = mode; a.K00 = k00.p; a.K11 = k11.p; a.K01 = k01.p; a.K10 = k10.p; a.eta = eta; a.dirR = dirR;
    a.omegas = dom.p; a.se_out = dse.p; a.tm_out = dtm.p; a.iters = dit.p; a.status = dst.p;
    a.gws = gws.p; a.gws_stride = mat_doubles;
    SCLMD_CUDA(cudaFuncSetAttribute(k_sig, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, (size_t)48 * 1024)));
    k_sig<<<grid, ST, smem>>>(a);
    SCLMD_CUDA(cudaGetLastError());
    SCLMD_CUDA(cudaDeviceSynchronize());
    std::vector<int> st(nw), it((size_t)2 * nw);
    SCLMD_CUDA(cudaMemcpy(st.data(), dst.p, nw * sizeof(int), cudaMemcpyDeviceToHost));
    SCLMD_CUDA(cudaMemcpy(it.data(), dit.p, (size_t)2 * nw * sizeof(int), cudaMemcpyDeviceToHost));
    if (mode != 1) {
        SCLMD_CUDA(cudaMemcpy(se_out, dse.p, (size_t)nw * mm2 * 2 * 8, cudaMemcpyDeviceToHost));
        if (iters_out && (mode == 0 || mode == 3)) memcpy(iters_out, it.data(), nw * sizeof(int));
    } else {
        SCLMD_CUDA(cudaMemcpy(tm_out, dtm.p, nw * 8, cudaMemcpyDeviceToHost));
    }
    for (int i = 0; i < nw; ++i) {
        if (st[i] == 1) {
            set_error("Iteration number exceeded 100, please increase eta (omega[%d]=%g)", i, omegas[i]);
            return SCLMD_ERR_NOCONV;
        }
        if (st[i] == 2) {
            set_error("sig: singular matrix at omega[%d]=%g", i, omegas[i]);
            return SCLMD_ERR_STATE;
        }
    }
    return SCLMD_OK;
}

}  // namespace

extern "C" {

int sclmd_sig_selfenergy(int device, int m, const double *K00, const double *K11, const double *K01, const double *K10, double eta,
                         char direction, const double *omegas, int nw, double *se_out, int32_t *iters_out) {
    SCLMD_REQUIRE(direction == 'R' || direction == 'L', "Wrong direction, should only be R or L");
    SCLMD_REQUIRE(se_out, "sclmd_sig_selfenergy: NULL output");
    return run_sig(device, m, K00, K11, K01, K10, eta, 0, direction == 'R', omegas, nw, se_out, nullptr, iters_out);
}

// sig.sgf (selfenergy.py:105-131): surface Green function inv((w + i eta)^2 - s) after the decimation, sgf_out[nw][m][m] complex
int sclmd_sig_sgf(int device, int m, const double *K00, const double *K11, const double *K01, const double *K10, double eta,
                  char direction, const double *omegas, int nw, double *sgf_out, int32_t *iters_out) {
    SCLMD_REQUIRE(direction == 'R' || direction == 'L', "Wrong direction, should only be R or L");
    SCLMD_REQUIRE(sgf_out, "sclmd_sig_sgf: NULL output");
    return run_sig(device, m, K00, K11, K01, K10, eta, 3, direction == 'R', omegas, nw, sgf_out, nullptr, iters_out);
}

// sig.retargf (selfenergy.py:145-147): G(w) = inv((w + 1e-8 i)^2 - K00 - Sigma_L - Sigma_R), green_out[nw][m][m] complex (interleaved)
int sclmd_sig_green(int device, int m, const double *K00, const double *K11, const double *K01, const double *K10, double eta,
                    const double *omegas, int nw, double *green_out) {
    SCLMD_REQUIRE(green_out, "sclmd_sig_green: NULL output");
    return run_sig(device, m, K00, K11, K01, K10, eta, 2, 0, omegas, nw, green_out, nullptr, nullptr);
}

int sclmd_sig_tm(int device, int m, const double *K00, const double *K11, const double *K01, const double *K10, double eta,
                 const double *omegas, int nw, double *tm_out) {
    SCLMD_REQUIRE(tm_out, "sclmd_sig_tm: NULL output");
    return run_sig(device, m, K00, K11, K01, K10, eta, 1, 0, omegas, nw, nullptr, tm_out, nullptr);
}

}  // extern "C"
