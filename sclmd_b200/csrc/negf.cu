// Ballistic phonon transmission sweeps (replaces bpt.retargf / bpt.tm / bpt.ps / bpt.gettm,
// sclmd/negf.py:104-119,153-157,206-208,228-242).
//
// For every frequency one CTA factorises  M(w) = (w + 1e-9 i)^2 I - K - Sigma_L(w) - Sigma_R(w)
// (Sigma = -i w/damp on the bath dofs, negf.py:153-157) by blocked LU with partial pivoting and
// carries the right-hand sides along as extra columns, so the trailing update also performs the
// forward substitution.  Only the rows of G = M^-1 that the observable needs are back-substituted:
//   tm :  T = Re Tr[G Gamma_L G^dagger Gamma_R] = (2w/damp)^2 sum_{i in R, j in L} |G_ij|^2
//         (Gamma = -i(Sigma - Sigma^dagger) is diagonal, negf.py:214-215)  -> columns L, rows R
//   ps :  -2 w^2 n_B Tr Im G[sel,sel]  (negf.py:232)                       -> columns sel, rows sel
// The reference forms two full inverses and three dense n^3 products per frequency instead.
//
// Layout: augmented matrix W = [M | E] column-major, planar complex (real plane, imaginary
// plane), one scratch slot per resident CTA; CTAs loop over frequencies (persistent grid).
#include <algorithm>
#include <memory>

#include "common.cuh"

using namespace sclmd;

namespace {

constexpr int NB = 16;    // panel width
constexpr int CW = 64;    // column chunk of the trailing update
constexpr int NT = 256;   // threads per CTA

struct LuArgs {
    int n, ld, nrhs, ncols, nw, mode;   // mode 0 = tm, 1 = ps
    const double *K;                     // [n][n] row-major (symmetric)
    const double *sig_mask;              // [n] number of leads touching each dof (0/1/2)
    const int *rhs;                      // [nrhs] unit-vector index of every right-hand side
    const int *rows;                     // tm: [nrows] rows R ; ps: unused (rows == rhs)
    int nrows, row_stop;
    double damp, eps;                    // eps = 1e-9 broadening
    const double *omegas;                // [nw]
    const double *weight;                // ps: n_B(w) per frequency; tm: unused
    double *W;                           // [grid][2][ncols*ld]
    double *out;                         // [nw]
    int *status;                         // [nw] 0 ok, 1 singular pivot
};

__device__ __forceinline__ void cfma_sub(double &cr, double &ci, double ar, double ai, double br, double bi) {
    // c -= a*b
    cr = fma(-ar, br, cr); cr = fma(ai, bi, cr);
    ci = fma(-ar, bi, ci); ci = fma(-ai, br, ci);
}

// C[rows r0..r1) x cols [c0, c0+cw)  -=  P[rows][0..kb) . U[0..kb)[cols]
//   P planar in smem with leading dim pld (row index relative to prow0), U planar in smem [kb][CW]
__device__ __forceinline__ void rank_update(double *__restrict__ Wre, double *__restrict__ Wim, int ld, int r0, int r1, int c0, int cw,
                                            const double *__restrict__ Pre, const double *__restrict__ Pim, int pld, int prow0,
                                            const double *__restrict__ Ure, const double *__restrict__ Uim, int kb) {
    const int rg = threadIdx.x % 16, cg = threadIdx.x / 16;   // 16 row groups x 16 column groups, 4x4 tiles
    const int cbase = cg * 4;
    if (cbase >= cw) return;
    for (int rb = r0 + rg * 4; rb < r1; rb += 64) {
        double ar[4][4], ai[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) ar[i][j] = ai[i][j] = 0.0;
        for (int kk = 0; kk < kb; ++kk) {
            double pr[4], pi[4], ur[4], ui[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = min(rb + i, r1 - 1) - prow0;
                pr[i] = Pre[kk * pld + r];
                pi[i] = Pim[kk * pld + r];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ur[j] = Ure[kk * CW + cbase + j];
                ui[j] = Uim[kk * CW + cbase + j];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) cfma_sub(ar[i][j], ai[i][j], pr[i], pi[i], ur[j], ui[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (cbase + j >= cw) continue;
            const size_t col = (size_t)(c0 + cbase + j) * ld;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (rb + i >= r1) continue;
                Wre[col + rb + i] += ar[i][j];
                Wim[col + rb + i] += ai[i][j];
            }
        }
    }
}

__global__ void __launch_bounds__(NT) k_bpt_lu(const LuArgs a) {
    extern __shared__ double sm[];
    const int n = a.n, ld = a.ld, ncols = a.ncols;
    double *Pre = sm, *Pim = Pre + (size_t)NB * ld;           // panel [NB][ld] (column kk contiguous over rows)
    double *Ure = Pim + (size_t)NB * ld, *Uim = Ure + NB * CW;  // chunk [NB][CW]
    double *red = Uim + NB * CW;                                // [64]
    __shared__ int piv[NB];
    __shared__ int s_arg;
    __shared__ int s_bad;
    double *Wre = a.W + (size_t)blockIdx.x * 2 * (size_t)ncols * ld, *Wim = Wre + (size_t)ncols * ld;

    for (int iw = blockIdx.x; iw < a.nw; iw += gridDim.x) {
        const double w = a.omegas[iw];
        // (w + i eps)^2 = w^2 - eps^2 + 2 i w eps ;  Sigma_ii = -i w/damp * mask_i  ->  M_ii += i w/damp * mask_i
        const double zr = w * w - a.eps * a.eps, zi = 2.0 * w * a.eps, sg = w / a.damp;
        __syncthreads();
        if (threadIdx.x == 0) s_bad = 0;
        for (size_t e = threadIdx.x; e < (size_t)n * n; e += NT) {
            const int i = (int)(e % n), j = (int)(e / n);
            Wre[(size_t)j * ld + i] = (i == j ? zr : 0.0) - a.K[(size_t)i * n + j];
            Wim[(size_t)j * ld + i] = i == j ? zi + sg * a.sig_mask[i] : 0.0;
        }
        for (size_t e = threadIdx.x; e < (size_t)a.nrhs * n; e += NT) {
            const int i = (int)(e % n), c = (int)(e / n);
            Wre[(size_t)(n + c) * ld + i] = a.rhs[c] == i ? 1.0 : 0.0;
            Wim[(size_t)(n + c) * ld + i] = 0.0;
        }
        __syncthreads();

        // ---------------- blocked LU, right-hand sides carried as columns n..ncols
        for (int k0 = 0; k0 < n; k0 += NB) {
            const int kb = min(NB, n - k0), m = n - k0;
            for (int e = threadIdx.x; e < kb * m; e += NT) {
                const int kk = e / m, i = e % m;
                Pre[kk * ld + i] = Wre[(size_t)(k0 + kk) * ld + k0 + i];
                Pim[kk * ld + i] = Wim[(size_t)(k0 + kk) * ld + k0 + i];
            }
            __syncthreads();
            for (int j = 0; j < kb; ++j) {
                // pivot: max |re|+|im| over rows j..m (izamax convention)
                double best = -1.0;
                int arg = j;
                for (int i = j + threadIdx.x; i < m; i += NT) {
                    const double v = fabs(Pre[j * ld + i]) + fabs(Pim[j * ld + i]);
                    if (v > best) { best = v; arg = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
                }
                if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = best; red[32 + (threadIdx.x >> 5)] = (double)arg; }
                __syncthreads();
                if (threadIdx.x == 0) {
                    double b = red[0];
                    int ar = (int)red[32];
                    for (int q = 1; q < NT / 32; ++q)
                        if (red[q] > b || (red[q] == b && (int)red[32 + q] < ar)) { b = red[q]; ar = (int)red[32 + q]; }
                    piv[j] = ar;
                    s_arg = ar;
                    if (!(b > 0.0)) s_bad = 1;
                }
                __syncthreads();
                const int r = s_arg;
                if (r != j && threadIdx.x < kb) {
                    const int kk = threadIdx.x;
                    double t = Pre[kk * ld + j]; Pre[kk * ld + j] = Pre[kk * ld + r]; Pre[kk * ld + r] = t;
                    t = Pim[kk * ld + j]; Pim[kk * ld + j] = Pim[kk * ld + r]; Pim[kk * ld + r] = t;
                }
                __syncthreads();
                const double dr = Pre[j * ld + j], di = Pim[j * ld + j];
                const double dn = dr * dr + di * di;
                const double ir = dr / dn, ii = -di / dn;    // 1/pivot
                __syncthreads();
                for (int i = j + 1 + threadIdx.x; i < m; i += NT) {
                    const double xr = Pre[j * ld + i], xi = Pim[j * ld + i];
                    Pre[j * ld + i] = xr * ir - xi * ii;
                    Pim[j * ld + i] = xr * ii + xi * ir;
                }
                __syncthreads();
                const int rem = kb - 1 - j, below = m - 1 - j;
                for (int e = threadIdx.x; e < rem * below; e += NT) {
                    const int jj = j + 1 + e / below, i = j + 1 + e % below;
                    double cr = Pre[jj * ld + i], ci = Pim[jj * ld + i];
                    cfma_sub(cr, ci, Pre[j * ld + i], Pim[j * ld + i], Pre[jj * ld + j], Pim[jj * ld + j]);
                    Pre[jj * ld + i] = cr;
                    Pim[jj * ld + i] = ci;
                }
                __syncthreads();
            }
            // panel back to W (U11 / L11 / L21 needed by the back substitution and nothing else)
            for (int e = threadIdx.x; e < kb * m; e += NT) {
                const int kk = e / m, i = e % m;
                Wre[(size_t)(k0 + kk) * ld + k0 + i] = Pre[kk * ld + i];
                Wim[(size_t)(k0 + kk) * ld + k0 + i] = Pim[kk * ld + i];
            }
            // remaining columns in chunks: row swaps, U12 = L11^-1 A12, then A22 -= L21 U12
            for (int c0 = k0 + kb; c0 < ncols; c0 += CW) {
                const int cw = min(CW, ncols - c0);
                __syncthreads();
                if (threadIdx.x < cw) {
                    const size_t col = (size_t)(c0 + threadIdx.x) * ld + k0;
                    for (int jj = 0; jj < kb; ++jj) {
                        const int r = piv[jj];
                        if (r != jj) {
                            double t = Wre[col + jj]; Wre[col + jj] = Wre[col + r]; Wre[col + r] = t;
                            t = Wim[col + jj]; Wim[col + jj] = Wim[col + r]; Wim[col + r] = t;
                        }
                    }
                    double ur[NB], ui[NB];
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) {
                        ur[jj] = jj < kb ? Wre[col + jj] : 0.0;
                        ui[jj] = jj < kb ? Wim[col + jj] : 0.0;
                    }
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) {
#pragma unroll
                        for (int i2 = 0; i2 < NB; ++i2)
                            if (i2 > jj && i2 < kb && jj < kb) cfma_sub(ur[i2], ui[i2], Pre[jj * ld + i2], Pim[jj * ld + i2], ur[jj], ui[jj]);
                    }
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) {
                        if (jj < kb) {
                            Wre[col + jj] = ur[jj];
                            Wim[col + jj] = ui[jj];
                        }
                        Ure[jj * CW + threadIdx.x] = ur[jj];
                        Uim[jj * CW + threadIdx.x] = ui[jj];
                    }
                }
                __syncthreads();
                rank_update(Wre, Wim, ld, k0 + kb, n, c0, cw, Pre, Pim, ld, k0, Ure, Uim, kb);
            }
            __syncthreads();
        }

        // ---------------- back substitution on the right-hand sides, bottom block first, down to row_stop
        const int last = ((n - 1) / NB) * NB;
        for (int k0 = last; k0 >= 0 && k0 + NB > a.row_stop; k0 -= NB) {
            const int kb = min(NB, n - k0);
            // U[0:k0+kb, k0:k0+kb] -> panel buffer (rows relative to 0)
            const int rows = k0 + kb;
            for (int e = threadIdx.x; e < kb * rows; e += NT) {
                const int kk = e / rows, i = e % rows;
                Pre[kk * ld + i] = Wre[(size_t)(k0 + kk) * ld + i];
                Pim[kk * ld + i] = Wim[(size_t)(k0 + kk) * ld + i];
            }
            const int rlo = max(0, (a.row_stop / NB) * NB);
            for (int c0 = n; c0 < ncols; c0 += CW) {
                const int cw = min(CW, ncols - c0);
                __syncthreads();
                if (threadIdx.x < cw) {
                    const size_t col = (size_t)(c0 + threadIdx.x) * ld + k0;
                    double xr[NB], xi[NB];
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) {
                        xr[jj] = jj < kb ? Wre[col + jj] : 0.0;
                        xi[jj] = jj < kb ? Wim[col + jj] : 0.0;
                    }
#pragma unroll
                    for (int jj = NB - 1; jj >= 0; --jj) {
                        if (jj < kb) {
                            const double dr = Pre[jj * ld + k0 + jj], di = Pim[jj * ld + k0 + jj];
                            const double dn = dr * dr + di * di;
                            const double tr = (xr[jj] * dr + xi[jj] * di) / dn, ti = (xi[jj] * dr - xr[jj] * di) / dn;
                            xr[jj] = tr; xi[jj] = ti;
#pragma unroll
                            for (int i2 = 0; i2 < NB; ++i2)
                                if (i2 < jj) cfma_sub(xr[i2], xi[i2], Pre[jj * ld + k0 + i2], Pim[jj * ld + k0 + i2], tr, ti);
                        }
                    }
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) {
                        if (jj < kb) {
                            Wre[col + jj] = xr[jj];
                            Wim[col + jj] = xi[jj];
                        }
                        Ure[jj * CW + threadIdx.x] = xr[jj];
                        Uim[jj * CW + threadIdx.x] = xi[jj];
                    }
                }
                __syncthreads();
                if (k0 > rlo) rank_update(Wre, Wim, ld, rlo, k0, c0, cw, Pre, Pim, ld, 0, Ure, Uim, kb);
            }
            __syncthreads();
        }

        // ---------------- observable
        double acc = 0.0;
        if (a.mode == 0) {
            for (int e = threadIdx.x; e < a.nrows * a.nrhs; e += NT) {
                const int i = a.rows[e % a.nrows], c = e / a.nrows;
                const double xr = Wre[(size_t)(n + c) * ld + i], xi = Wim[(size_t)(n + c) * ld + i];
                acc += xr * xr + xi * xi;
            }
        } else {
            for (int c = threadIdx.x; c < a.nrhs; c += NT) acc += Wim[(size_t)(n + c) * ld + a.rhs[c]];
        }
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) {
            const double gam = 2.0 * w / a.damp;
            a.out[iw] = a.mode == 0 ? gam * gam * acc : -2.0 * w * w * a.weight[iw] * acc;
            a.status[iw] = s_bad;
        }
        __syncthreads();
    }
}

int run_lu(int device, int n, const double *K, const int32_t *idxL, int nL, const int32_t *idxR, int nR, double damp,
           const double *omegas, int nw, int mode, const double *weight, const int32_t *sel, int nsel, double *out) {
    SCLMD_REQUIRE(n > 0 && K && idxL && idxR && nL > 0 && nR > 0 && omegas && nw > 0 && out && damp != 0.0, "bpt: bad arguments");
    if (int e = select_device(device)) return e;
    std::vector<double> mask(n, 0.0);
    for (int i = 0; i < nL; ++i) {
        SCLMD_REQUIRE(idxL[i] >= 0 && idxL[i] < n, "bpt: left bath dof %d out of range", idxL[i]);
        mask[idxL[i]] += 1.0;
    }
    for (int i = 0; i < nR; ++i) {
        SCLMD_REQUIRE(idxR[i] >= 0 && idxR[i] < n, "bpt: right bath dof %d out of range", idxR[i]);
        mask[idxR[i]] += 1.0;
    }
    std::vector<int> rhs, rows;
    int row_stop = 0;
    if (mode == 0) {
        rhs.assign(idxL, idxL + nL);
        rows.assign(idxR, idxR + nR);
        row_stop = *std::min_element(rows.begin(), rows.end());
    } else {
        SCLMD_REQUIRE(sel && nsel > 0 && weight, "bpt.ps: empty selection");
        for (int i = 0; i < nsel; ++i) SCLMD_REQUIRE(sel[i] >= 0 && sel[i] < n, "bpt.ps: selected dof %d out of range", sel[i]);
        rhs.assign(sel, sel + nsel);
        rows = rhs;
        row_stop = *std::min_element(rows.begin(), rows.end());
    }
    LuArgs a{};
    a.n = n; a.ld = round_up(n, 4); a.nrhs = (int)rhs.size(); a.ncols = n + a.nrhs; a.nw = nw; a.mode = mode;
    a.nrows = (int)rows.size(); a.row_stop = row_stop; a.damp = damp; a.eps = 1e-9;
    const size_t smem = ((size_t)2 * NB * a.ld + 2 * NB * CW + 64) * sizeof(double);
    SCLMD_REQUIRE(smem <= 220 * 1024, "bpt: n=%d too large for the shared-memory panel (max ~850)", n);
    const int grid = std::min(nw, 2 * sm_count(device));
    DevBuf<double> dK, dmask, dom, dwt, W, dout;
    DevBuf<int> drhs, drows, dstat;
    SCLMD_CUDA(dK.alloc((size_t)n * n)); SCLMD_CUDA(dmask.alloc(n)); SCLMD_CUDA(dom.alloc(nw)); SCLMD_CUDA(dwt.alloc(nw));
    SCLMD_CUDA(W.alloc((size_t)grid * 2 * a.ncols * a.ld)); SCLMD_CUDA(dout.alloc(nw));
    SCLMD_CUDA(drhs.alloc(rhs.size())); SCLMD_CUDA(drows.alloc(rows.size())); SCLMD_CUDA(dstat.alloc(nw));
    SCLMD_CUDA(cudaMemcpy(dK.p, K, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(dmask.p, mask.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(dom.p, omegas, nw * sizeof(double), cudaMemcpyHostToDevice));
    if (weight) SCLMD_CUDA(cudaMemcpy(dwt.p, weight, nw * sizeof(double), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(drhs.p, rhs.data(), rhs.size() * sizeof(int), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(drows.p, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice));
    a.K = dK.p; a.sig_mask = dmask.p; a.rhs = drhs.p; a.rows = drows.p; a.omegas = dom.p; a.weight = dwt.p;
    a.W = W.p; a.out = dout.p; a.status = dstat.p;
    SCLMD_CUDA(cudaFuncSetAttribute(k_bpt_lu, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bpt_lu<<<grid, NT, smem>>>(a);
    SCLMD_CUDA(cudaGetLastError());
    SCLMD_CUDA(cudaDeviceSynchronize());
    SCLMD_CUDA(cudaMemcpy(out, dout.p, nw * sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<int> st(nw);
    SCLMD_CUDA(cudaMemcpy(st.data(), dstat.p, nw * sizeof(int), cudaMemcpyDeviceToHost));
    for (int i = 0; i < nw; ++i)
        if (st[i]) {
            set_error("bpt: singular matrix at omega[%d]=%g (numpy.linalg.LinAlgError in the reference)", i, omegas[i]);
            return SCLMD_ERR_STATE;
        }
    return SCLMD_OK;
}

}  // namespace

extern "C" {

int sclmd_bpt_tm(int device, int n, const double *K, const int32_t *idxL, int nL, const int32_t *idxR, int nR, double damp,
                 const double *omegas, int nw, double *tm_out) {
    return run_lu(device, n, K, idxL, nL, idxR, nR, damp, omegas, nw, 0, nullptr, nullptr, 0, tm_out);
}

int sclmd_bpt_ps(int device, int n, const double *K, const int32_t *idxL, int nL, const int32_t *idxR, int nR, double damp,
                 const double *omegas, const double *nb, int nw, const int32_t *sel, int nsel, double *ps_out) {
    return run_lu(device, n, K, idxL, nL, idxR, nR, damp, omegas, nw, 1, nb, sel, nsel, ps_out);
}

}  // extern "C"
