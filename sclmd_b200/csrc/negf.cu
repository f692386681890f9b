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
#include <cstdlib>
#include <memory>

#include "common.cuh"

using namespace sclmd;

namespace {

constexpr int NB = 16;    // panel width
constexpr int CW = 64;    // column chunk of the trailing update
constexpr int NT = 256;   // threads per CTA
constexpr int UW = CW + 4; // padded row of the U chunk in shared memory (conflict-free DMMA fragment loads)

struct LuArgs {
    int n, np, ld, lw, nrhs, ncols, nw, mode;   // mode 0 = tm, 1 = ps; ld: panel leading dim (smem), lw: leading dim of W
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
    // optional dense self-energy block of a biased electron bath on dofs [b0, b0+nb) (negf.py:162-190):
    //   Sigma_b^r = -i w bdamp - bias chiminus ;  Sigma^K_b = kr1 bdamp + kr2 chiplus + i ki chiminus (per-frequency scalars)
    int b0, nb;
    const double *bdamp, *chiplus, *chiminus;   // [nb][nb]
    double bias;
    const double *kd, *kr1, *kr2, *ki;          // [nw] Keldysh weights (mode 2)
    double *Xs;                                 // [grid][2][np][nrhs] first-pass solution (mode 2)
    long long *timing;                   // optional [8] cycle counters of CTA 0 (build, panel, trsm, update, backsub, observable)
};

__device__ __forceinline__ void cfma_sub(double &cr, double &ci, double ar, double ai, double br, double bi) {
    // c -= a*b
    cr = fma(-ar, br, cr); cr = fma(ai, bi, cr);
    ci = fma(-ar, bi, ci); ci = fma(-ai, br, ci);
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// C[rows r0..r1) x cols [c0, c0+cw)  -=  P[rows][0..NB) . U[0..NB)[cols]      (complex, planar; W row-major, leading dim lw)
//   P in smem: Pre/Pim[kk*pld + (row - prow0)], columns kk >= kb are zero;  U in smem: Ure/Uim[kk*UW + col], rows >= kb zero.
// FP64 tensor path: the complex product is four real DMMA.8x8x4 per fragment pair
//   Re += Pre.Ure + (-Pim).Uim ,  Im += Pre.Uim + Pim.Ure.
// 8 warps = 2 row groups (32 rows) x 4 column groups (16 columns): one pass covers 64 rows x 64 columns.
// The C tile is fetched into registers BEFORE the DMMA loop (it does not depend on it): every lane has 16 independent
// 16-byte loads per plane in flight while the tensor pipe works, so the update streams W at memory speed.
__device__ __forceinline__ void rank_update(double *__restrict__ Wre, double *__restrict__ Wim, int lw, int r0, int r1, int c0, int cw,
                                            const double *__restrict__ Pre, const double *__restrict__ Pim, int pld, int prow0,
                                            const double *__restrict__ Ure, const double *__restrict__ Uim) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp & 1, wc = warp >> 1;
    const int fr = lane >> 2, fk = lane & 3;
    if (wc * 16 >= cw) return;
    for (int rb = r0 + wr * 32; rb < r1; rb += 64) {
        double2 cre[4][2], cim[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = min(rb + i * 8 + fr, r1 - 1);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int col = min(wc * 16 + j * 8 + 2 * fk, CW - 2);      // c0, lw even -> 16-byte aligned pair
                const size_t o = (size_t)row * lw + c0 + col;
                cre[i][j] = *reinterpret_cast<const double2 *>(Wre + o);
                cim[i][j] = *reinterpret_cast<const double2 *>(Wim + o);
            }
        }
        double are[4][2][2], aim[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) are[i][j][0] = are[i][j][1] = aim[i][j][0] = aim[i][j][1] = 0.0;
#pragma unroll
        for (int kk = 0; kk < NB; kk += 4) {
            double pr[4], pi[4], pn[4], ur[2], ui[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = min(rb + i * 8 + fr, r1 - 1) - prow0;
                pr[i] = Pre[(kk + fk) * pld + r];
                pi[i] = Pim[(kk + fk) * pld + r];
                pn[i] = -pi[i];
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                ur[j] = Ure[(kk + fk) * UW + wc * 16 + j * 8 + fr];
                ui[j] = Uim[(kk + fk) * UW + wc * 16 + j * 8 + fr];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    dmma(are[i][j][0], are[i][j][1], pr[i], ur[j]);
                    dmma(are[i][j][0], are[i][j][1], pn[i], ui[j]);
                    dmma(aim[i][j][0], aim[i][j][1], pr[i], ui[j]);
                    dmma(aim[i][j][0], aim[i][j][1], pi[i], ur[j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = rb + i * 8 + fr;
            if (row >= r1) continue;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int col = wc * 16 + j * 8 + 2 * fk;
                if (col >= cw) continue;
                const size_t o = (size_t)row * lw + c0 + col;
                if (col + 1 < cw) {
                    *reinterpret_cast<double2 *>(Wre + o) = make_double2(cre[i][j].x - are[i][j][0], cre[i][j].y - are[i][j][1]);
                    *reinterpret_cast<double2 *>(Wim + o) = make_double2(cim[i][j].x - aim[i][j][0], cim[i][j].y - aim[i][j][1]);
                } else {
                    Wre[o] = cre[i][j].x - are[i][j][0];
                    Wim[o] = cim[i][j].x - aim[i][j][0];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(NT) k_bpt_lu(const LuArgs a) {
    extern __shared__ double sm[];
    // working dimension n is even (a decoupled identity dof pads an odd system): every chunk start k0+kb and the first
    // right-hand-side column are then even, which keeps the 16-byte tile accesses of the update aligned
    const int nl = a.n, n = a.np, ld = a.ld, lw = a.lw, ncols = a.ncols;
    double *Pre = sm, *Pim = Pre + (size_t)NB * ld;           // panel [NB][ld] (column kk contiguous over rows)
    double *Ure = Pim + (size_t)NB * ld, *Uim = Ure + NB * UW;  // chunk [NB][UW]
    double *red = Uim + NB * UW;                                // [64]
    __shared__ int piv[NB];
    __shared__ int sw_dst[2 * NB], sw_src[2 * NB], sw_n;   // net effect of the panel's row interchanges on the touched rows
    __shared__ int s_bad;
    // W = [M | E] row-major, planar: row swaps, the U12 solve and the tile traffic of the update are all coalesced
    double *Wre = a.W + (size_t)blockIdx.x * 2 * (size_t)n * lw, *Wim = Wre + (size_t)n * lw;

    long long tacc[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64();
#define TICK(k) do { if (a.timing && blockIdx.x == 0 && threadIdx.x == 0) { const long long _n = clock64(); tacc[k] += _n - tlast; tlast = _n; } } while (0)
    for (int iw = blockIdx.x; iw < a.nw; iw += gridDim.x) {
        const double w = a.omegas[iw];
        TICK(5);
        // (w + i eps)^2 = w^2 - eps^2 + 2 i w eps ;  Sigma_ii = -i w/damp * mask_i  ->  M_ii += i w/damp * mask_i
        const double zr = w * w - a.eps * a.eps, zi = 2.0 * w * a.eps, sg = w / a.damp;
        __syncthreads();
        if (threadIdx.x == 0) s_bad = 0;
        // mode 2 (biased power spectrum, negf.py:236) needs G^a[:,sel] and G^r[sel,:]: two factorisations per frequency,
        //   pass 0: M^a = z^2 - K - Sigma^r-dagger            (advanced; negf.py:210-212 keeps the +i eps of z)
        //   pass 1: (M^r)^T                                   (rows of G^r are columns of its transpose)
        const int npass = a.mode == 2 ? 2 : 1;
        for (int pass = 0; pass < npass; ++pass) {
        const double sgn = (a.mode == 2 && pass == 0) ? -1.0 : 1.0;
        const bool tblk = a.mode == 2;
        for (int i = threadIdx.x >> 5; i < n; i += NT / 32) {          // warp per row, lanes over columns: coalesced, 4-deep ILP
            const double *krow = a.K + (size_t)min(i, nl - 1) * nl;
            double *wre = Wre + (size_t)i * lw, *wim = Wim + (size_t)i * lw;
            for (int j0 = threadIdx.x & 31; j0 < ncols; j0 += 128) {
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u;
                    v[u] = (j < nl && i < nl) ? krow[j] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u;
                    if (j < n) {
                        double mr = (i == j ? (i < nl ? zr : 1.0) : 0.0) - v[u];
                        double mi = (i == j && i < nl) ? zi + sgn * sg * a.sig_mask[i] : 0.0;
                        if (a.nb > 0 && i >= a.b0 && i < a.b0 + a.nb && j >= a.b0 && j < a.b0 + a.nb) {
                            const int bi = tblk ? j - a.b0 : i - a.b0, bj = tblk ? i - a.b0 : j - a.b0;
                            mr += a.bias * a.chiminus[bi * a.nb + bj];          // M -= Sigma_b :  +bias chi-  and  +i w bdamp
                            mi += sgn * w * a.bdamp[bi * a.nb + bj];
                        }
                        wre[j] = mr;
                        wim[j] = mi;
                    } else if (j < ncols) {
                        wre[j] = a.rhs[j - n] == i ? 1.0 : 0.0;
                        wim[j] = 0.0;
                    }
                }
            }
        }
        __syncthreads();

        TICK(0);
        // ---------------- blocked LU, right-hand sides carried as columns n..ncols
        for (int k0 = 0; k0 < n; k0 += NB) {
            const int kb = min(NB, n - k0), m = n - k0;
            for (int e = threadIdx.x; e < NB * m; e += NT) {
                const int kk = e % NB, i = e / NB;
                const bool ok = kk < kb;
                Pre[kk * ld + i] = ok ? Wre[(size_t)(k0 + i) * lw + k0 + kk] : 0.0;
                Pim[kk * ld + i] = ok ? Wim[(size_t)(k0 + i) * lw + k0 + kk] : 0.0;
            }
            __syncthreads();
            for (int j = 0; j < kb; ++j) {
                // pivot: max |re|+|im| over rows j..m (izamax convention), ties -> smallest row
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
                double *cand = red + (j & 1) * 32;       // double-buffered: no barrier needed before the next column's writes
                if ((threadIdx.x & 31) == 0) { cand[threadIdx.x >> 5] = best; cand[16 + (threadIdx.x >> 5)] = (double)arg; }
                __syncthreads();
                best = cand[0];
                arg = (int)cand[16];
#pragma unroll
                for (int q = 1; q < NT / 32; ++q)
                    if (cand[q] > best || (cand[q] == best && (int)cand[16 + q] < arg)) { best = cand[q]; arg = (int)cand[16 + q]; }
                const int r = arg;                        // every thread derives the same pivot
                if (threadIdx.x == 0) {
                    piv[j] = r;
                    if (!(best > 0.0)) s_bad = 1;
                }
                if (r != j && threadIdx.x < NB) {         // swap rows j <-> r of the panel (all NB columns)
                    const int kk = threadIdx.x;
                    double t = Pre[kk * ld + j]; Pre[kk * ld + j] = Pre[kk * ld + r]; Pre[kk * ld + r] = t;
                    t = Pim[kk * ld + j]; Pim[kk * ld + j] = Pim[kk * ld + r]; Pim[kk * ld + r] = t;
                }
                __syncthreads();
                const double dr = Pre[j * ld + j], di = Pim[j * ld + j];
                const double dn = dr * dr + di * di;
                const double ir = dr / dn, ii = -di / dn;    // 1/pivot
                // scale column j and apply the rank-1 update to the rest of the panel; a thread owns whole rows and
                // first gathers its row into registers so the shared-memory round trips overlap
                for (int i = j + 1 + threadIdx.x; i < m; i += NT) {
                    double vr[NB], vi[NB];
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) {
                        vr[jj] = Pre[jj * ld + i];
                        vi[jj] = Pim[jj * ld + i];
                    }
                    double lr = 0.0, li = 0.0;
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj)
                        if (jj == j) { lr = vr[jj] * ir - vi[jj] * ii; li = vr[jj] * ii + vi[jj] * ir; }
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) {
                        if (jj == j) {
                            Pre[jj * ld + i] = lr;
                            Pim[jj * ld + i] = li;
                        } else if (jj > j && jj < kb) {
                            cfma_sub(vr[jj], vi[jj], lr, li, Pre[jj * ld + j], Pim[jj * ld + j]);
                            Pre[jj * ld + i] = vr[jj];
                            Pim[jj * ld + i] = vi[jj];
                        }
                    }
                }
                __syncthreads();
            }
            for (int e = threadIdx.x; e < kb * m; e += NT) {
                const int kk = e % kb, i = e / kb;
                Wre[(size_t)(k0 + i) * lw + k0 + kk] = Pre[kk * ld + i];
                Wim[(size_t)(k0 + i) * lw + k0 + kk] = Pim[kk * ld + i];
            }
            __syncthreads();
            TICK(1);
            // net row permutation of this panel (LAPACK laswp without the 16 dependent round trips): touched rows are the
            // kb top rows plus the distinct pivot rows below them; sw_src[t] = original row that ends up in row sw_dst[t]
            if (threadIdx.x == 0) {
                int nt = kb;
                for (int t = 0; t < kb; ++t) sw_dst[t] = sw_src[t] = t;
                for (int jj = 0; jj < kb; ++jj) {
                    const int r = piv[jj];
                    int pos = -1;
                    for (int t = 0; t < nt; ++t) if (sw_dst[t] == r) { pos = t; break; }
                    if (pos < 0) { pos = nt++; sw_dst[pos] = sw_src[pos] = r; }
                    const int tmp = sw_src[jj]; sw_src[jj] = sw_src[pos]; sw_src[pos] = tmp;
                }
                sw_n = nt;
            }
            __syncthreads();
            // every remaining column (thread per column, coalesced across the warp): interchanges, then U12 = L11^-1 A12
            for (int c = k0 + kb + threadIdx.x; c < ncols; c += NT) {
                double *wre = Wre + (size_t)k0 * lw + c, *wim = Wim + (size_t)k0 * lw + c;
                const int nt = sw_n;
                double vr[2 * NB], vi[2 * NB];
#pragma unroll
                for (int t = 0; t < 2 * NB; ++t) {
                    if (t < nt) {
                        vr[t] = wre[(size_t)sw_src[t] * lw];
                        vi[t] = wim[(size_t)sw_src[t] * lw];
                    } else {
                        vr[t] = vi[t] = 0.0;
                    }
                }
#pragma unroll
                for (int t = 0; t < 2 * NB; ++t)     // displaced rows below the panel
                    if (t >= kb && t < nt && sw_dst[t] != sw_src[t]) { wre[(size_t)sw_dst[t] * lw] = vr[t]; wim[(size_t)sw_dst[t] * lw] = vi[t]; }
                double ur[NB], ui[NB];
#pragma unroll
                for (int jj = 0; jj < NB; ++jj) {
                    ur[jj] = jj < kb ? vr[jj] : 0.0;
                    ui[jj] = jj < kb ? vi[jj] : 0.0;
                }
#pragma unroll
                for (int jj = 0; jj < NB; ++jj) {
#pragma unroll
                    for (int i2 = 0; i2 < NB; ++i2)
                        if (i2 > jj && i2 < kb && jj < kb) cfma_sub(ur[i2], ui[i2], Pre[jj * ld + i2], Pim[jj * ld + i2], ur[jj], ui[jj]);
                }
#pragma unroll
                for (int jj = 0; jj < NB; ++jj)
                    if (jj < kb) {
                        wre[(size_t)jj * lw] = ur[jj];
                        wim[(size_t)jj * lw] = ui[jj];
                    }
            }
            __syncthreads();
            TICK(2);
            // trailing update in column chunks: A22 -= L21 U12 (also advances the carried right-hand sides)
            for (int c0 = k0 + kb; c0 < ncols; c0 += CW) {
                const int cw = min(CW, ncols - c0);
                for (int e = threadIdx.x; e < NB * CW; e += NT) {
                    const int cc = e % CW, jj = e / CW;
                    const bool ok = cc < cw && jj < kb;
                    Ure[jj * UW + cc] = ok ? Wre[(size_t)(k0 + jj) * lw + c0 + cc] : 0.0;
                    Uim[jj * UW + cc] = ok ? Wim[(size_t)(k0 + jj) * lw + c0 + cc] : 0.0;
                }
                __syncthreads();
                rank_update(Wre, Wim, lw, k0 + kb, n, c0, cw, Pre, Pim, ld, k0, Ure, Uim);
                __syncthreads();
            }
            TICK(3);
        }

        // ---------------- back substitution on the right-hand sides, bottom block first, down to row_stop
        const int last = ((n - 1) / NB) * NB;
        for (int k0 = last; k0 >= 0 && k0 + NB > a.row_stop; k0 -= NB) {
            const int kb = min(NB, n - k0);
            // U[rlo:k0+kb, k0:k0+kb] -> panel buffer (rows relative to 0; only the rows still needed)
            const int rows = k0 + kb;
            const int rlo = max(0, (a.row_stop / NB) * NB);
            const int nr = rows - rlo;
            for (int e = threadIdx.x; e < NB * nr; e += NT) {
                const int kk = e % NB, i = rlo + e / NB;
                const bool ok = kk < kb;
                Pre[kk * ld + i] = ok ? Wre[(size_t)i * lw + k0 + kk] : 0.0;
                Pim[kk * ld + i] = ok ? Wim[(size_t)i * lw + k0 + kk] : 0.0;
            }
            __syncthreads();
            for (int c = n + threadIdx.x; c < ncols; c += NT) {    // X1 = U11^-1 Y1, thread per right-hand side
                double *wre = Wre + (size_t)k0 * lw + c, *wim = Wim + (size_t)k0 * lw + c;
                double xr[NB], xi[NB];
#pragma unroll
                for (int jj = 0; jj < NB; ++jj) {
                    xr[jj] = jj < kb ? wre[(size_t)jj * lw] : 0.0;
                    xi[jj] = jj < kb ? wim[(size_t)jj * lw] : 0.0;
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
                for (int jj = 0; jj < NB; ++jj)
                    if (jj < kb) {
                        wre[(size_t)jj * lw] = xr[jj];
                        wim[(size_t)jj * lw] = xi[jj];
                    }
            }
            __syncthreads();
            if (k0 > rlo) {
                for (int c0 = n; c0 < ncols; c0 += CW) {           // Y_above -= U_above,blk X1
                    const int cw = min(CW, ncols - c0);
                    for (int e = threadIdx.x; e < NB * CW; e += NT) {
                        const int cc = e % CW, jj = e / CW;
                        const bool ok = cc < cw && jj < kb;
                        Ure[jj * UW + cc] = ok ? Wre[(size_t)(k0 + jj) * lw + c0 + cc] : 0.0;
                        Uim[jj * UW + cc] = ok ? Wim[(size_t)(k0 + jj) * lw + c0 + cc] : 0.0;
                    }
                    __syncthreads();
                    rank_update(Wre, Wim, lw, rlo, k0, c0, cw, Pre, Pim, ld, 0, Ure, Uim);
                    __syncthreads();
                }
            }
        }

        if (a.mode == 2 && pass == 0) {     // keep G^a[:,sel] while the second factorisation runs
            double *xr = a.Xs + (size_t)blockIdx.x * 2 * n * a.nrhs, *xi = xr + (size_t)n * a.nrhs;
            for (int e = threadIdx.x; e < n * a.nrhs; e += NT) {
                const int c = e % a.nrhs, i = e / a.nrhs;
                xr[e] = Wre[(size_t)i * lw + n + c];
                xi[e] = Wim[(size_t)i * lw + n + c];
            }
            __syncthreads();
        }
        }   // pass
        TICK(4);
        // ---------------- observable
        double acc = 0.0;
        if (a.mode == 2) {
            // w^2 Re sum_c [G^r Sigma^K G^a]_cc  with  Z[a,c] = G^r[sel_c,a] (pass 1) and X[b,c] = G^a[b,sel_c] (pass 0)
            const double *xr = a.Xs + (size_t)blockIdx.x * 2 * n * a.nrhs, *xi = xr + (size_t)n * a.nrhs;
            const double kd = a.kd[iw], kr1 = a.kr1[iw], kr2 = a.kr2[iw], ki = a.ki[iw];
            for (int e = threadIdx.x; e < nl * a.nrhs; e += NT) {          // diagonal lead part: kd * mask_a
                const int c = e % a.nrhs, i = e / a.nrhs;
                const double m = a.sig_mask[i];
                if (m != 0.0) {
                    const double zr2 = Wre[(size_t)i * lw + n + c], zi2 = Wim[(size_t)i * lw + n + c];
                    acc += kd * m * (zr2 * xr[(size_t)i * a.nrhs + c] - zi2 * xi[(size_t)i * a.nrhs + c]);
                }
            }
            for (int e = threadIdx.x; e < a.nb * a.nb * a.nrhs; e += NT) { // dense bias block
                const int c = e % a.nrhs, ab = e / a.nrhs, ia = ab / a.nb, ib = ab % a.nb;
                const double skr = kr1 * a.bdamp[ab] + kr2 * a.chiplus[ab], ski = ki * a.chiminus[ab];
                const size_t ra = (size_t)(a.b0 + ia) * lw + n + c, rb = (size_t)(a.b0 + ib) * a.nrhs + c;
                const double zr2 = Wre[ra], zi2 = Wim[ra], xr2 = xr[rb], xi2 = xi[rb];
                // Re[ z * sk * x ]
                const double tr = skr * xr2 - ski * xi2, ti = skr * xi2 + ski * xr2;
                acc += zr2 * tr - zi2 * ti;
            }
        } else if (a.mode == 0) {
            for (int e = threadIdx.x; e < a.nrows * a.nrhs; e += NT) {
                const int c = e % a.nrhs, i = a.rows[e / a.nrhs];
                const double xr = Wre[(size_t)i * lw + n + c], xi = Wim[(size_t)i * lw + n + c];
                acc += xr * xr + xi * xi;
            }
        } else {
            for (int c = threadIdx.x; c < a.nrhs; c += NT) acc += Wim[(size_t)a.rhs[c] * lw + n + c];
        }
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) {
            const double gam = 2.0 * w / a.damp;
            a.out[iw] = a.mode == 0 ? gam * gam * acc : (a.mode == 1 ? -2.0 * w * w * a.weight[iw] * acc : w * w * acc);
            a.status[iw] = s_bad;
        }
        __syncthreads();
    }
    if (a.timing && blockIdx.x == 0 && threadIdx.x == 0)
        for (int k = 0; k < 6; ++k) a.timing[k] = tacc[k];
#undef TICK
}

struct BiasBlock {
    int b0 = 0, nb = 0;
    const double *bdamp = nullptr, *chiplus = nullptr, *chiminus = nullptr;
    double bias = 0.0;
    const double *kd = nullptr, *kr1 = nullptr, *kr2 = nullptr, *ki = nullptr;
};

int run_lu(int device, int n, const double *K, const int32_t *idxL, int nL, const int32_t *idxR, int nR, double damp,
           const double *omegas, int nw, int mode, const double *weight, const int32_t *sel, int nsel, double *out,
           const BiasBlock *bb = nullptr) {
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
        SCLMD_REQUIRE(sel && nsel > 0 && (weight || mode == 2), "bpt.ps: empty selection");
        for (int i = 0; i < nsel; ++i) SCLMD_REQUIRE(sel[i] >= 0 && sel[i] < n, "bpt.ps: selected dof %d out of range", sel[i]);
        rhs.assign(sel, sel + nsel);
        rows = rhs;
        row_stop = *std::min_element(rows.begin(), rows.end());
        if (mode == 2) {   // the Keldysh self-energy lives on the lead dofs and on the bias block: those rows of both solutions
            row_stop = 0;
            for (int i = 0; i < n; ++i) if (mask[i] != 0.0) { row_stop = i; break; }
            if (bb && bb->nb > 0) row_stop = std::min(row_stop, bb->b0);
        }
    }
    if (bb && bb->nb > 0)
        SCLMD_REQUIRE(bb->b0 >= 0 && bb->b0 + bb->nb <= n && bb->bdamp && bb->chiminus && bb->chiplus, "bpt: bad bias block");
    SCLMD_REQUIRE(mode != 2 || (bb && bb->nb > 0 && bb->kd && bb->kr1 && bb->kr2 && bb->ki), "bpt.ps (biased): missing Keldysh weights");
    LuArgs a{};
    a.n = n; a.np = n + (n & 1); a.nrhs = (int)rhs.size(); a.ld = round_up(a.np, 16) + 4;   // ld % 16 == 4: conflict-free 64-bit DMMA fragment loads from the panel
    if (a.ld - 16 >= a.np) a.ld -= 16;
    a.ncols = a.np + a.nrhs; a.nw = nw; a.mode = mode;
    a.lw = round_up(a.ncols + CW, 2);   // slack of one chunk: the update prefetches whole 64-column tiles
    a.nrows = (int)rows.size(); a.row_stop = row_stop; a.damp = damp; a.eps = 1e-9;
    const size_t smem = ((size_t)2 * NB * a.ld + 2 * NB * UW + 64) * sizeof(double);
    SCLMD_REQUIRE(smem <= 220 * 1024, "bpt: n=%d too large for the shared-memory panel (max ~850)", n);
    const int grid = std::min(nw, 2 * sm_count(device));
    DevBuf<double> dK, dmask, dom, dwt, W, dout;
    DevBuf<int> drhs, drows, dstat;
    SCLMD_CUDA(dK.alloc((size_t)n * n)); SCLMD_CUDA(dmask.alloc(n)); SCLMD_CUDA(dom.alloc(nw)); SCLMD_CUDA(dwt.alloc(nw));
    SCLMD_CUDA(W.alloc((size_t)grid * 2 * a.np * a.lw)); SCLMD_CUDA(dout.alloc(nw));
    SCLMD_CUDA(drhs.alloc(rhs.size())); SCLMD_CUDA(drows.alloc(rows.size())); SCLMD_CUDA(dstat.alloc(nw));
    SCLMD_CUDA(cudaMemcpy(dK.p, K, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(dmask.p, mask.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(dom.p, omegas, nw * sizeof(double), cudaMemcpyHostToDevice));
    if (weight) SCLMD_CUDA(cudaMemcpy(dwt.p, weight, nw * sizeof(double), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(drhs.p, rhs.data(), rhs.size() * sizeof(int), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(drows.p, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice));
    a.K = dK.p; a.sig_mask = dmask.p; a.rhs = drhs.p; a.rows = drows.p; a.omegas = dom.p; a.weight = dwt.p;
    a.W = W.p; a.out = dout.p; a.status = dstat.p;
    DevBuf<double> dbd, dcp, dcm, dkw, dXs;
    if (bb && bb->nb > 0) {
        const size_t n2 = (size_t)bb->nb * bb->nb;
        SCLMD_CUDA(dbd.alloc(n2)); SCLMD_CUDA(dcp.alloc(n2)); SCLMD_CUDA(dcm.alloc(n2));
        SCLMD_CUDA(cudaMemcpy(dbd.p, bb->bdamp, n2 * sizeof(double), cudaMemcpyHostToDevice));
        SCLMD_CUDA(cudaMemcpy(dcp.p, bb->chiplus, n2 * sizeof(double), cudaMemcpyHostToDevice));
        SCLMD_CUDA(cudaMemcpy(dcm.p, bb->chiminus, n2 * sizeof(double), cudaMemcpyHostToDevice));
        a.b0 = bb->b0; a.nb = bb->nb; a.bdamp = dbd.p; a.chiplus = dcp.p; a.chiminus = dcm.p; a.bias = bb->bias;
        if (mode == 2) {
            SCLMD_CUDA(dkw.alloc((size_t)4 * nw));
            SCLMD_CUDA(cudaMemcpy(dkw.p, bb->kd, nw * sizeof(double), cudaMemcpyHostToDevice));
            SCLMD_CUDA(cudaMemcpy(dkw.p + nw, bb->kr1, nw * sizeof(double), cudaMemcpyHostToDevice));
            SCLMD_CUDA(cudaMemcpy(dkw.p + 2 * (size_t)nw, bb->kr2, nw * sizeof(double), cudaMemcpyHostToDevice));
            SCLMD_CUDA(cudaMemcpy(dkw.p + 3 * (size_t)nw, bb->ki, nw * sizeof(double), cudaMemcpyHostToDevice));
            a.kd = dkw.p; a.kr1 = dkw.p + nw; a.kr2 = dkw.p + 2 * (size_t)nw; a.ki = dkw.p + 3 * (size_t)nw;
            SCLMD_CUDA(dXs.alloc((size_t)grid * 2 * a.np * a.nrhs));
            a.Xs = dXs.p;
        }
    }
    DevBuf<long long> dtim;
    const bool want_timing = getenv("SCLMD_BPT_TIMING") != nullptr;
    if (want_timing) { SCLMD_CUDA(dtim.alloc(8)); a.timing = dtim.p; }
    SCLMD_CUDA(cudaFuncSetAttribute(k_bpt_lu, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    SCLMD_CUDA(cudaEventCreate(&e0));
    SCLMD_CUDA(cudaEventCreate(&e1));
    SCLMD_CUDA(cudaEventRecord(e0));
    k_bpt_lu<<<grid, NT, smem>>>(a);
    SCLMD_CUDA(cudaGetLastError());
    SCLMD_CUDA(cudaEventRecord(e1));
    SCLMD_CUDA(cudaDeviceSynchronize());
    float kms = 0;
    SCLMD_CUDA(cudaEventElapsedTime(&kms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (want_timing) fprintf(stderr, "[bpt] kernel %.3f ms for %d frequencies on %d CTAs (%.0f omega/s device-only)\n", kms, nw, grid, nw / (kms * 1e-3));
    SCLMD_CUDA(cudaMemcpy(out, dout.p, nw * sizeof(double), cudaMemcpyDeviceToHost));
    if (want_timing) {
        long long t[8];
        SCLMD_CUDA(cudaMemcpy(t, dtim.p, sizeof(t), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[bpt timing, CTA 0 cycles] build %lld panel %lld swap+trsm %lld update %lld backsub %lld observable+loop %lld\n", t[0], t[1], t[2], t[3], t[4], t[5]);
    }
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

// bpt.tm with a biased electron bath attached (bpt.setbias, negf.py:27-37): G includes Sigma_b^r = -i w bdamp - bias chiminus
// on the contiguous dof block [b0, b0+nb) (negf.py:162-172); bias in angular units (eV/hbar, as bpt stores it)
int sclmd_bpt_tm_bias(int device, int n, const double *K, const int32_t *idxL, int nL, const int32_t *idxR, int nR, double damp,
                      int b0, int nb, const double *bdamp, const double *chiplus, const double *chiminus, double bias,
                      const double *omegas, int nw, double *tm_out) {
    BiasBlock bb;
    bb.b0 = b0; bb.nb = nb; bb.bdamp = bdamp; bb.chiplus = chiplus; bb.chiminus = chiminus; bb.bias = bias;
    return run_lu(device, n, K, idxL, nL, idxR, nR, damp, omegas, nw, 0, nullptr, nullptr, 0, tm_out, &bb);
}

// bpt.ps with bias (negf.py:234-236): w^2 Re Tr[(G^r Sigma^K G^a)[sel,sel]], Sigma^K = totalkselfenergy (negf.py:177-193)
//   = kd(w) on the lead dofs (kd = (2w/damp) n_B) + kr1 bdamp + kr2 chiplus + i ki chiminus on the bias block
int sclmd_bpt_ps_bias(int device, int n, const double *K, const int32_t *idxL, int nL, const int32_t *idxR, int nR, double damp,
                      int b0, int nb, const double *bdamp, const double *chiplus, const double *chiminus, double bias,
                      const double *omegas, const double *kd, const double *kr1, const double *kr2, const double *ki, int nw,
                      const int32_t *sel, int nsel, double *ps_out) {
    BiasBlock bb;
    bb.b0 = b0; bb.nb = nb; bb.bdamp = bdamp; bb.chiplus = chiplus; bb.chiminus = chiminus; bb.bias = bias;
    bb.kd = kd; bb.kr1 = kr1; bb.kr2 = kr2; bb.ki = ki;
    return run_lu(device, n, K, idxL, nL, idxR, nR, damp, omegas, nw, 2, nullptr, sel, nsel, ps_out, &bb);
}

}  // extern "C"
