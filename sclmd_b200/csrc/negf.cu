// Ballistic phonon transmission sweeps (replaces bpt.retargf / bpt.advangf / bpt.tm / bpt.ps / bpt.gettm / bpt.getps,
// sclmd/negf.py:104-150,153-193,206-212,228-242).
//
// For every frequency  M(w) = (w + 1e-9 i)^2 I - K - Sigma_L(w) - Sigma_R(w) [- Sigma_bias(w)]  is factorised by blocked LU with
// partial pivoting, the right-hand sides (unit vectors) ride along as extra columns so the trailing update also performs the
// forward substitution, and only the rows of G = M^-1 that the observable needs are back-substituted:
//   tm :  T = Re Tr[G Gamma_L G^dagger Gamma_R] = (2w/damp)^2 sum_{i in R, j in L} |G_ij|^2   (Gamma diagonal, negf.py:214-215)
//   ps :  -2 w^2 n_B Tr Im G[sel,sel]  (negf.py:232)   /   w^2 Re Tr[(G^r Sigma^K G^a)[sel,sel]]  (negf.py:236, two factorisations)
// The reference forms two full inverses and three dense n^3 products per frequency instead.
//
// B200 design.  A rank-k update of an HBM-resident matrix moves 32 bytes per 8k flops, so k = 16 needs ~9 TB/s at the FP64 peak
// and is HBM-bound; the factorisation is therefore organised as a BATCH of frequencies moving in lock step through
//   k_build       M(w) and the unit columns, one pass over W
//   k_panel       16-column sub-panel, one CTA per frequency, rows held in REGISTERS, one barrier per column; then - thread per
//                 column - the U rows of the other columns of the current 64-column block
//   k_block_trsm  once per 64-column block, persistent CTA per frequency over 32-column chunks right of the block:
//                 U12 = L11^-1 A12 (64 rows) with the in-block updates on the FP64 tensor pipe, next chunk prefetched in registers
//   k_gemm        complex rank-16 (panel columns of the block) / rank-64 (trailing matrix) update on the FP64 tensor pipe
//                 (4 real DMMA.8x8x4 per complex fragment pair), 64x64 tiles, K slabs through a 3-stage cp.async ring
//   k_block_trsm_upper + k_gemm   back substitution of the rows >= row_stop in 64-row blocks from the bottom, on the tensor pipe
//   k_observe     reduction to one number per frequency
// with several batches in flight on separate streams so that the latency-bound kernels of one batch fill the tails of another.
// Pivoting is IMPLICIT over the whole factorisation: no row of W ever moves.  Every frequency carries an index list
// act[position] -> physical row; choosing a pivot permutes 16-32 entries of that list (pivot j to position kk+j, the displaced
// entries into the vacated positions) and every kernel addresses rows through it.
// The dofs are re-ordered [not needed | right-hand-side dofs | rows needed]: the unit columns are then zero in every row above
// their own dof, and a 64x64 update tile whose U12 slab is exactly zero is skipped (bit-identical result, ~1/3 fewer flops).
//
// Layout: W = [M | E] row-major, planar complex (real plane, imaginary plane), nrp x lw doubles per plane (rows padded to 64
// plus one spare tile, columns to 64, pads zero), one slot per frequency of a batch.
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <numeric>

#include "common.cuh"

using namespace sclmd;

namespace {

constexpr int TS = 64;        // outer block (rank of the trailing update) and GEMM tile edge
constexpr int LDS_T = TS + 4; // padded shared-memory row (== 4 mod 16 doubles: conflict-free 64-bit DMMA fragment loads)
constexpr int GK = 16, GSTG = 3, GLDA = GK + 4;                 // k_gemm: slab depth, ring stages, leading dim of the A slab
constexpr int GSTAGE = 2 * TS * GLDA + 2 * GK * LDS_T;          // doubles per stage: Are | Aim | Bre | Bim
constexpr size_t GEMM_SMEM = (size_t)GSTG * GSTAGE * sizeof(double);
constexpr int TC = 32, LDT = TC + 4;                            // k_block_trsm: columns per chunk, leading dim of the chunk tile
constexpr size_t TRSM_SMEM = (size_t)(2 * TS * LDS_T + 2 * TS * LDT) * sizeof(double);

struct Geo {
    int nl;      // logical dimension
    int np;      // even working dimension (an identity dof pads an odd system)
    int nrp;     // rows allocated (multiple of 64, plus one spare tile: update tiles start at any multiple of 8)
    int lw;      // leading dimension (multiple of 64)
    int nrhs, ncols;
    size_t plane;    // nrp * lw
};

struct Problem {
    const double *K;          // [nl][nl] permuted
    const double *mask;       // [nl] number of leads on each dof
    const int *rhs;           // [nrhs] row index of the 1 in every unit column
    const int *rows;          // [nrows] rows of G entering the tm observable
    int nrows;
    const int *bmap;          // [nl] index inside the bias block or -1
    const int *bpos;          // [nb] row of every bias-block dof
    int nb;
    const double *bdamp, *chiplus, *chiminus;   // [nb][nb]
    double bias, damp, eps;
    const double *omegas, *weight, *kd, *kr1, *kr2, *ki;   // per frequency
};

__device__ __forceinline__ void cfma_sub(double &cr, double &ci, double ar, double ai, double br, double bi) {
    // c -= a*b
    cr = fma(-ar, br, cr); cr = fma(ai, bi, cr);
    ci = fma(-ar, bi, ci); ci = fma(-ai, br, ci);
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------ build
// grid (nrp, batch): one row of W per CTA.  sgn = -1 builds the advanced matrix (Sigma^r-dagger), tblk transposes the bias block
// (mode 2: pass 0 = M^a, pass 1 = (M^r)^T; negf.py:210-212 keeps the +i eps of z in advangf).
__global__ void __launch_bounds__(128) k_build(Geo g, Problem p, double *W, int *act, int w0, double sgn, int tblk) {
    const int i = blockIdx.x, b = blockIdx.y;
    if (threadIdx.x == 0) act[(size_t)b * g.nrp + i] = i;
    const double w = p.omegas[w0 + b];
    const double zr = w * w - p.eps * p.eps, zi = 2.0 * w * p.eps, sg = w / p.damp;
    double *wre = W + (size_t)b * 2 * g.plane + (size_t)i * g.lw, *wim = wre + g.plane;
    const int bi0 = (i < g.nl && p.nb > 0) ? p.bmap[i] : -1;
    for (int j = threadIdx.x; j < g.lw; j += blockDim.x) {
        double mr = 0.0, mi = 0.0;
        if (i < g.np && j < g.np) {
            if (i < g.nl && j < g.nl) {
                mr = -p.K[(size_t)i * g.nl + j];
                if (i == j) { mr += zr; mi = zi + sgn * sg * p.mask[i]; }
                if (bi0 >= 0) {
                    const int bj0 = p.bmap[j];
                    if (bj0 >= 0) {
                        const int bi = tblk ? bj0 : bi0, bj = tblk ? bi0 : bj0;
                        mr += p.bias * p.chiminus[bi * p.nb + bj];      // M -= Sigma_b :  +bias chi-  and  +i w bdamp
                        mi += sgn * w * p.bdamp[bi * p.nb + bj];
                    }
                }
            } else if (i == j) {
                mr = 1.0;
            }
        } else if (i < g.np && j < g.ncols) {
            mr = p.rhs[j - g.np] == i ? 1.0 : 0.0;
        }
        wre[j] = mr;
        wim[j] = mi;
    }
}

// ------------------------------------------------------------------------------------------------ sub-panel
// One CTA per frequency factorises columns [kk, kk+kb) over the rows at positions [kk, np).  Thread t owns positions
// kk + t + r*NT (r < RPT) in registers.  A chosen row is frozen (it becomes a row of U), the winner of each warp publishes its
// row next to its magnitude, so one barrier per column suffices.  Afterwards the index list is permuted (pivot j to position
// kk+j, unchosen top entries into the vacated positions); the rows themselves are written back where they came from.
// Pivot rule: max |re|+|im| (izamax), ties to the smallest position.
template <int NT, int RPT, int NBW>
__global__ void __launch_bounds__(NT, RPT == 1 ? 512 / NT : 1) k_panel(Geo g, double *W, int *act, int *status, int w0, int kk, int kb, int rend) {
    constexpr int NWARP = NT / 32;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = g.np - kk, lw = g.lw;
    double *wre = W + (size_t)b * 2 * g.plane, *wim = wre + g.plane;
    int *actb = act + (size_t)b * g.nrp;
    __shared__ double c_best[2][NWARP];
    __shared__ int c_arg[2][NWARP];
    __shared__ double c_row[2][NWARP][2][NBW];
    __shared__ int s_piv[NBW], s_bad;
    __shared__ int s_src[2 * NBW], s_dst[2 * NBW], s_n, s_phys[NBW];
    __shared__ double s_lr[NBW][NBW + 1], s_li[NBW][NBW + 1];      // L11 in pivot order for the U rows of the side columns

    double ar[RPT][NBW], ai[RPT][NBW];
    int ord[RPT];
    size_t off[RPT];
    bool live[RPT];      // valid row, not yet chosen as a pivot
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int rel = tid + r * NT;
        live[r] = rel < m;
        ord[r] = -1;
        off[r] = (size_t)actb[kk + min(rel, m - 1)] * lw + kk;
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
#pragma unroll
        for (int jj = 0; jj < NBW; jj += 2) {
            double2 vr = make_double2(0.0, 0.0), vi = vr;
            if (live[r] && jj < kb) {          // kb is even: a pair is inside or outside the panel as a whole
                vr = *reinterpret_cast<const double2 *>(wre + off[r] + jj);
                vi = *reinterpret_cast<const double2 *>(wim + off[r] + jj);
            }
            ar[r][jj] = vr.x; ar[r][jj + 1] = vr.y;
            ai[r][jj] = vi.x; ai[r][jj + 1] = vi.y;
        }
    }
    if (tid == 0) s_bad = 0;

#pragma unroll
    for (int j = 0; j < NBW; ++j) {
        if (j < kb) {
            double best = -1.0;
            int arg = INT_MAX;
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const double v = fabs(ar[r][j]) + fabs(ai[r][j]);
                if (live[r] && v > best) { best = v; arg = tid + r * NT; }
            }
            const double mybest = best;
            const int myarg = arg;
            {   // warp arg-max with three redux.sync: non-negative doubles order like their bit patterns (a NaN wins and is
                // reported as a singular pivot below); ties go to the smallest position
                const long long key = __double_as_longlong(best);          // -1.0 (no live row) is negative
                const int hi = (int)(key >> 32);
                const int mh = __reduce_max_sync(0xffffffffu, hi);
                const unsigned lo = hi == mh ? (unsigned)(key & 0xffffffffll) : 0u;
                const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
                const bool top = hi == mh && lo == ml && mh >= 0;
                arg = __reduce_min_sync(0xffffffffu, top ? myarg : INT_MAX);
                best = __longlong_as_double(((long long)mh << 32) | (long long)ml);
            }
            const int buf = j & 1;
            if (arg == myarg && mybest >= 0.0) {      // this lane holds the warp's candidate: publish the rest of its row
#pragma unroll
                for (int r = 0; r < RPT; ++r)
                    if (tid + r * NT == arg) {
#pragma unroll
                        for (int jj = j; jj < NBW; ++jj) {
                            c_row[buf][warp][0][jj] = ar[r][jj];
                            c_row[buf][warp][1][jj] = ai[r][jj];
                        }
                    }
            }
            if (lane == 0) { c_best[buf][warp] = best; c_arg[buf][warp] = arg; }
            __syncthreads();
            double wbest = c_best[buf][0];
            int warg = c_arg[buf][0], ww = 0;
#pragma unroll
            for (int q = 1; q < NWARP; ++q) {
                const double qb = c_best[buf][q];
                const int qa = c_arg[buf][q];
                if (qb > wbest || (qb == wbest && qa < warg)) { wbest = qb; warg = qa; ww = q; }
            }
            if (tid == 0) {
                s_piv[j] = warg == INT_MAX ? j : warg;
                if (!(wbest > 0.0)) s_bad = 1;
            }
            const double dr = c_row[buf][ww][0][j], di = c_row[buf][ww][1][j];
            const double rn = __drcp_rn(dr * dr + di * di);
            const double ir = dr * rn, ii = -di * rn;          // 1 / pivot
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                if (!live[r]) continue;
                if (tid + r * NT == warg) {
                    live[r] = false;
                    ord[r] = j;
                } else {
                    const double lr = ar[r][j] * ir - ai[r][j] * ii, li = ar[r][j] * ii + ai[r][j] * ir;
                    ar[r][j] = lr;
                    ai[r][j] = li;
#pragma unroll
                    for (int jj = j + 1; jj < NBW; ++jj)
                        cfma_sub(ar[r][jj], ai[r][jj], lr, li, c_row[buf][ww][0][jj], c_row[buf][ww][1][jj]);
                }
            }
        }
    }

    // permutation of the index list: pivot j -> position kk+j ; unchosen top entries -> the vacated positions (warp 0, ballots)
    __syncthreads();
    if (warp == 0) {
        const int pj = lane < kb ? s_piv[lane] : INT_MAX;                     // lane j: relative position of pivot j
        const unsigned chosen = __reduce_or_sync(0xffffffffu, pj < kb ? 1u << pj : 0u);   // top positions taken as pivots
        const unsigned below = __ballot_sync(0xffffffffu, lane < kb && pj >= kb);         // pivots that vacate a lower position
        const bool loose = lane < kb && !((chosen >> lane) & 1u);            // top entry that must move down
        const unsigned loosem = __ballot_sync(0xffffffffu, loose);
        const int q = __popc(loosem & ((1u << lane) - 1u));
        const int vl = __fns(below, 0, q + 1);                                // lane of the q-th vacating pivot
        const int vpos = __shfl_sync(0xffffffffu, pj, loose ? (vl & 31) : 0);
        if (lane < kb) {
            s_src[lane] = kk + pj;
            s_dst[lane] = kk + lane;
        }
        if (loose) {
            s_src[kb + q] = kk + lane;
            s_dst[kb + q] = kk + vpos;
        }
        if (lane == 0) {
            s_n = kb + __popc(below);
            if (s_bad) status[w0 + b] = 1;
        }
    }
    __syncthreads();
    int moved = 0;
    if (tid < s_n) moved = actb[s_src[tid]];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {          // rows go back where they came from (L below the pivots, U in the pivot rows)
        if (tid + r * NT >= m) continue;
#pragma unroll
        for (int jj = 0; jj < NBW; jj += 2)
            if (jj < kb) {
                *reinterpret_cast<double2 *>(wre + off[r] + jj) = make_double2(ar[r][jj], ar[r][jj + 1]);
                *reinterpret_cast<double2 *>(wim + off[r] + jj) = make_double2(ai[r][jj], ai[r][jj + 1]);
            }
        if (ord[r] >= 0) {
#pragma unroll
            for (int jj = 0; jj < NBW; ++jj) { s_lr[ord[r]][jj] = ar[r][jj]; s_li[ord[r]][jj] = ai[r][jj]; }
        }
    }
    __syncthreads();
    if (tid < s_n) {
        actb[s_dst[tid]] = moved;
        if (tid < kb) s_phys[tid] = moved;
    }
    __syncthreads();
    // the remaining panel columns [kk+kb, rend) of this 64-column block, thread per column: U = L11^-1 A on the pivot rows
    // (k_gemm then updates them below).  Columns right of the block are handled once per block by k_block_trsm.
    const int nside = rend - kk - kb;
    if (tid < nside) {
        const int c = kk + kb + tid;
        double vr[NBW], vi[NBW];
#pragma unroll
        for (int t = 0; t < NBW; ++t) {
            vr[t] = vi[t] = 0.0;
            if (t < kb) {
                vr[t] = wre[(size_t)s_phys[t] * lw + c];
                vi[t] = wim[(size_t)s_phys[t] * lw + c];
            }
        }
#pragma unroll
        for (int jj = 0; jj < NBW; ++jj)
#pragma unroll
            for (int i2 = jj + 1; i2 < NBW; ++i2)
                if (i2 < kb) cfma_sub(vr[i2], vi[i2], s_lr[i2][jj], s_li[i2][jj], vr[jj], vi[jj]);
#pragma unroll
        for (int t = 0; t < NBW; ++t)
            if (t < kb) {
                wre[(size_t)s_phys[t] * lw + c] = vr[t];
                wim[(size_t)s_phys[t] * lw + c] = vi[t];
            }
    }
}

// ------------------------------------------------------------------------------------------------ U12 of a 64-column block
// Persistent CTA per frequency over 32-column chunks right of the block (c >= rend).  L (64x64, strictly lower, pivot order)
// sits in shared memory as the DMMA A operand; a chunk (64 pivot rows x 32 columns) is staged in shared memory as the B / C
// operand.  Per sub-panel: U_j = inv(L_jj) T_j, then  T_below -= L_below U_j, both on the tensor pipe (the diagonal blocks are
// inverted once per CTA).  The next chunk is prefetched into registers meanwhile; chunks that are still exactly zero are skipped.
template <int NBW>
__global__ void __launch_bounds__(256, 2) k_block_trsm(Geo g, double *W, const int *act, int *flags, int k0, int rend) {
    extern __shared__ double sm[];
    double *Lr = sm, *Li = Lr + TS * LDS_T, *Tr = Li + TS * LDS_T, *Ti = Tr + TS * LDT;
    __shared__ int s_row[TS];
    const int b = blockIdx.x, lw = g.lw, nblk = rend - k0, nsub = (nblk + NBW - 1) / NBW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fk = lane & 3;
    double *wre = W + (size_t)b * 2 * g.plane, *wim = wre + g.plane;
    const int *actb = act + (size_t)b * g.nrp;
    if (tid < TS) s_row[tid] = actb[k0 + min(tid, nblk - 1)];
    __syncthreads();
    for (int e = tid; e < TS * TS; e += 256) {
        const int kx = e % TS, i = e / TS;
        const bool ok = i < nblk && kx < i;
        Lr[i * LDS_T + kx] = ok ? wre[(size_t)s_row[i] * lw + k0 + kx] : 0.0;
        Li[i * LDS_T + kx] = ok ? wim[(size_t)s_row[i] * lw + k0 + kx] : 0.0;
    }
    __syncthreads();
    // the diagonal blocks are replaced by their inverses (unit lower triangular, NBW x NBW): the solve of a sub-panel's rows
    // becomes a DMMA product that all warps share.  Thread (j, c) builds column c of inv(L_jj) by forward substitution.
    {
        const int j = tid / NBW, c = tid % NBW, jb = j * NBW;
        double xr[NBW], xi[NBW];
        if (tid < nsub * NBW) {
#pragma unroll
            for (int i = 0; i < NBW; ++i) { xr[i] = i == c ? 1.0 : 0.0; xi[i] = 0.0; }
#pragma unroll
            for (int jj = 0; jj < NBW; ++jj)
#pragma unroll
                for (int i2 = jj + 1; i2 < NBW; ++i2)
                    cfma_sub(xr[i2], xi[i2], Lr[(jb + i2) * LDS_T + jb + jj], Li[(jb + i2) * LDS_T + jb + jj], xr[jj], xi[jj]);
        }
        __syncthreads();
        if (tid < nsub * NBW) {
#pragma unroll
            for (int i = 0; i < NBW; ++i) {
                Lr[(jb + i) * LDS_T + jb + c] = xr[i];
                Li[(jb + i) * LDS_T + jb + c] = xi[i];
            }
        }
    }
    const int nchunk = (g.ncols - rend + TC - 1) / TC;
    double2 cur[8], nxt[8];
    auto fetch = [&](int ch, double2 (&buf)[8]) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = tid + u * 256, pl = e >> 10, row = (e >> 4) & 63, col = rend + ch * TC + (e & 15) * 2;
            buf[u] = make_double2(0.0, 0.0);
            if (row < nblk && col < g.ncols) buf[u] = *reinterpret_cast<const double2 *>((pl ? wim : wre) + (size_t)s_row[row] * lw + col);
        }
    };
    fetch(0, cur);
    for (int ch = 0; ch < nchunk; ++ch) {
        int any = 0;
        __syncthreads();          // L staged (first pass) / previous chunk written back
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = tid + u * 256, pl = e >> 10, row = (e >> 4) & 63, cp = (e & 15) * 2;
            *reinterpret_cast<double2 *>((pl ? Ti : Tr) + row * LDT + cp) = cur[u];
            any |= (cur[u].x != 0.0) | (cur[u].y != 0.0);
        }
        any = __syncthreads_or(any);
        if (tid == 0) flags[(size_t)b * (g.lw / TC) + ch] = any;          // k_gemm skips update tiles whose U12 chunks are all zero
        if (ch + 1 < nchunk) fetch(ch + 1, nxt);
        if (any) {
            for (int j = 0; j < nsub; ++j) {
                const int jb = j * NBW, kb = min(NBW, nblk - jb);
                if (warp < TC / 8) {      // U_j = inv(L_jj) T_j : warp w owns the 8-column fragment w (reads its inputs, then writes)
                    const int cbase = warp * 8;
                    double ur[NBW / 4], ui[NBW / 4];
#pragma unroll
                    for (int k4 = 0; k4 < NBW; k4 += 4) {
                        ur[k4 / 4] = Tr[(jb + k4 + fk) * LDT + cbase + fr];
                        ui[k4 / 4] = Ti[(jb + k4 + fk) * LDT + cbase + fr];
                    }
                    double2 yr[NBW / 8], yi[NBW / 8];
#pragma unroll
                    for (int mf = 0; mf < NBW / 8; ++mf) {
                        yr[mf] = yi[mf] = make_double2(0.0, 0.0);
#pragma unroll
                        for (int k4 = 0; k4 < NBW; k4 += 4) {
                            const double alr = Lr[(jb + mf * 8 + fr) * LDS_T + jb + k4 + fk], ali = Li[(jb + mf * 8 + fr) * LDS_T + jb + k4 + fk];
                            dmma(yr[mf].x, yr[mf].y, alr, ur[k4 / 4]);
                            dmma(yr[mf].x, yr[mf].y, -ali, ui[k4 / 4]);
                            dmma(yi[mf].x, yi[mf].y, alr, ui[k4 / 4]);
                            dmma(yi[mf].x, yi[mf].y, ali, ur[k4 / 4]);
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int mf = 0; mf < NBW / 8; ++mf) {
                        *reinterpret_cast<double2 *>(Tr + (jb + mf * 8 + fr) * LDT + cbase + 2 * fk) = yr[mf];
                        *reinterpret_cast<double2 *>(Ti + (jb + mf * 8 + fr) * LDT + cbase + 2 * fk) = yi[mf];
                    }
                }
                __syncthreads();
                // rows below inside the block: (nblk - jb - kb) rows in 8-row fragments x 4 column fragments over the 8 warps
                const int rb0 = jb + kb, nmf = (nblk - rb0 + 7) / 8;
                for (int q = warp; q < nmf * (TC / 8); q += 8) {
                    const int rbase = rb0 + (q / (TC / 8)) * 8, cbase = (q % (TC / 8)) * 8;
                    double2 cr = *reinterpret_cast<const double2 *>(Tr + (rbase + fr) * LDT + cbase + 2 * fk);
                    double2 ci = *reinterpret_cast<const double2 *>(Ti + (rbase + fr) * LDT + cbase + 2 * fk);
#pragma unroll
                    for (int k4 = 0; k4 < NBW; k4 += 4) {
                        if (k4 < kb) {
                            const double alr = Lr[(rbase + fr) * LDS_T + jb + k4 + fk], ali = Li[(rbase + fr) * LDS_T + jb + k4 + fk];
                            const double ur = Tr[(jb + k4 + fk) * LDT + cbase + fr], ui = Ti[(jb + k4 + fk) * LDT + cbase + fr];
                            dmma(cr.x, cr.y, -alr, ur);
                            dmma(cr.x, cr.y, ali, ui);
                            dmma(ci.x, ci.y, -alr, ui);
                            dmma(ci.x, ci.y, -ali, ur);
                        }
                    }
                    *reinterpret_cast<double2 *>(Tr + (rbase + fr) * LDT + cbase + 2 * fk) = cr;
                    *reinterpret_cast<double2 *>(Ti + (rbase + fr) * LDT + cbase + 2 * fk) = ci;
                }
                __syncthreads();
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = tid + u * 256, pl = e >> 10, row = (e >> 4) & 63, cp = (e & 15) * 2, col = rend + ch * TC + cp;
                if (row < nblk && col < g.ncols)
                    *reinterpret_cast<double2 *>((pl ? wim : wre) + (size_t)s_row[row] * lw + col) =
                        *reinterpret_cast<const double2 *>((pl ? Ti : Tr) + row * LDT + cp);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) cur[u] = nxt[u];
    }
}

// ------------------------------------------------------------------------------------------------ back substitution, one block
// X_B = U_BB^-1 Y_B for the pivot rows at positions [b0, b1) (<= 64) and every right-hand side: the mirror image of
// k_block_trsm (upper triangular, non-unit diagonal, sub-blocks from the bottom up).  U_BB sits in shared memory as the DMMA A
// operand with its 16x16 diagonal blocks inverted once per CTA; the rows above inside the block are updated on the tensor pipe.
// The rows above the block get  Y -= U[above, B] X_B  from k_gemm.
__global__ void __launch_bounds__(256, 2) k_block_trsm_upper(Geo g, double *W, const int *act, int b0, int b1) {
    extern __shared__ double sm[];
    constexpr int NBW = 16;
    double *Ur = sm, *Ui = Ur + TS * LDS_T, *Tr = Ui + TS * LDS_T, *Ti = Tr + TS * LDT;
    __shared__ int s_row[TS];
    const int b = blockIdx.x, lw = g.lw, nblk = b1 - b0, nsub = (nblk + NBW - 1) / NBW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fr = lane >> 2, fk = lane & 3;
    double *wre = W + (size_t)b * 2 * g.plane, *wim = wre + g.plane;
    const int *actb = act + (size_t)b * g.nrp;
    if (tid < TS) s_row[tid] = actb[b0 + min(tid, nblk - 1)];
    __syncthreads();
    for (int e = tid; e < TS * TS; e += 256) {
        const int kx = e % TS, i = e / TS;
        const bool ok = i < nblk && kx < nblk && kx >= i;
        Ur[i * LDS_T + kx] = ok ? wre[(size_t)s_row[i] * lw + b0 + kx] : (i == kx ? 1.0 : 0.0);   // identity beyond a partial block
        Ui[i * LDS_T + kx] = ok ? wim[(size_t)s_row[i] * lw + b0 + kx] : 0.0;
    }
    __syncthreads();
    {   // thread (j, c): column c of inv(U_jj) by back substitution; x_i = 0 below the diagonal
        const int j = tid / NBW, c = tid % NBW, jb = j * NBW;
        double xr[NBW], xi[NBW];
        if (tid < nsub * NBW) {
#pragma unroll
            for (int i = NBW - 1; i >= 0; --i) {
                double sr = i == c ? 1.0 : 0.0, si = 0.0;
#pragma unroll
                for (int k = i + 1; k < NBW; ++k)
                    if (k <= c) cfma_sub(sr, si, Ur[(jb + i) * LDS_T + jb + k], Ui[(jb + i) * LDS_T + jb + k], xr[k], xi[k]);
                const double dr = Ur[(jb + i) * LDS_T + jb + i], di = Ui[(jb + i) * LDS_T + jb + i];
                const double rn = 1.0 / (dr * dr + di * di);
                const bool on = i <= c;
                xr[i] = on ? (sr * dr + si * di) * rn : 0.0;
                xi[i] = on ? (si * dr - sr * di) * rn : 0.0;
            }
        }
        __syncthreads();
        if (tid < nsub * NBW) {
#pragma unroll
            for (int i = 0; i < NBW; ++i) {
                Ur[(jb + i) * LDS_T + jb + c] = xr[i];
                Ui[(jb + i) * LDS_T + jb + c] = xi[i];
            }
        }
    }
    const int nchunk = (g.nrhs + TC - 1) / TC;
    for (int ch = 0; ch < nchunk; ++ch) {
        __syncthreads();          // inverses staged (first pass) / previous chunk written back
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = tid + u * 256, pl = e >> 10, row = (e >> 4) & 63, cp = (e & 15) * 2, col = g.np + ch * TC + cp;
            double2 v = make_double2(0.0, 0.0);
            if (row < nblk && col < g.ncols) v = *reinterpret_cast<const double2 *>((pl ? wim : wre) + (size_t)s_row[row] * lw + col);
            *reinterpret_cast<double2 *>((pl ? Ti : Tr) + row * LDT + cp) = v;
        }
        __syncthreads();
        for (int j = nsub - 1; j >= 0; --j) {
            const int jb = j * NBW;
            if (warp < TC / 8) {      // X_j = inv(U_jj) T_j : warp w owns the 8-column fragment w (reads its inputs, then writes)
                const int cbase = warp * 8;
                double ur[NBW / 4], ui[NBW / 4];
#pragma unroll
                for (int k4 = 0; k4 < NBW; k4 += 4) {
                    ur[k4 / 4] = Tr[(jb + k4 + fk) * LDT + cbase + fr];
                    ui[k4 / 4] = Ti[(jb + k4 + fk) * LDT + cbase + fr];
                }
                double2 yr[NBW / 8], yi[NBW / 8];
#pragma unroll
                for (int mf = 0; mf < NBW / 8; ++mf) {
                    yr[mf] = yi[mf] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int k4 = 0; k4 < NBW; k4 += 4) {
                        const double alr = Ur[(jb + mf * 8 + fr) * LDS_T + jb + k4 + fk], ali = Ui[(jb + mf * 8 + fr) * LDS_T + jb + k4 + fk];
                        dmma(yr[mf].x, yr[mf].y, alr, ur[k4 / 4]);
                        dmma(yr[mf].x, yr[mf].y, -ali, ui[k4 / 4]);
                        dmma(yi[mf].x, yi[mf].y, alr, ui[k4 / 4]);
                        dmma(yi[mf].x, yi[mf].y, ali, ur[k4 / 4]);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int mf = 0; mf < NBW / 8; ++mf) {
                    *reinterpret_cast<double2 *>(Tr + (jb + mf * 8 + fr) * LDT + cbase + 2 * fk) = yr[mf];
                    *reinterpret_cast<double2 *>(Ti + (jb + mf * 8 + fr) * LDT + cbase + 2 * fk) = yi[mf];
                }
            }
            __syncthreads();
            // rows above inside the block: jb rows in 8-row fragments x 4 column fragments over the 8 warps
            for (int q = warp; q < (jb / 8) * (TC / 8); q += 8) {
                const int rbase = (q / (TC / 8)) * 8, cbase = (q % (TC / 8)) * 8;
                double2 cr = *reinterpret_cast<const double2 *>(Tr + (rbase + fr) * LDT + cbase + 2 * fk);
                double2 ci = *reinterpret_cast<const double2 *>(Ti + (rbase + fr) * LDT + cbase + 2 * fk);
#pragma unroll
                for (int k4 = 0; k4 < NBW; k4 += 4) {
                    const double alr = Ur[(rbase + fr) * LDS_T + jb + k4 + fk], ali = Ui[(rbase + fr) * LDS_T + jb + k4 + fk];
                    const double xr = Tr[(jb + k4 + fk) * LDT + cbase + fr], xi = Ti[(jb + k4 + fk) * LDT + cbase + fr];
                    dmma(cr.x, cr.y, -alr, xr);
                    dmma(cr.x, cr.y, ali, xi);
                    dmma(ci.x, ci.y, -alr, xi);
                    dmma(ci.x, ci.y, -ali, xr);
                }
                *reinterpret_cast<double2 *>(Tr + (rbase + fr) * LDT + cbase + 2 * fk) = cr;
                *reinterpret_cast<double2 *>(Ti + (rbase + fr) * LDT + cbase + 2 * fk) = ci;
            }
            __syncthreads();
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = tid + u * 256, pl = e >> 10, row = (e >> 4) & 63, cp = (e & 15) * 2, col = g.np + ch * TC + cp;
            if (row < nblk && col < g.ncols)
                *reinterpret_cast<double2 *>((pl ? wim : wre) + (size_t)s_row[row] * lw + col) =
                    *reinterpret_cast<const double2 *>((pl ? Ti : Tr) + row * LDT + cp);
        }
    }
}

// ------------------------------------------------------------------------------------------------ rank-K update
// C[r0 + 64 by .. , c0 + 64 bx ..) -= A B  with  A = W[rows, ka:ka+K]  (L21)  and  B = W[ka:ka+K, cols]  (U12), K <= 64, complex planar.
// 8 warps = 2 row groups (32 rows) x 4 column groups (16 columns); the C tile is the accumulator itself:
//   Re += (-Ar) Br + Ai Bi ,  Im += (-Ar) Bi + (-Ai) Br      (four real DMMA.8x8x4 per fragment pair)
// The K range is cut into 16-deep slabs moved by 16-byte cp.async through a 3-stage shared-memory ring (A as [row][k], B as
// [k][col], both padded to a leading dimension == 4 mod 16 doubles: conflict-free 64-bit fragment loads), so the loads of the
// next slabs and of the C tile overlap the tensor work.  Tiles in the unit-column region whose B slab is exactly zero leave C
// unchanged and return early.
__device__ __forceinline__ void cp_async16(double *dst, const double *src, bool pred) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
    const int nbytes = pred ? 16 : 0;      // src-size 0: the 16 bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(src), "r"(nbytes) : "memory");
}

__global__ void __launch_bounds__(256, 2) k_gemm(Geo g, double *W, const int *act, const int *flags, unsigned long long *tiles, int r0, int c0,
                                                 int c1, int ka, int K) {
    extern __shared__ double sm[];
    __shared__ int s_ra[TS], s_rb[TS];          // physical rows of the C / A tile and of the B slab
    const int b = blockIdx.z, lw = g.lw;
    const int tr = r0 + TS * blockIdx.y, tc = c0 + TS * blockIdx.x;
    double *wre = W + (size_t)b * 2 * g.plane, *wim = wre + g.plane;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp & 1, wc = warp >> 1, fr = lane >> 2, fk = lane & 3;
    if (flags && tc >= g.np) {       // unit columns: skip tiles whose U12 slab is still exactly zero (recorded by k_block_trsm)
        const int ch = (tc - c0) / TC, nch = (g.ncols - c0 + TC - 1) / TC;
        const int *fl = flags + (size_t)b * (g.lw / TC);
        if (!(fl[ch] | (ch + 1 < nch ? fl[ch + 1] : 0))) return;
    }
    if (tiles && threadIdx.x == 0) atomicAdd(tiles, 1ull);
    {
        const int *actb = act + (size_t)b * g.nrp;
        if (tid < TS) s_ra[tid] = actb[tr + tid];
        else if (tid < 2 * TS) s_rb[tid - TS] = actb[ka + min(tid - TS, K - 1)];
    }
    __syncthreads();

    const int nslab = (K + GK - 1) / GK;
    auto issue = [&](int slab) {
        double *st = sm + (size_t)(slab % GSTG) * GSTAGE;
        const int kc = slab * GK;
#pragma unroll
        for (int u = 0; u < 4; ++u) {          // A: 2 planes x 64 rows x 8 chunks
            const int e = tid + u * 256, pl = e >> 9, r = (e >> 3) & 63, kp = (e & 7) * 2;
            const bool ok = kc + kp < K;
            const double *src = (pl ? wim : wre) + (size_t)s_ra[r] * lw + ka + (ok ? kc + kp : 0);
            cp_async16(st + pl * (TS * GLDA) + r * GLDA + kp, src, ok);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {          // B: 2 planes x 16 k x 32 chunks
            const int e = tid + u * 256, pl = e >> 9, k = (e >> 5) & 15, cp = (e & 31) * 2;
            const bool ok = kc + k < K && tc + cp < c1;
            const double *src = (pl ? wim : wre) + (size_t)s_rb[ok ? kc + k : 0] * lw + (ok ? tc + cp : tc);
            cp_async16(st + 2 * (TS * GLDA) + pl * (GK * LDS_T) + k * LDS_T + cp, src, ok);
        }
    };
#pragma unroll
    for (int s0 = 0; s0 < GSTG - 1; ++s0) {
        if (s0 < nslab) issue(s0);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    double2 cre[4][2], cim[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int row = s_ra[wr * 32 + i * 8 + fr], col = tc + wc * 16 + j * 8 + 2 * fk;
            cre[i][j] = cim[i][j] = make_double2(0.0, 0.0);
            if (col < c1) {
                cre[i][j] = *reinterpret_cast<const double2 *>(wre + (size_t)row * lw + col);
                cim[i][j] = *reinterpret_cast<const double2 *>(wim + (size_t)row * lw + col);
            }
        }

    for (int sl = 0; sl < nslab; ++sl) {
        asm volatile("cp.async.wait_group %0;" ::"n"(GSTG - 2) : "memory");
        __syncthreads();          // slab sl has landed for everyone, and everyone is done with the stage refilled next
        const double *Ar = sm + (size_t)(sl % GSTG) * GSTAGE, *Ai = Ar + TS * GLDA, *Br = Ai + TS * GLDA, *Bi = Br + GK * LDS_T;
#pragma unroll
        for (int k4 = 0; k4 < GK; k4 += 4) {
            double nar[4], pai[4], nai[4], br[2], bi[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int o = (wr * 32 + i * 8 + fr) * GLDA + k4 + fk;
                nar[i] = -Ar[o];
                pai[i] = Ai[o];
                nai[i] = -pai[i];
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int o = (k4 + fk) * LDS_T + wc * 16 + j * 8 + fr;
                br[j] = Br[o];
                bi[j] = Bi[o];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    dmma(cre[i][j].x, cre[i][j].y, nar[i], br[j]);
                    dmma(cre[i][j].x, cre[i][j].y, pai[i], bi[j]);
                    dmma(cim[i][j].x, cim[i][j].y, nar[i], bi[j]);
                    dmma(cim[i][j].x, cim[i][j].y, nai[i], br[j]);
                }
            // the refill of the stage freed by the barrier is issued by warp w after its k-step (w mod 4): staggered, some warp
            // of every scheduler is always issuing DMMAs while another does the address arithmetic
            if (k4 / 4 == (((warp & 3) + 2 * (warp >> 2)) & 3)) {
                if (sl + GSTG - 1 < nslab) issue(sl + GSTG - 1);
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int row = s_ra[wr * 32 + i * 8 + fr], col = tc + wc * 16 + j * 8 + 2 * fk;
            if (col < c1) {
                *reinterpret_cast<double2 *>(wre + (size_t)row * lw + col) = cre[i][j];
                *reinterpret_cast<double2 *>(wim + (size_t)row * lw + col) = cim[i][j];
            }
        }
}

// keep G^a[:,sel] (pass 0 of the biased power spectrum) while the second factorisation runs; Xs is indexed by position
__global__ void k_save(Geo g, const double *W, const int *act, double *Xs) {
    const int b = blockIdx.y;
    const double *wre = W + (size_t)b * 2 * g.plane, *wim = wre + g.plane;
    const int *actb = act + (size_t)b * g.nrp;
    double *xr = Xs + (size_t)b * 2 * g.np * g.nrhs, *xi = xr + (size_t)g.np * g.nrhs;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < g.np * g.nrhs; e += gridDim.x * blockDim.x) {
        const int c = e % g.nrhs, i = e / g.nrhs;
        xr[e] = wre[(size_t)actb[i] * g.lw + g.np + c];
        xi[e] = wim[(size_t)actb[i] * g.lw + g.np + c];
    }
}

// the full solution X = M^-1 (mode 3: bpt.retargf / bpt.advangf): out[b][i][j] = (re, im) of row i of unit column j
__global__ void k_extract(Geo g, const double *W, const int *act, double *out, int *status, int w0) {
    const int b = blockIdx.y, n = g.nl;
    const double *wre = W + (size_t)b * 2 * g.plane, *wim = wre + g.plane;
    const int *actb = act + (size_t)b * g.nrp;
    double2 *o = reinterpret_cast<double2 *>(out) + (size_t)b * n * n;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) {
        const int j = e % n, i = e / n;
        const size_t src = (size_t)actb[i] * g.lw + g.np + j;
        const double2 v = make_double2(wre[src], wim[src]);
        o[e] = v;
        if (!isfinite(v.x) || !isfinite(v.y)) status[w0 + b] = 1;
    }
}

// ------------------------------------------------------------------------------------------------ observable
__global__ void __launch_bounds__(256) k_observe(Geo g, Problem p, const double *W, const int *act, const double *Xs, double *out, int *status,
                                                 int w0, int mode) {
    __shared__ double red[32];
    const int b = blockIdx.x, iw = w0 + b, lw = g.lw, n = g.np;
    const double *wre = W + (size_t)b * 2 * g.plane, *wim = wre + g.plane;
    const int *actb = act + (size_t)b * g.nrp;
    const double w = p.omegas[iw];
    double acc = 0.0;
    if (mode == 2) {
        // w^2 Re sum_c [G^r Sigma^K G^a]_cc  with  Z[a,c] = G^r[sel_c,a] (pass 1, in W) and X[b,c] = G^a[b,sel_c] (pass 0, in Xs)
        const double *xr = Xs + (size_t)b * 2 * n * g.nrhs, *xi = xr + (size_t)n * g.nrhs;
        const double kd = p.kd[iw], kr1 = p.kr1[iw], kr2 = p.kr2[iw], ki = p.ki[iw];
        for (int e = threadIdx.x; e < g.nl * g.nrhs; e += blockDim.x) {          // diagonal lead part: kd * mask_a
            const int c = e % g.nrhs, i = e / g.nrhs;
            const double m = p.mask[i];
            if (m != 0.0) {
                const size_t o = (size_t)actb[i] * lw + n + c;
                acc += kd * m * (wre[o] * xr[(size_t)i * g.nrhs + c] - wim[o] * xi[(size_t)i * g.nrhs + c]);
            }
        }
        for (int e = threadIdx.x; e < p.nb * p.nb * g.nrhs; e += blockDim.x) { // dense bias block
            const int c = e % g.nrhs, ab = e / g.nrhs, ia = ab / p.nb, ib = ab % p.nb;
            const double skr = kr1 * p.bdamp[ab] + kr2 * p.chiplus[ab], ski = ki * p.chiminus[ab];
            const size_t ra = (size_t)actb[p.bpos[ia]] * lw + n + c, rb = (size_t)p.bpos[ib] * g.nrhs + c;
            const double zr2 = wre[ra], zi2 = wim[ra], xr2 = xr[rb], xi2 = xi[rb];
            const double tr = skr * xr2 - ski * xi2, ti = skr * xi2 + ski * xr2;     // Re[ z * sk * x ]
            acc += zr2 * tr - zi2 * ti;
        }
    } else if (mode == 0) {
        for (int e = threadIdx.x; e < p.nrows * g.nrhs; e += blockDim.x) {
            const int c = e % g.nrhs, i = actb[p.rows[e / g.nrhs]];
            const double xr = wre[(size_t)i * lw + n + c], xi = wim[(size_t)i * lw + n + c];
            acc += xr * xr + xi * xi;
        }
    } else {
        for (int c = threadIdx.x; c < g.nrhs; c += blockDim.x) acc += wim[(size_t)actb[p.rhs[c]] * lw + n + c];
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
        const double gam = 2.0 * w / p.damp;
        out[iw] = mode == 0 ? gam * gam * acc : (mode == 1 ? -2.0 * w * w * p.weight[iw] * acc : w * w * acc);
        if (!isfinite(acc)) status[iw] = 1;      // a zero / NaN pivot somewhere upstream: reported as a singular matrix
    }
}


// Workspace of a sclmd_bpt handle, kept between its sweeps (cudaMalloc of several GB costs more than a whole sweep): grow-only raw
// buffers and the streams of the batch slots; freed by sclmd_bpt_destroy.  One sweep at a time per handle (its mutex).
struct RawBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t want) {
        if (want <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) bytes = want; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};
struct StreamSlot {
    cudaStream_t st = nullptr;
    RawBuf W, Xs, act, flags;
};
constexpr int NSLOT = 6;
struct Workspace {
    int device = -1;
    StreamSlot slot[NSLOT];
    RawBuf arena;
    void release() {
        for (auto &s : slot) {
            s.W.release(); s.Xs.release(); s.act.release(); s.flags.release();
            if (s.st) cudaStreamDestroy(s.st);
            s.st = nullptr;
        }
        arena.release();
    }
};
// Optional per-kernel-class timing of a handle (sclmd_bpt_set_profiling): one stream, a CUDA event pair around every launch.
// classes: 0 build, 1 panel, 2 panel-column update (k_gemm, rank 16), 3 block trsm, 4 trailing update (k_gemm, rank 64),
//          5 back substitution, 6 observable/save
constexpr int NPROF = 7;
struct Prof {
    bool on = false;
    std::vector<cudaEvent_t> ev;      // pairs
    std::vector<int> kind;
    double ms[NPROF] = {0};
    long long n[NPROF] = {0};
    double last_device_ms = 0.0;
    unsigned long long tiles = 0;     // rank-64 update tiles that ran the tensor loop (zero tiles return early)
    double gemm_flops = 0.0;
    void begin(int k, cudaStream_t st) {
        if (!on) return;
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st);
        ev.push_back(a); ev.push_back(b); kind.push_back(k);
    }
    void end(cudaStream_t st) { if (on) cudaEventRecord(ev.back(), st); }
    void collect() {
        for (size_t i = 0; i < kind.size(); ++i) {
            float t = 0;
            if (cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]) == cudaSuccess) { ms[kind[i]] += t; n[kind[i]]++; }
            cudaEventDestroy(ev[2 * i]); cudaEventDestroy(ev[2 * i + 1]);
        }
        ev.clear(); kind.clear();
    }
};
__global__ void k_permute_k(const double *__restrict__ K, const int *__restrict__ order, int n, double *__restrict__ Kp) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * n) return;
    const int i = (int)(e / n), j = (int)(e % n);
    Kp[e] = K[(size_t)order[i] * n + order[j]];
}

}  // namespace

// The handle of the frequency sweeps: the junction (K on the device, lead dofs, damping, optional bias block), the batch workspace,
// the streams and the profiling state.  Nothing of the sweeps lives outside it.
struct sclmd_bpt {
    int device = 0, n = 0;
    double damp = 0.0;
    DevBuf<double> dK;                  // K in the caller's numbering, uploaded once
    std::vector<int> idxL, idxR;
    std::vector<double> mask;           // number of leads a dof belongs to
    int b0 = 0, nb = 0;
    std::vector<double> bdamp, chiplus, chiminus;
    double bias = 0.0;
    Workspace ws;
    Prof prof;
    std::mutex mu;
};

namespace {

// enqueue the factorisation + solve of one batch (nbat frequencies starting at w0) on slot s
cudaError_t enqueue_lu(Prof &pf, const Geo &g, const Problem &p, StreamSlot &s, int *status, int w0, int nbat, int row_stop, double sgn, int tblk,
                       unsigned long long *tiles) {
    cudaStream_t st = s.st;
    double *W = static_cast<double *>(s.W.p);
    int *act = static_cast<int *>(s.act.p), *flags = static_cast<int *>(s.flags.p);
    pf.begin(0, st);
    k_build<<<dim3(g.nrp, nbat), 128, 0, st>>>(g, p, W, act, w0, sgn, tblk);
    pf.end(st);
    for (int k0 = 0; k0 < g.np; k0 += TS) {       // (blocks start on multiples of 64: tiles stay 512-byte aligned)
        const int rend = std::min(k0 + TS, g.np);
        // rows per thread x sub-panel width is bounded by the register file (32 complex numbers per thread): the taller the
        // panel, the narrower its sub-panels; k_block_trsm and the rank-64 update do not depend on that width
        const int mh = g.np - k0;
        const int nbw = mh > 2048 ? 2 : (mh > 1024 ? 4 : (mh > 512 ? 8 : 16));
        for (int kk = k0; kk < rend; kk += nbw) {
            const int m = g.np - kk, kb = std::min(nbw, rend - kk);
            pf.begin(1, st);
            if (nbw == 2) k_panel<256, 16, 2><<<nbat, 256, 0, st>>>(g, W, act, status, w0, kk, kb, rend);
            else if (nbw == 4) k_panel<256, 8, 4><<<nbat, 256, 0, st>>>(g, W, act, status, w0, kk, kb, rend);
            else if (nbw == 8) k_panel<512, 2, 8><<<nbat, 512, 0, st>>>(g, W, act, status, w0, kk, kb, rend);
            else if (m > 256) k_panel<256, 2, 16><<<nbat, 256, 0, st>>>(g, W, act, status, w0, kk, kb, rend);
            else if (m > 128) k_panel<256, 1, 16><<<nbat, 256, 0, st>>>(g, W, act, status, w0, kk, kb, rend);
            else if (m > 64) k_panel<128, 1, 16><<<nbat, 128, 0, st>>>(g, W, act, status, w0, kk, kb, rend);
            else k_panel<64, 1, 16><<<nbat, 64, 0, st>>>(g, W, act, status, w0, kk, kb, rend);
            pf.end(st);
            // the remaining panel columns of this block, every row below the sub-panel: rank-kb update
            if (kk + kb < rend) {
                pf.begin(2, st);
                k_gemm<<<dim3(1, cdiv(g.np - (kk + kb), TS), nbat), 256, GEMM_SMEM, st>>>(g, W, act, nullptr, nullptr, kk + kb, kk + kb, rend, kk, kb);
                pf.end(st);
            }
        }
        if (g.ncols > rend) {
            pf.begin(3, st);
            k_block_trsm<16><<<nbat, 256, TRSM_SMEM, st>>>(g, W, act, flags, k0, rend);      // its own 16-row blocking of L11
            pf.end(st);
        }
        if (rend < g.np) {     // trailing matrix and carried right-hand sides: rank-64 update
            pf.begin(4, st);
            k_gemm<<<dim3(cdiv(g.ncols - rend, TS), cdiv(g.np - rend, TS), nbat), 256, GEMM_SMEM, st>>>(g, W, act, flags, tiles, rend, rend, g.ncols, k0, rend - k0);
            pf.end(st);
        }
    }
    // back substitution for the rows >= row_stop, 64-row blocks from the bottom: triangular solve of the block (k_block_trsm_upper),
    // then the rows above it on the tensor pipe (k_gemm, rank = block height)
    pf.begin(5, st);
    const int rs = (row_stop / TS) * TS;
    for (int b0 = ((g.np - 1) / TS) * TS; b0 >= rs; b0 -= TS) {
        const int b1 = std::min(b0 + TS, g.np);
        k_block_trsm_upper<<<nbat, 256, TRSM_SMEM, st>>>(g, W, act, b0, b1);
        if (b0 > rs)
            k_gemm<<<dim3(cdiv(g.nrhs, TS), (b0 - rs) / TS, nbat), 256, GEMM_SMEM, st>>>(g, W, act, nullptr, nullptr, rs, g.np, g.ncols, b0, b1 - b0);
    }
    pf.end(st);
    return cudaGetLastError();
}

struct Keldysh {
    const double *kd = nullptr, *kr1 = nullptr, *kr2 = nullptr, *ki = nullptr;
};

int run_lu(sclmd_bpt &h, const double *omegas, int nw, int mode, const double *weight, const int32_t *sel, int nsel, double *out,
           const Keldysh *kw = nullptr, int advanced = 0) {
    // mode 0 transmission, 1 power spectrum, 2 biased power spectrum, 3 the full Green function (out = [nw][n][n] complex, interleaved)
    SCLMD_REQUIRE(omegas && nw > 0 && out, "bpt: bad arguments");
    std::lock_guard<std::mutex> lock(h.mu);
    const int n = h.n, device = h.device, nb = h.nb, b0 = h.b0;
    const double damp = h.damp;
    if (int e = select_device(device)) return e;
    SCLMD_REQUIRE(mode != 2 || (nb > 0 && kw && kw->kd && kw->kr1 && kw->kr2 && kw->ki), "bpt.ps (biased): no bias block set or missing Keldysh weights");
    const std::vector<double> &mask = h.mask;
    // right-hand sides and the rows of the solution the observable reads (original numbering)
    std::vector<int> rhs_o, rows_o;
    if (mode == 0) {
        rhs_o = h.idxL;
        rows_o = h.idxR;
    } else if (mode == 3) {
        rhs_o.resize(n);
        std::iota(rhs_o.begin(), rhs_o.end(), 0);
        rows_o = rhs_o;
    } else {
        SCLMD_REQUIRE(sel && nsel > 0 && (weight || mode == 2), "bpt.ps: empty selection");
        for (int i = 0; i < nsel; ++i) SCLMD_REQUIRE(sel[i] >= 0 && sel[i] < n, "bpt.ps: selected dof %d out of range", sel[i]);
        rhs_o.assign(sel, sel + nsel);
        rows_o = rhs_o;
        if (mode == 2) {   // the Keldysh self-energy lives on the lead dofs and on the bias block: those rows of both solutions
            rows_o.clear();
            for (int i = 0; i < n; ++i)
                if (mask[i] != 0.0 || (i >= b0 && i < b0 + nb)) rows_o.push_back(i);
        }
    }
    // ordering [neither | right-hand-side dofs | needed rows]: unit columns stay zero above their dof, back substitution is short
    std::vector<int> key(n, 0), order(n), pos(n);
    for (int d : rhs_o) key[d] = std::max(key[d], 1);
    for (int d : rows_o) key[d] = 2;
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
    for (int i = 0; i < n; ++i) pos[order[i]] = i;
    std::vector<double> maskp(n);
    std::vector<int> bmap(n, -1), bpos(std::max(nb, 1), 0);
    for (int i = 0; i < n; ++i) {
        maskp[i] = mask[order[i]];
        if (order[i] >= b0 && order[i] < b0 + nb) { bmap[i] = order[i] - b0; bpos[order[i] - b0] = i; }
    }
    std::vector<int> rhs(rhs_o.size()), rows(rows_o.size());
    for (size_t i = 0; i < rhs_o.size(); ++i) rhs[i] = pos[rhs_o[i]];
    for (size_t i = 0; i < rows_o.size(); ++i) rows[i] = pos[rows_o[i]];
    const int row_stop = *std::min_element(rows.begin(), rows.end());

    Geo g{};
    g.nl = n; g.np = n + (n & 1); g.nrhs = (int)rhs.size(); g.ncols = g.np + g.nrhs;
    g.nrp = round_up(g.np, TS) + TS; g.lw = round_up(g.ncols, TS); g.plane = (size_t)g.nrp * g.lw;

    // batches: several in flight on separate streams; slot memory bounded
    const int sms = sm_count(device);
    const size_t per_w = 2 * g.plane * sizeof(double);
    // A/B switches: SCLMD_BPT_BATCH = frequencies per batch in units of the SM count (default 4), SCLMD_BPT_SLOTS = batches in flight (default 3)
    static const int batch_mult = getenv("SCLMD_BPT_BATCH") ? std::max(1, atoi(getenv("SCLMD_BPT_BATCH"))) : 4;
    static const int slots_max = getenv("SCLMD_BPT_SLOTS") ? std::min(NSLOT, std::max(1, atoi(getenv("SCLMD_BPT_SLOTS")))) : 3;
    int bmax = batch_mult * sms;
    while (bmax > 8 && (size_t)bmax * per_w * slots_max > ((size_t)24 << 30)) bmax /= 2;
    const int nbatch = cdiv(nw, bmax), bsz = cdiv(nw, nbatch);
    Prof &prof = h.prof;
    const int nslots = prof.on ? 1 : std::min(slots_max, nbatch);

    Workspace &ws = h.ws;
    StreamSlot *slots = ws.slot;
    // the per-call device arrays come out of one cached arena (no cudaMalloc / cudaFree per sweep)
    const size_t nb2 = (size_t)nb * nb;
    const size_t arena_bytes = sizeof(double) * ((size_t)n * n + n + 7 * (size_t)nw + 3 * nb2) +
                               sizeof(int) * (rhs.size() + rows.size() + (size_t)nw + 2 * (size_t)n + bpos.size()) + 256 * 24;
    SCLMD_CUDA(ws.arena.reserve(((arena_bytes >> 24) + 1) << 24));      // 16 MB granules: sweeps of similar length reuse the arena
    char *cursor = static_cast<char *>(ws.arena.p);
    auto take = [&](size_t bytes) { char *r = cursor; cursor += (bytes + 255) / 256 * 256; return static_cast<void *>(r); };
    auto take_d = [&](size_t cnt) { return static_cast<double *>(take(cnt * sizeof(double))); };
    auto take_i = [&](size_t cnt) { return static_cast<int *>(take(cnt * sizeof(int))); };
    double *dK = take_d((size_t)n * n), *dmask = take_d(n), *dom = take_d(nw), *dwt = take_d(nw), *dout = take_d(nw);
    int *drhs = take_i(rhs.size()), *drows = take_i(rows.size()), *dstat = take_i(nw), *dbmap = take_i(n), *dbpos = take_i(bpos.size());
    int *dorder = take_i(n);
    // K stays on the device with the handle; only its re-ordering for this observable is rebuilt (a device gather)
    SCLMD_CUDA(cudaMemcpy(dorder, order.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    k_permute_k<<<(unsigned)(((size_t)n * n + 255) / 256), 256>>>(h.dK.p, dorder, n, dK);
    SCLMD_CUDA(cudaGetLastError());
    SCLMD_CUDA(cudaMemcpy(dmask, maskp.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(dom, omegas, nw * sizeof(double), cudaMemcpyHostToDevice));
    if (weight) SCLMD_CUDA(cudaMemcpy(dwt, weight, nw * sizeof(double), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(drhs, rhs.data(), rhs.size() * sizeof(int), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(drows, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(dbmap, bmap.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(dbpos, bpos.data(), bpos.size() * sizeof(int), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemset(dstat, 0, nw * sizeof(int)));
    Problem p{};
    p.K = dK; p.mask = dmask; p.rhs = drhs; p.rows = drows; p.nrows = (int)rows.size(); p.bmap = dbmap; p.bpos = dbpos;
    p.nb = nb; p.damp = damp; p.eps = 1e-9; p.omegas = dom; p.weight = dwt;
    if (nb > 0) {
        double *dbd = take_d(nb2), *dcp = take_d(nb2), *dcm = take_d(nb2);
        SCLMD_CUDA(cudaMemcpy(dbd, h.bdamp.data(), nb2 * sizeof(double), cudaMemcpyHostToDevice));
        SCLMD_CUDA(cudaMemcpy(dcp, h.chiplus.data(), nb2 * sizeof(double), cudaMemcpyHostToDevice));
        SCLMD_CUDA(cudaMemcpy(dcm, h.chiminus.data(), nb2 * sizeof(double), cudaMemcpyHostToDevice));
        p.bdamp = dbd; p.chiplus = dcp; p.chiminus = dcm; p.bias = h.bias;
        if (mode == 2) {
            double *dkw = take_d((size_t)4 * nw);
            const double *src[4] = {kw->kd, kw->kr1, kw->kr2, kw->ki};
            for (int q = 0; q < 4; ++q) SCLMD_CUDA(cudaMemcpy(dkw + (size_t)q * nw, src[q], nw * sizeof(double), cudaMemcpyHostToDevice));
            p.kd = dkw; p.kr1 = dkw + nw; p.kr2 = dkw + 2 * (size_t)nw; p.ki = dkw + 3 * (size_t)nw;
        }
    }
    for (int s = 0; s < nslots; ++s) {
        if (!slots[s].st) SCLMD_CUDA(cudaStreamCreateWithFlags(&slots[s].st, cudaStreamNonBlocking));
        SCLMD_CUDA(slots[s].W.reserve((size_t)bsz * 2 * g.plane * sizeof(double)));       // k_build writes every element
        SCLMD_CUDA(slots[s].act.reserve((size_t)bsz * g.nrp * sizeof(int)));
        SCLMD_CUDA(slots[s].flags.reserve((size_t)bsz * (g.lw / TC) * sizeof(int)));
        if (mode == 2) SCLMD_CUDA(slots[s].Xs.reserve((size_t)bsz * 2 * g.np * g.nrhs * sizeof(double)));
    }
    SCLMD_CUDA(cudaFuncSetAttribute(k_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
    SCLMD_CUDA(cudaFuncSetAttribute(k_block_trsm<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_SMEM));
    SCLMD_CUDA(cudaFuncSetAttribute(k_block_trsm_upper, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_SMEM));

    DevBuf<unsigned long long> dtiles;
    if (prof.on) SCLMD_CUDA(dtiles.alloc(1));
    DevBuf<double> dgreen;
    if (mode == 3) {
        SCLMD_REQUIRE((size_t)nw * n * n <= ((size_t)1 << 28), "bpt: %d full Green functions of order %d do not fit the output buffer; sweep in pieces", nw, n);
        SCLMD_CUDA(dgreen.alloc((size_t)nw * n * n * 2));
    }
    const bool want_timing = getenv("SCLMD_BPT_TIMING") != nullptr;
    cudaEvent_t e0, e1;
    SCLMD_CUDA(cudaEventCreate(&e0));
    SCLMD_CUDA(cudaEventCreate(&e1));
    SCLMD_CUDA(cudaDeviceSynchronize());
    SCLMD_CUDA(cudaEventRecord(e0, slots[0].st));
    for (int ib = 0; ib < nbatch; ++ib) {
        StreamSlot &s = slots[ib % nslots];
        const int w0 = ib * bsz, nbat = std::min(bsz, nw - w0);
        if (nbat <= 0) break;
        if (mode == 2) {
            SCLMD_CUDA(enqueue_lu(prof, g, p, s, dstat, w0, nbat, row_stop, -1.0, 1, dtiles.p));
            k_save<<<dim3(cdiv(g.np * g.nrhs, 256), nbat), 256, 0, s.st>>>(g, static_cast<const double *>(s.W.p), static_cast<const int *>(s.act.p), static_cast<double *>(s.Xs.p));
            SCLMD_CUDA(enqueue_lu(prof, g, p, s, dstat, w0, nbat, row_stop, 1.0, 1, dtiles.p));
        } else if (mode == 3) {
            SCLMD_CUDA(enqueue_lu(prof, g, p, s, dstat, w0, nbat, row_stop, advanced ? -1.0 : 1.0, advanced ? 1 : 0, dtiles.p));
            k_extract<<<dim3(cdiv(n * n, 256), nbat), 256, 0, s.st>>>(g, static_cast<const double *>(s.W.p), static_cast<const int *>(s.act.p),
                                                                      dgreen.p + (size_t)w0 * n * n * 2, dstat, w0);
            SCLMD_CUDA(cudaGetLastError());
            continue;
        } else {
            SCLMD_CUDA(enqueue_lu(prof, g, p, s, dstat, w0, nbat, row_stop, 1.0, 0, dtiles.p));
        }
        prof.begin(6, s.st);
        k_observe<<<nbat, 256, 0, s.st>>>(g, p, static_cast<const double *>(s.W.p), static_cast<const int *>(s.act.p), static_cast<const double *>(s.Xs.p), dout, dstat, w0, mode);
        prof.end(s.st);
        SCLMD_CUDA(cudaGetLastError());
    }
    // the end marker waits for every slot
    for (int s = 1; s < nslots; ++s) {
        cudaEvent_t ev;
        SCLMD_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        SCLMD_CUDA(cudaEventRecord(ev, slots[s].st));
        SCLMD_CUDA(cudaStreamWaitEvent(slots[0].st, ev, 0));
        SCLMD_CUDA(cudaEventDestroy(ev));
    }
    SCLMD_CUDA(cudaEventRecord(e1, slots[0].st));
    SCLMD_CUDA(cudaDeviceSynchronize());
    float kms = 0;
    SCLMD_CUDA(cudaEventElapsedTime(&kms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    prof.last_device_ms = kms;
    if (prof.on) {
        prof.collect();
        unsigned long long t = 0;
        SCLMD_CUDA(cudaMemcpy(&t, dtiles.p, sizeof(t), cudaMemcpyDeviceToHost));
        prof.tiles += t;
        prof.gemm_flops += (double)t * 8.0 * TS * TS * TS;      // full 64-deep tiles (the last block of an n that is not a multiple of 64 is shallower)
    }
    if (want_timing)
        fprintf(stderr, "[bpt] device %.3f ms for %d frequencies in %d batches of %d on %d streams (%.0f omega/s device-only)\n", kms, nw,
                nbatch, bsz, nslots, nw / (kms * 1e-3));
    if (mode == 3) SCLMD_CUDA(cudaMemcpy(out, dgreen.p, (size_t)nw * n * n * 2 * sizeof(double), cudaMemcpyDeviceToHost));
    else SCLMD_CUDA(cudaMemcpy(out, dout, nw * sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<int> stv(nw);
    SCLMD_CUDA(cudaMemcpy(stv.data(), dstat, nw * sizeof(int), cudaMemcpyDeviceToHost));
    for (int i = 0; i < nw; ++i)
        if (stv[i]) {
            set_error("bpt: singular matrix at omega[%d]=%g (numpy.linalg.LinAlgError in the reference)", i, omegas[i]);
            return SCLMD_ERR_STATE;
        }
    return SCLMD_OK;
}

}  // namespace

extern "C" {

// bpt.__init__ / getdynmat (negf.py:8-25, 39-102) leave the reduced dynamical matrix, the lead dofs and the damping: that is the handle.
// K is uploaded once; every sweep of the handle reuses it, the batch workspace and the streams.
int sclmd_bpt_create(int device, int n, const double *K, const int32_t *idxL, int nL, const int32_t *idxR, int nR, double damp,
                     sclmd_bpt **out) {
    SCLMD_REQUIRE(out, "sclmd_bpt_create: NULL output");
    *out = nullptr;
    SCLMD_REQUIRE(n > 0 && K && idxL && idxR && nL > 0 && nR > 0 && damp != 0.0, "sclmd_bpt_create: bad arguments");
    SCLMD_REQUIRE(n <= 4096, "bpt: n=%d exceeds the register-resident panel (n <= 4096)", n);
    if (int e = select_device(device)) return e;
    std::unique_ptr<sclmd_bpt> h(new sclmd_bpt());
    h->device = device; h->n = n; h->damp = damp;
    h->mask.assign(n, 0.0);
    for (int i = 0; i < nL; ++i) {
        SCLMD_REQUIRE(idxL[i] >= 0 && idxL[i] < n, "bpt: left bath dof %d out of range", idxL[i]);
        h->mask[idxL[i]] += 1.0;
    }
    for (int i = 0; i < nR; ++i) {
        SCLMD_REQUIRE(idxR[i] >= 0 && idxR[i] < n, "bpt: right bath dof %d out of range", idxR[i]);
        h->mask[idxR[i]] += 1.0;
    }
    h->idxL.assign(idxL, idxL + nL);
    h->idxR.assign(idxR, idxR + nR);
    SCLMD_CUDA(h->dK.alloc_raw((size_t)n * n));
    SCLMD_CUDA(cudaMemcpy(h->dK.p, K, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice));
    *out = h.release();
    return SCLMD_OK;
}

int sclmd_bpt_destroy(sclmd_bpt *h) {
    if (!h) return SCLMD_OK;
    {
        std::lock_guard<std::mutex> lock(h->mu);
        if (cudaSetDevice(h->device) == cudaSuccess) {
            cudaDeviceSynchronize();
            h->ws.release();
        }
        h->prof.collect();
    }
    delete h;
    return SCLMD_OK;
}

// bpt.setbias (negf.py:27-37): Sigma_b^r = -i w bdamp - bias chiminus on the contiguous dof block [b0, b0+nb) (negf.py:162-172);
// bias in angular units (eV/hbar, as bpt stores it); nb == 0 removes the block
int sclmd_bpt_set_bias(sclmd_bpt *h, int b0, int nb, const double *bdamp, const double *chiplus, const double *chiminus, double bias) {
    SCLMD_REQUIRE(h, "sclmd_bpt_set_bias: NULL handle");
    std::lock_guard<std::mutex> lock(h->mu);
    if (nb == 0) {
        h->b0 = h->nb = 0; h->bias = 0.0;
        h->bdamp.clear(); h->chiplus.clear(); h->chiminus.clear();
        return SCLMD_OK;
    }
    SCLMD_REQUIRE(nb > 0 && b0 >= 0 && b0 + nb <= h->n && bdamp && chiplus && chiminus, "bpt: bad bias block");
    const size_t nb2 = (size_t)nb * nb;
    h->b0 = b0; h->nb = nb; h->bias = bias;
    h->bdamp.assign(bdamp, bdamp + nb2);
    h->chiplus.assign(chiplus, chiplus + nb2);
    h->chiminus.assign(chiminus, chiminus + nb2);
    return SCLMD_OK;
}

// per-kernel-class timing of this handle's sweeps (one stream, event pair per launch); switching it on resets the totals
int sclmd_bpt_set_profiling(sclmd_bpt *h, int on) {
    SCLMD_REQUIRE(h, "sclmd_bpt_set_profiling: NULL handle");
    std::lock_guard<std::mutex> lock(h->mu);
    h->prof.collect();
    h->prof = Prof();
    h->prof.on = on != 0;
    return SCLMD_OK;
}

// ms[7], n[7]: build, panel, rank-16 panel-column update, block trsm, rank-64 trailing update, back substitution, observable;
// gemm_flops: flops executed by the rank-64 update tiles; last_device_ms: device time of the most recent sweep (always kept)
int sclmd_bpt_get_profile(sclmd_bpt *h, double *ms, int64_t *n, double *gemm_flops, double *last_device_ms) {
    SCLMD_REQUIRE(h, "sclmd_bpt_get_profile: NULL handle");
    std::lock_guard<std::mutex> lock(h->mu);
    for (int i = 0; i < NPROF; ++i) {
        if (ms) ms[i] = h->prof.ms[i];
        if (n) n[i] = h->prof.n[i];
    }
    if (gemm_flops) *gemm_flops = h->prof.gemm_flops;
    if (last_device_ms) *last_device_ms = h->prof.last_device_ms;
    return SCLMD_OK;
}

// frees the device scratch the noise generator caches between calls (the sweeps' workspaces belong to their handles)
int sclmd_release_workspace(void) {
    sclmd::release_noise_scratch();
    return SCLMD_OK;
}

// bpt.tm (negf.py:240-242); with a bias block set G includes its retarded self-energy
int sclmd_bpt_tm(sclmd_bpt *h, const double *omegas, int nw, double *tm_out) {
    SCLMD_REQUIRE(h, "sclmd_bpt_tm: NULL handle");
    return run_lu(*h, omegas, nw, 0, nullptr, nullptr, 0, tm_out);
}

// bpt.ps without bias (negf.py:232)
int sclmd_bpt_ps(sclmd_bpt *h, const double *omegas, const double *nb, int nw, const int32_t *sel, int nsel, double *ps_out) {
    SCLMD_REQUIRE(h, "sclmd_bpt_ps: NULL handle");
    SCLMD_REQUIRE(h->nb == 0, "sclmd_bpt_ps: the handle carries a bias block; use sclmd_bpt_ps_bias");
    return run_lu(*h, omegas, nw, 1, nb, sel, nsel, ps_out);
}

// bpt.retargf / bpt.advangf (negf.py:206-212): the full Green function of every frequency, G[nw][n][n] complex (interleaved re, im).
// advangf keeps the +i eps of z (negf.py:212) and takes the advanced self-energies.
int sclmd_bpt_green(sclmd_bpt *h, const double *omegas, int nw, int advanced, double *green_out) {
    SCLMD_REQUIRE(h, "sclmd_bpt_green: NULL handle");
    return run_lu(*h, omegas, nw, 3, nullptr, nullptr, 0, green_out, nullptr, advanced);
}

// bpt.ps with bias (negf.py:234-236): w^2 Re Tr[(G^r Sigma^K G^a)[sel,sel]], Sigma^K = totalkselfenergy (negf.py:177-193)
//   = kd(w) on the lead dofs (kd = (2w/damp) n_B) + kr1 bdamp + kr2 chiplus + i ki chiminus on the bias block
int sclmd_bpt_ps_bias(sclmd_bpt *h, const double *omegas, const double *kd, const double *kr1, const double *kr2, const double *ki, int nw,
                      const int32_t *sel, int nsel, double *ps_out) {
    SCLMD_REQUIRE(h, "sclmd_bpt_ps_bias: NULL handle");
    Keldysh kw;
    kw.kd = kd; kw.kr1 = kr1; kw.kr2 = kr2; kw.ki = ki;
    return run_lu(*h, omegas, nw, 2, nullptr, sel, nsel, ps_out, &kw);
}

}  // extern "C"
