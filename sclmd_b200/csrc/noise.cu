// Quantum coloured-noise generation for an ensemble (replaces noise.phnoise / noise.enoise /
// noise.vargau / myfft.iFourier1D; sclmd/noise.py:50-100,149-206,273-305, functions.py:36-53).
//
//   1. covariance  A_w = sum_m (cre + i cim)[w][m] * basis[idx[w][m]], hermitianised
//      (phnoise: two neighbouring gamma-grid nodes from flinterp, functions.py:117-134;
//       enoise : efric, exip, exim with the Bose weights of noise.py:172-185)
//   2. factor      L_w = V sqrt(clamp+(lambda))  -- one-sided Jacobi, one CTA per frequency
//      (same clamp as vargau: eigenvalues <= 0 contribute nothing, noise.py:299-303)
//   3. draws       xi ~ N(0,1) from Philox4x32-10 + Box-Muller, counter = (w, k, trajectory)
//   4. x_w = L_w xi_w  as a batched DMMA GEMM over frequencies
//   5. mirror to negative frequencies (noise.py:87-94) and Fourier transform with in-house
//      Stockham radix-2/3/4/5 kernels (no cuFFT); two real-output series share one complex
//      transform.  series = Re(FFT)/(dt*nmd)  (functions.py:51-53, baths.py:191,408)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <memory>
#include <mutex>

#include "common.cuh"
#include "dgemm.cuh"
#include "dgemm_tma.cuh"

using namespace sclmd;

namespace {

// ------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
// two independent standard normals from one Philox4x32-10 block (Box-Muller: both branches): counter = (w, column pair, trajectory)
__device__ __forceinline__ double2 philox_normal2(uint64_t seed, uint32_t w, uint32_t kpair, uint64_t traj) {
    uint32_t c[4] = {w, kpair, (uint32_t)traj, (uint32_t)(traj >> 32)};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    const uint64_t a = ((uint64_t)c[0] << 32) | c[1], b = ((uint64_t)c[2] << 32) | c[3];
    const double u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);  // (0,1)
    const double u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    return make_double2(rad * cs, rad * sn);
}

// xi[w][traj][ncp] (pads zero): one thread per column pair
__global__ void k_fill_xi(double *__restrict__ xi, int nw, int ntraj, int nc, int ncp, uint64_t seed, long long traj0) {
    const int hp = ncp / 2;
    const size_t n = (size_t)nw * ntraj * hp;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const int kp = (int)(e % hp);
        const size_t r = e / hp;
        const int tr = (int)(r % ntraj), w = (int)(r / ntraj);
        double2 v = philox_normal2(seed, (uint32_t)w, (uint32_t)kp, (uint64_t)(traj0 + tr));
        if (2 * kp + 1 >= nc) v.y = 0.0;
        *reinterpret_cast<double2 *>(xi + 2 * e) = v;
    }
}
// injected draws: host layout [ntraj][nw][nc] -> [w][traj][ncp]
__global__ void k_scatter_xi(const double *__restrict__ src, double *__restrict__ xi, int nw, int ntraj, int nc, int ncp) {
    const size_t n = (size_t)nw * ntraj * ncp;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(e % ncp);
        const size_t r = e / ncp;
        const int tr = (int)(r % ntraj), w = (int)(r / ntraj);
        xi[e] = k < nc ? src[((size_t)tr * nw + w) * nc + k] : 0.0;
    }
}

// diagonal factors (diagonal spectra, e.g. the diagonal memory kernels of config 5): x[traj][w][c] = Ld[w][c] * xi, the draws made in
// place (same Philox counters as k_fill_xi, so the series equal those of the general path bit for bit) or read from injected xi[w][traj][ncp].
// One thread per column pair; consecutive threads write consecutive 16-byte pieces of a spectrum row.
__global__ void k_fill_x_diag(double *__restrict__ X, const double *__restrict__ Ld, const double *__restrict__ xi, int nw, int ntraj, int nc,
                              int ncp, uint64_t seed, long long traj0) {
    const int hp = ncp / 2;
    const size_t n = (size_t)nw * ntraj * hp;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const int kp = (int)(e % hp);
        const size_t r = e / hp;
        const int w = (int)(r % nw), tr = (int)(r / nw);           // X is [traj][w][ncp]
        double2 v;
        if (xi) v = *reinterpret_cast<const double2 *>(xi + ((size_t)w * ntraj + tr) * ncp + 2 * kp);
        else v = philox_normal2(seed, (uint32_t)w, (uint32_t)kp, (uint64_t)(traj0 + tr));
        const double2 l = *reinterpret_cast<const double2 *>(Ld + (size_t)w * ncp + 2 * kp);
        v.x *= l.x;
        v.y = 2 * kp + 1 < nc ? v.y * l.y : 0.0;
        *reinterpret_cast<double2 *>(X + 2 * e) = v;
    }
}

// ------------------------------------------------------------------ covariance + Jacobi factor
// G: column-major n x n (complex: interleaved), lives in shared or global memory.
template <bool CPLX>
__device__ void jacobi_onesided(double *G, int n, double normF2, int *flag) {
    constexpr int E = CPLX ? 2 : 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int m = n + (n & 1);  // players (a dummy if n is odd)
    const double floor2 = 1e-28 * normF2;
    const double tol = 8.9e-16 * sqrt((double)n);   // rounding level of an n-term dot product
    for (int sweep = 0; sweep < 40; ++sweep) {
        if (threadIdx.x == 0) *flag = 0;
        __syncthreads();
        for (int r = 0; r < m - 1; ++r) {
            for (int k = warp; k < m / 2; k += nwarps) {
                int p = (r + k) % (m - 1);
                int q = k == 0 ? m - 1 : (r - k + m - 1) % (m - 1);
                if (p >= n || q >= n) continue;
                if (p > q) { const int t = p; p = q; q = t; }
                double *gp = G + (size_t)p * n * E, *gq = G + (size_t)q * n * E;
                double a = 0, b = 0, cr = 0, ci = 0;
                for (int i = lane; i < n; i += 32) {
                    if (CPLX) {
                        const double pr = gp[2 * i], pi = gp[2 * i + 1], qr = gq[2 * i], qi = gq[2 * i + 1];
                        a += pr * pr + pi * pi;
                        b += qr * qr + qi * qi;
                        cr += pr * qr + pi * qi;   // conj(gp) * gq
                        ci += pr * qi - pi * qr;
                    } else {
                        const double pv = gp[i], qv = gq[i];
                        a += pv * pv; b += qv * qv; cr += pv * qv;
                    }
                }
                a = warp_sum(a); b = warp_sum(b); cr = warp_sum(cr);
                if (CPLX) ci = warp_sum(ci);
                const double gabs = CPLX ? sqrt(cr * cr + ci * ci) : fabs(cr);
                if (gabs <= tol * sqrt(a * b) || gabs <= floor2) continue;
                if (lane == 0) *flag = 1;
                const double zeta = (b - a) / (2.0 * gabs);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                // phase: gq~ = e^{-i phi} gq with gamma = |gamma| e^{i phi}
                const double er = CPLX ? cr / gabs : (cr >= 0 ? 1.0 : -1.0), ei = CPLX ? -ci / gabs : 0.0;
                for (int i = lane; i < n; i += 32) {
                    if (CPLX) {
                        const double pr = gp[2 * i], pi = gp[2 * i + 1], qr0 = gq[2 * i], qi0 = gq[2 * i + 1];
                        const double qr = qr0 * er - qi0 * ei, qi = qr0 * ei + qi0 * er;
                        gp[2 * i] = c * pr - s * qr; gp[2 * i + 1] = c * pi - s * qi;
                        gq[2 * i] = s * pr + c * qr; gq[2 * i + 1] = s * pi + c * qi;
                    } else {
                        const double pv = gp[i], qv = gq[i] * er;
                        gp[i] = c * pv - s * qv;
                        gq[i] = s * pv + c * qv;
                    }
                }
            }
            __syncthreads();
        }
        const int any = *flag;
        __syncthreads();
        if (!any) break;
    }
}

// One CTA per matrix.  A_w assembled from the basis, copied into G (shared if it fits, else the
// global scratch), rotated to orthogonal columns g_k = lambda_k v_k; then
//   lambda_k = g_k^H A g_k / |g_k|^2,   L[c][k] = g_k[c]/sqrt(lambda_k) if lambda_k > 0 else 0.
// Output L as real rows [nc][ncp] (+ an imaginary block [nc][ncp] after it when CPLX).
template <bool CPLX>
__global__ void __launch_bounds__(512) k_factor(const double *__restrict__ basis, int nc, int ncp, int nterm,
                                                 const int *__restrict__ idx, const double *__restrict__ cre,
                                                 const double *__restrict__ cim, double *__restrict__ Aglob,
                                                 double *__restrict__ Gglob, int use_smem, double *__restrict__ L,
                                                 double *__restrict__ evals, const int *__restrict__ wlist, int w0) {
    constexpr int E = CPLX ? 2 : 1;
    extern __shared__ double sm[];
    __shared__ double red[32];
    __shared__ int flag;
    const int w = wlist ? wlist[w0 + blockIdx.x] : w0 + blockIdx.x, n = nc;      // frequency index; scratch is indexed by the CTA
    const size_t nn = (size_t)n * n * E;
    double *A = Aglob + (size_t)blockIdx.x * nn;        // column-major copy of the hermitianised covariance
    double *G = use_smem ? sm : Gglob + (size_t)blockIdx.x * nn;
    double nf = 0.0;
    for (size_t e = threadIdx.x; e < (size_t)n * n; e += blockDim.x) {
        const int i = (int)(e % n), j = (int)(e / n);   // element (i,j), column-major
        double ar = 0, ai = 0;
        for (int m = 0; m < nterm; ++m) {
            const int bi = idx[w * nterm + m];
            if (bi < 0) continue;
            const double *B = basis + (size_t)bi * n * n;
            const double bij = B[(size_t)i * n + j], bji = B[(size_t)j * n + i];
            const double r = cre[w * nterm + m], im = CPLX ? cim[w * nterm + m] : 0.0;
            // 0.5*(a_ij + conj(a_ji)) with a = (r + i im) * B   (functions.py:198-200)
            ar += 0.5 * r * (bij + bji);
            ai += 0.5 * im * (bij - bji);
        }
        if (CPLX) {
            A[2 * e] = ar; A[2 * e + 1] = ai; G[2 * e] = ar; G[2 * e + 1] = ai;
            nf += ar * ar + ai * ai;
        } else {
            A[e] = ar; G[e] = ar;
            nf += ar * ar;
        }
    }
    nf = block_sum(nf, red);
    if (threadIdx.x == 0) red[0] = nf;
    __syncthreads();
    nf = red[0];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    double *Lw = L + (size_t)w * n * ncp * E;
    // attempt 0: rotate A itself (keeps tiny eigenvalues relatively accurate; exact for PSD input).
    // If A has eigenvalue pairs +-lambda the columns of A V need not be eigenvectors: detected by the
    // residual below, and attempt 1 repeats the rotation on the positive definite A + |A|_F I.
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (attempt == 1) {
            const double sigma = sqrt(nf);
            for (size_t e = threadIdx.x; e < (size_t)n * n; e += blockDim.x) {
                const int i = (int)(e % n), j = (int)(e / n);
                if (CPLX) { G[2 * e] = A[2 * e] + (i == j ? sigma : 0.0); G[2 * e + 1] = A[2 * e + 1]; }
                else G[e] = A[e] + (i == j ? sigma : 0.0);
            }
            __syncthreads();
        }
        if (nf > 0.0) jacobi_onesided<CPLX>(G, n, attempt ? 4.0 * nf * n : nf, &flag);
        __syncthreads();
        if (threadIdx.x == 0) flag = 0;
        __syncthreads();
        for (int k = warp; k < n; k += nwarps) {
            const double *g = G + (size_t)k * n * E;
            double num = 0, den = 0, w2 = 0;
            for (int i = lane; i < n; i += 32) {
                // (A g)_i = sum_j A_ij g_j ; A Hermitian, column-major: A_ij = conj(A_ji) -> read column i
                double sr = 0, si = 0;
                const double *acol = A + (size_t)i * n * E;
                for (int j = 0; j < n; ++j) {
                    if (CPLX) {
                        const double ar = acol[2 * j], ai = -acol[2 * j + 1];
                        sr += ar * g[2 * j] - ai * g[2 * j + 1];
                        si += ar * g[2 * j + 1] + ai * g[2 * j];
                    } else {
                        sr += acol[j] * g[j];
                    }
                }
                if (CPLX) {
                    num += g[2 * i] * sr + g[2 * i + 1] * si;
                    den += g[2 * i] * g[2 * i] + g[2 * i + 1] * g[2 * i + 1];
                } else {
                    num += g[i] * sr;
                    den += g[i] * g[i];
                }
                w2 += sr * sr + si * si;
            }
            num = warp_sum(num); den = warp_sum(den); w2 = warp_sum(w2);
            const double lam = den > 0 ? num / den : 0.0;
            // |A v - lam v|^2 = |A g|^2/|g|^2 - lam^2 ; an eigenvector has this at rounding level
            const double res2 = den > 0 ? w2 / den - lam * lam : 0.0;
            if (attempt == 0 && res2 > 1e-20 * nf && lane == 0) flag = 1;
            const double sc = (lam > 0 && den > 0) ? sqrt(lam / den) : 0.0;   // v_k sqrt(lam) = g_k/|g_k| sqrt(lam)
            if (lane == 0 && evals) evals[(size_t)w * n + k] = lam;
            for (int i = lane; i < n; i += 32) {
                if (CPLX) {
                    Lw[(size_t)i * ncp + k] = g[2 * i] * sc;
                    Lw[(size_t)(n + i) * ncp + k] = g[2 * i + 1] * sc;
                } else {
                    Lw[(size_t)i * ncp + k] = g[i] * sc;
                }
            }
        }
        __syncthreads();
        if (!flag) break;
        __syncthreads();
    }
}


// Pivoted Cholesky of a positive SEMI-definite covariance, one CTA per frequency: L L^H = A = clamp+(A) -- any such factor gives
// the same Gaussian law as the reference's V sqrt(lambda+) (noise.py:82-84, 299-303), at n^3/3 flops and n barriers instead of
// Jacobi sweeps.  A(w) = hermitianize(sum_m (cre + i cim) basis[idx]) is evaluated on the fly (never stored); thread i owns row i
// and the remaining diagonal d_i of the Schur complement.  Diagonal pivoting stops when max d_i <= 1e-13 max|a_ii|: for a PSD matrix
// every remaining entry is then below that bound.  A frequency is handed to the Jacobi kernel (status = 1) when the matrix is not
// PSD: a remaining diagonal below -1e-11 max|a_ii|, or |(A - L L^H) v|_2 > 1e-11 max|a_ii| for a +-1 pseudo-random vector v.
// L is kept column-major ([column k][row i]) in shared memory (n*n*E*8 <= 200 KB) or in a global scratch.
template <bool CPLX>
__global__ void __launch_bounds__(1024) k_chol(const double *__restrict__ basis, int nc, int ncp, int nterm, const int *__restrict__ idx,
                                                const double *__restrict__ cre, const double *__restrict__ cim, double *__restrict__ Lglob,
                                                int use_smem, double *__restrict__ L, int *__restrict__ status) {
    constexpr int E = CPLX ? 2 : 1;
    extern __shared__ double sm[];
    __shared__ double red[32], sval[4];
    __shared__ int redi[32], sidx;
    const int w = blockIdx.x, n = nc, i = threadIdx.x;
    const bool row = i < n;
    double *Lc = use_smem ? sm : Lglob + (size_t)blockIdx.x * n * n * E;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    auto aij = [&](int r, int c, double &ar, double &ai) {       // hermitianised covariance element (functions.py:198-200)
        ar = 0.0;
        ai = 0.0;
        for (int m = 0; m < nterm; ++m) {
            const int bi = idx[w * nterm + m];
            if (bi < 0) continue;
            const double *B = basis + (size_t)bi * n * n;
            const double brc = B[(size_t)r * n + c], bcr = B[(size_t)c * n + r];
            ar += 0.5 * cre[w * nterm + m] * (brc + bcr);
            if (CPLX) ai += 0.5 * cim[w * nterm + m] * (brc - bcr);
        }
    };
    auto block_max = [&](double v, int id, double &vmax, int &imax) {   // arg-max over the CTA (first index on ties)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, id, o);
            if (ov > v || (ov == v && oi < id)) { v = ov; id = oi; }
        }
        __syncthreads();
        if (lane == 0) { red[warp] = v; redi[warp] = id; }
        __syncthreads();
        if (warp == 0) {
            v = lane < nwarp ? red[lane] : -1e300;
            id = lane < nwarp ? redi[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int oi = __shfl_xor_sync(0xffffffffu, id, o);
                if (ov > v || (ov == v && oi < id)) { v = ov; id = oi; }
            }
            if (lane == 0) { sval[0] = v; sidx = id; }
        }
        __syncthreads();
        vmax = sval[0];
        imax = sidx;
    };
    double d = 0.0, dummy;
    if (row) aij(i, i, d, dummy);
    double amax;
    int piv;
    block_max(row ? fabs(d) : -1.0, i, amax, piv);
    double *Lw = L + (size_t)w * n * ncp * E;
    if (!(amax > 0.0)) {                                  // A == 0 (above the spectral cut-off): L = 0
        for (int e = threadIdx.x; e < n * ncp * E; e += blockDim.x) Lw[e] = 0.0;
        if (threadIdx.x == 0) status[w] = 0;
        return;
    }
    const double tol = 1e-13 * amax;
    bool done = !row;
    int rank = 0;
    for (int j = 0; j < n; ++j) {
        double dmax;
        block_max(done ? -1e300 : d, i, dmax, piv);
        if (!(dmax > tol)) break;
        rank = j + 1;
        const double inv = 1.0 / sqrt(dmax);
        if (row) {
            double lr = 0.0, li = 0.0;
            if (i == piv) {
                lr = sqrt(dmax);
            } else if (!done) {
                double ar, ai;
                aij(i, piv, ar, ai);
                for (int k = 0; k < j; ++k) {             // a_{i,piv} - sum_k L_ik conj(L_{piv,k})
                    const double *lk = Lc + (size_t)k * n * E;
                    if (CPLX) {
                        const double xr = lk[2 * i], xi = lk[2 * i + 1], yr = lk[2 * piv], yi = lk[2 * piv + 1];
                        ar -= xr * yr + xi * yi;
                        ai -= xi * yr - xr * yi;
                    } else {
                        ar -= lk[i] * lk[piv];
                    }
                }
                lr = ar * inv;
                li = ai * inv;
                d -= lr * lr + li * li;
            }
            double *lj = Lc + (size_t)j * n * E;
            if (CPLX) { lj[2 * i] = lr; lj[2 * i + 1] = li; } else lj[i] = lr;
            if (i == piv) { done = true; d = 0.0; }
        }
        __syncthreads();
    }
    // not PSD?  (1) a clearly negative remaining diagonal
    double dmin;
    block_max(done ? -1e300 : -d, i, dmin, piv);
    int bad = (-dmin < -1e-11 * amax) ? 1 : 0;            // dmin = max(-d_i)  ->  min d_i = -dmin
    // (2) residual of the factorisation on a pseudo-random +-1 vector: y = L^H v, r = A v - L y
    auto sgn = [&](int k) { return ((unsigned)(k * 2654435761u + (unsigned)w * 40503u) >> 15) & 1u ? 1.0 : -1.0; };
    // y lives in the first unused column of Lc (columns >= rank are free); with full rank, in the output buffer before it is written
    double *ybuf = rank < n ? Lc + (size_t)rank * n * E : Lw;
    __syncthreads();
    if (row && i < rank) {                                 // thread k = i: y_k = sum_r conj(L_rk) v_r
        const double *lk = Lc + (size_t)i * n * E;
        double yr = 0.0, yi = 0.0;
        for (int r = 0; r < n; ++r) {
            const double v = sgn(r);
            if (CPLX) { yr += lk[2 * r] * v; yi -= lk[2 * r + 1] * v; } else yr += lk[r] * v;
        }
        if (CPLX) { ybuf[2 * i] = yr; ybuf[2 * i + 1] = yi; } else ybuf[i] = yr;
    }
    __syncthreads();
    double r2 = 0.0;
    if (row) {
        double rr = 0.0, ri = 0.0;
        for (int c = 0; c < n; ++c) {
            double ar, ai;
            aij(i, c, ar, ai);
            const double v = sgn(c);
            rr += ar * v;
            ri += ai * v;
        }
        for (int k = 0; k < rank; ++k) {
            const double *lk = Lc + (size_t)k * n * E;
            if (CPLX) {
                const double xr = lk[2 * i], xi = lk[2 * i + 1], yr = ybuf[2 * k], yi = ybuf[2 * k + 1];
                rr -= xr * yr - xi * yi;
                ri -= xr * yi + xi * yr;
            } else {
                rr -= lk[i] * ybuf[k];
            }
        }
        r2 = rr * rr + ri * ri;
    }
    r2 = block_sum(r2, red);
    if (threadIdx.x == 0) sval[1] = r2;
    __syncthreads();
    if (sqrt(sval[1]) > 1e-11 * amax) bad = 1;
    if (threadIdx.x == 0) status[w] = bad;
    if (bad) return;
    __syncthreads();
    for (int e = threadIdx.x; e < n * ncp; e += blockDim.x) {      // L rows [n][ncp] (+ the imaginary block), zero beyond the rank
        const int r = e / ncp, k = e % ncp;
        double lr = 0.0, li = 0.0;
        if (k < rank) {
            const double *lk = Lc + (size_t)k * n * E;
            if (CPLX) { lr = lk[2 * r]; li = lk[2 * r + 1]; } else lr = lk[r];
        }
        Lw[(size_t)r * ncp + k] = lr;
        if (CPLX) Lw[(size_t)(n + r) * ncp + k] = li;
    }
}

// single-basis shortcut: A_w = c_w * B  =>  L_w = V diag(sqrt(max(c_w*lambda, 0)))
// L0 holds g_k/sqrt(|lambda_k|)-style data?  No: we pass V (unit eigenvectors) and lambda explicitly.
__global__ void k_scale_factor(const double *__restrict__ V, const double *__restrict__ lam, const double *__restrict__ cw,
                               int nc, int ncp, double *__restrict__ L) {
    const int w = blockIdx.x;
    const double c = cw[w];
    for (int e = threadIdx.x; e < nc * ncp; e += blockDim.x) {
        const int k = e % ncp;
        double v = 0.0;
        if (k < nc) {
            const double s = c * lam[k];
            v = s > 0 ? V[e] * sqrt(s) : 0.0;
        }
        L[(size_t)w * nc * ncp + e] = v;
    }
}
// V[c][k] = L1[c][k] / sqrt(lambda_k) where L1 is the factor of B itself (lambda_k > 0), for lambda_k < 0 we
// need the eigenvector too (c_w may be negative): recover from the factor of -B.
__global__ void k_unit_vectors(const double *__restrict__ Lpos, const double *__restrict__ Lneg, const double *__restrict__ lpos,
                               const double *__restrict__ lneg, int nc, int ncp, double *__restrict__ V, double *__restrict__ lam) {
    // columns of Lpos with lpos>0 are eigenvectors of B with eigenvalue lpos; columns of Lneg (factor of -B)
    // with lneg>0 are eigenvectors of B with eigenvalue -lneg.  Pack both sets (at most nc non-null in total).
    __shared__ int map[1024];
    __shared__ int cnt;
    if (threadIdx.x == 0) {
        int c = 0;
        for (int k = 0; k < nc && c < nc; ++k) if (lpos[k] > 0) map[c++] = k;
        for (int k = 0; k < nc && c < nc; ++k) if (lneg[k] > 0) map[c++] = -(k + 1);
        cnt = c;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nc * ncp; e += blockDim.x) {
        const int i = e / ncp, k = e % ncp;
        double v = 0.0;
        if (k < cnt) {
            const int src = map[k];
            v = src >= 0 ? Lpos[(size_t)i * ncp + src] / sqrt(lpos[src]) : Lneg[(size_t)i * ncp + (-src - 1)] / sqrt(lneg[-src - 1]);
        }
        V[e] = v;
    }
    for (int k = threadIdx.x; k < nc; k += blockDim.x) {
        double l = 0.0;
        if (k < cnt) l = map[k] >= 0 ? lpos[map[k]] : -lneg[-map[k] - 1];
        lam[k] = l;
    }
}

// ------------------------------------------------------------------ FFT (Stockham autosort, radix 2/3/4/5)
struct FftPlan {
    int n, npass, radix[24];
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

template <int R>
__device__ __forceinline__ void dft_small(double2 (&v)[R]) {
    if (R == 2) {
        const double2 a = v[0], b = v[1];
        v[0] = make_double2(a.x + b.x, a.y + b.y);
        v[1] = make_double2(a.x - b.x, a.y - b.y);
    } else if (R == 4) {  // forward (exp(-i...)): multiply by -i = (y, -x)
        const double2 a = make_double2(v[0].x + v[2].x, v[0].y + v[2].y), b = make_double2(v[0].x - v[2].x, v[0].y - v[2].y);
        const double2 c = make_double2(v[1].x + v[3].x, v[1].y + v[3].y), d = make_double2(v[1].x - v[3].x, v[1].y - v[3].y);
        v[0] = make_double2(a.x + c.x, a.y + c.y);
        v[2] = make_double2(a.x - c.x, a.y - c.y);
        v[1] = make_double2(b.x + d.y, b.y - d.x);
        v[3] = make_double2(b.x - d.y, b.y + d.x);
    } else {
        double2 o[R];
#pragma unroll
        for (int m = 0; m < R; ++m) {
            double2 s = v[0];
#pragma unroll
            for (int j = 1; j < R; ++j) {
                double sn, cs;
                sincospi(-2.0 * ((m * j) % R) / R, &sn, &cs);
                s.x += v[j].x * cs - v[j].y * sn;
                s.y += v[j].x * sn + v[j].y * cs;
            }
            o[m] = s;
        }
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = o[m];
    }
}

template <int R>
__device__ __forceinline__ void stockham_pass(const double2 *__restrict__ in, double2 *__restrict__ out, int n, int ns, int batch) {
    const int per = n / R;
    for (int e = threadIdx.x; e < per * batch; e += blockDim.x) {
        const int f = e / per, j = e % per;
        const int k = j % ns;
        const double2 *src = in + (size_t)f * n;
        double2 *dst = out + (size_t)f * n;
        double2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) {
            v[m] = src[j + m * per];
            if (m > 0 && k > 0) {
                double sn, cs;
                sincospi(-2.0 * (double)(m * k) / (double)(ns * R), &sn, &cs);
                v[m] = cmul(v[m], make_double2(cs, sn));
            }
        }
        dft_small<R>(v);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int m = 0; m < R; ++m) dst[j0 + m * ns] = v[m];
    }
}

// in-place over two ping-pong buffers; returns the buffer holding the result (natural order)
__device__ double2 *fft_smem(double2 *a, double2 *b, const FftPlan &pl, int batch) {
    int ns = 1;
    for (int p = 0; p < pl.npass; ++p) {
        const int R = pl.radix[p];
        if (R == 4) stockham_pass<4>(a, b, pl.n, ns, batch);
        else if (R == 2) stockham_pass<2>(a, b, pl.n, ns, batch);
        else if (R == 5) stockham_pass<5>(a, b, pl.n, ns, batch);
        else stockham_pass<3>(a, b, pl.n, ns, batch);
        ns *= R;
        __syncthreads();
        double2 *t = a; a = b; b = t;
    }
    return a;
}

// value of the mirrored, column-packed spectrum Z[k] = Xa[k] + i Xb[k] (noise.py:87-94):
//   X[k] = x_k (k < h),  conj(x_{N-k}) (k >= h)     with x stored as [traj][w][ncx] (re block, then im block)
__device__ __forceinline__ double2 load_Z(const double *__restrict__ X, int k, int N, int h, size_t wstride, size_t off, int imoff,
                                          bool has_b) {
    const bool mir = k >= h;
    const int w = mir ? N - k : k;
    const double *r = X + (size_t)w * wstride + off;
    double ar = r[0], br = has_b ? r[1] : 0.0, ai = 0.0, bi = 0.0;
    if (imoff) {
        ai = r[imoff];
        bi = has_b ? r[imoff + 1] : 0.0;
    }
    if (mir) { ai = -ai; bi = -bi; }
    // bins 0 and N/2 are unpaired: their imaginary parts only feed Im(series), which the
    // reference discards (np.real, baths.py:191,408); dropping them makes both packed transforms real
    if (w == 0 || k == h) { ai = 0.0; bi = 0.0; }
    return make_double2(ar - bi, ai + br);   // (ar + i ai) + i (br + i bi)
}


// ---- in-place transform in shared memory (decimation in time) ----
// One buffer of N complex numbers instead of the Stockham ping-pong pair: N = 8192 (128 KB) fits one CTA.  The input is loaded in
// digit-reversed order (a gather from global memory anyway), the passes then work in place and leave the result in natural order.
// Pass p (radix R_p, ns = R_0 ... R_{p-1}): for every group g and k < ns the elements g ns R_p + k + m ns (m < R_p) are multiplied by
// w^{m k}, w = exp(-2 pi i / (ns R_p)), and replaced by their R_p-point DFT.  Twiddles come from a table tw[j] = exp(-2 pi i j / N)
// (built once per plan, L1/L2-resident): one 16-byte load and R-2 complex products per butterfly instead of sincospi calls.
__global__ void k_twiddle(double2 *__restrict__ tw, int N) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    double sn, cs;
    sincospi(-2.0 * (double)j / (double)N, &sn, &cs);
    tw[j] = make_double2(cs, sn);
}
__device__ __forceinline__ int digit_reverse(int pos, const FftPlan &pl) {     // source index of position `pos`
    int n = 0;
    for (int q = 0; q < pl.npass; ++q) {      // digit q of pos (radix R_q, least significant first) becomes the next more significant digit of n
        n = n * pl.radix[q] + pos % pl.radix[q];
        pos /= pl.radix[q];
    }
    return n;
}
__global__ void k_digit_reverse(int *__restrict__ perm, FftPlan pl) {           // once per plan: the load order of the in-place transform
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos < pl.n) perm[pos] = digit_reverse(pos, pl);
}
// lg >= 0: ns = 2^lg (shifts instead of integer divisions; every pass of a power-of-two length)
template <int R>
__device__ __forceinline__ void dit_pass(double2 *__restrict__ a, int N, int ns, int lg, const double2 *__restrict__ tw) {
    const int per = N / R, np = ns * R, tstep = N / np;
    for (int j = threadIdx.x; j < per; j += blockDim.x) {
        const int k = lg >= 0 ? (j & (ns - 1)) : j % ns, g = lg >= 0 ? (j >> lg) : j / ns;
        double2 *base = a + g * np + k;
        double2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = base[m * ns];
        if (k > 0) {
            const double2 w1 = __ldg(tw + k * tstep);
            double2 wm = w1;
#pragma unroll
            for (int m = 1; m < R; ++m) {
                v[m] = cmul(v[m], wm);
                if (m + 1 < R) wm = cmul(wm, w1);
            }
        }
        dft_small<R>(v);
#pragma unroll
        for (int m = 0; m < R; ++m) base[m * ns] = v[m];
    }
}
__device__ void fft_inplace_smem(double2 *a, const FftPlan &pl, const double2 *__restrict__ tw) {
    int ns = 1;
    for (int p = 0; p < pl.npass; ++p) {
        const int R = pl.radix[p];
        const int lg = (ns & (ns - 1)) == 0 ? 31 - __clz(ns) : -1;
        if (R == 4) dit_pass<4>(a, pl.n, ns, lg, tw);
        else if (R == 2) dit_pass<2>(a, pl.n, ns, lg, tw);
        else if (R == 5) dit_pass<5>(a, pl.n, ns, lg, tw);
        else dit_pass<3>(a, pl.n, ns, lg, tw);
        ns *= R;
        __syncthreads();
    }
}
// ---- power-of-two lengths: big-radix passes with the butterflies in registers ----
// 8192 = 16 x 16 x 32: THREE passes instead of seven radix-4 ones, and the first / last of them read global memory / write the
// noise table directly, so a series crosses shared memory twice (write, read+write, read) instead of fourteen times -- the radix-4
// kernel above is bound by exactly that traffic.  Shared-memory index i is stored at i + (i >> 4): the stride-16 / stride-256
// accesses of the passes then fall on distinct banks.  A radix-R butterfly (R <= 32) is an unrolled radix-2 DIT network on registers
// with compile-time twiddles exp(-2 pi i k / 32).
__device__ constexpr double FFT_C32[32] = {1, 0.98078528040323043, 0.92387953251128674, 0.83146961230254524, 0.70710678118654757, 0.55557023301960229, 0.38268343236508984, 0.19509032201612833, 6.123233995736766e-17, -0.19509032201612819, -0.38268343236508973, -0.55557023301960196, -0.70710678118654746, -0.83146961230254535, -0.92387953251128674, -0.98078528040323043, -1, -0.98078528040323043, -0.92387953251128685, -0.83146961230254546, -0.70710678118654768, -0.55557023301960218, -0.38268343236509034, -0.19509032201612866, -1.8369701987210297e-16, 0.1950903220161283, 0.38268343236509, 0.55557023301960184, 0.70710678118654735, 0.83146961230254524, 0.92387953251128652, 0.98078528040323032};
__device__ constexpr double FFT_S32[32] = {-0, -0.19509032201612825, -0.38268343236508978, -0.55557023301960218, -0.70710678118654746, -0.83146961230254524, -0.92387953251128674, -0.98078528040323043, -1, -0.98078528040323043, -0.92387953251128674, -0.83146961230254546, -0.70710678118654757, -0.55557023301960218, -0.38268343236508989, -0.19509032201612861, -1.2246467991473532e-16, 0.19509032201612836, 0.38268343236508967, 0.55557023301960196, 0.70710678118654746, 0.83146961230254524, 0.92387953251128652, 0.98078528040323032, 1, 0.98078528040323043, 0.92387953251128663, 0.83146961230254546, 0.70710678118654768, 0.55557023301960218, 0.38268343236509039, 0.19509032201612872};     // -sin: forward transform
__device__ constexpr int FFT_BR32[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30, 1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};
__host__ __device__ constexpr int fft_log2(int r) { return r <= 1 ? 0 : 1 + fft_log2(r >> 1); }
template <int R, int LEN>
struct FftStage {       // one radix-2 stage (butterflies of span LEN) of the register network, then the next one
    static __device__ __forceinline__ void run(double2 (&t)[R]) {
#pragma unroll
        for (int i = 0; i < R; i += LEN) {
#pragma unroll
            for (int k = 0; k < LEN / 2; ++k) {
                constexpr int step = 32 / LEN;           // exp(-2 pi i k / LEN) = table entry k * 32 / LEN
                const int e = k * step;
                const double2 u = t[i + k], y = t[i + k + LEN / 2];
                double2 x;
                if (e == 0) x = y;
                else if (e == 8) x = make_double2(y.y, -y.x);      // times -i
                else x = make_double2(y.x * FFT_C32[e] - y.y * FFT_S32[e], y.x * FFT_S32[e] + y.y * FFT_C32[e]);
                t[i + k] = make_double2(u.x + x.x, u.y + x.y);
                t[i + k + LEN / 2] = make_double2(u.x - x.x, u.y - x.y);
            }
        }
        FftStage<R, LEN * 2>::run(t);
    }
};
template <int R>
struct FftStage<R, 2 * R> {
    static __device__ __forceinline__ void run(double2 (&)[R]) {}
};
template <int R>
__device__ __forceinline__ void fft_reg(double2 (&v)[R]) {
    constexpr int L = fft_log2(R);
    double2 t[R];
#pragma unroll
    for (int i = 0; i < R; ++i) t[FFT_BR32[i] >> (5 - L)] = v[i];        // bit reversal in L bits
    FftStage<R, 2>::run(t);
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = t[i];
}
__host__ __device__ __forceinline__ int fft_pad(int i) { return i + (i >> 4); }

struct BigFftArgs {
    const double *X;
    const double2 *tw;
    const int *perm;
    double2 *fs;
    int N, h, nc, imoff, has_b;
    size_t wstride, off;
    double scale;
    double *out;
    size_t out_nstride;
    const double2 *stage;    // real spectra: the h + 1 rows (x_a[w], x_b[w]) of this column pair staged in shared memory, or NULL
};
template <int R, bool FIRST, bool LAST>
__device__ __forceinline__ void big_pass(const BigFftArgs &a, int ns, int lg) {
    const int N = a.N, per = N / R, np = ns * R, tstep = N / np;
    for (int j = threadIdx.x; j < per; j += blockDim.x) {
        const int k = j & (ns - 1), g = j >> lg;
        const int i0 = g * np + k;
        double2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int i = i0 + m * ns;
            if (FIRST) {
                const int k2 = __ldg(a.perm + i);
                // real spectrum: Z[k] = x_a[w] + i x_b[w], w = k below the Nyquist bin, N - k above it (mirror without conjugation work)
                v[m] = a.stage ? a.stage[k2 >= a.h ? N - k2 : k2] : load_Z(a.X, k2, N, a.h, a.wstride, a.off, a.imoff, a.has_b != 0);
            } else {
                v[m] = a.fs[fft_pad(i)];
            }
        }
        if (!FIRST && k > 0) {          // the first pass has ns = 1: no twiddles
            const double2 w1 = __ldg(a.tw + k * tstep);
            double2 wm = w1;
#pragma unroll
            for (int m = 1; m < R; ++m) {
                if ((m & 7) == 0) wm = __ldg(a.tw + ((m * k * tstep) & (N - 1)));      // fresh from the table: bounds the rounding of the running product
                v[m] = cmul(v[m], wm);
                if (m + 1 < R && ((m + 1) & 7) != 0) wm = cmul(wm, w1);
            }
        }
        fft_reg<R>(v);
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int i = i0 + m * ns;
            if (LAST) {
                double *o = a.out + (size_t)i * a.out_nstride;
                if (a.has_b && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                    *reinterpret_cast<double2 *>(o) = make_double2(v[m].x * a.scale, v[m].y * a.scale);
                } else {
                    o[0] = v[m].x * a.scale;
                    if (a.has_b) o[1] = v[m].y * a.scale;
                }
            } else {
                a.fs[fft_pad(i)] = v[m];
            }
        }
    }
}
template <bool FIRST, bool LAST>
__device__ __forceinline__ void big_pass_r(int R, const BigFftArgs &a, int ns, int lg) {
    if (R == 16) big_pass<16, FIRST, LAST>(a, ns, lg);
    else if (R == 32) big_pass<32, FIRST, LAST>(a, ns, lg);
    else if (R == 8) big_pass<8, FIRST, LAST>(a, ns, lg);
    else if (R == 4) big_pass<4, FIRST, LAST>(a, ns, lg);
    else big_pass<2, FIRST, LAST>(a, ns, lg);
}
__global__ void __launch_bounds__(256, 1) k_fft_big(const double *__restrict__ X, FftPlan pl, const double2 *__restrict__ tw, const int *__restrict__ perm,
                                                     int nc, int ncp, int ncx, int imoff, double scale, double *__restrict__ out, size_t out_tstride,
                                                     size_t out_nstride, int stage_ok) {
    extern __shared__ double2 fs[];
    const int pair = blockIdx.x, tr = blockIdx.y, c0 = 2 * pair;
    const bool stage_rows = stage_ok && c0 + 1 < nc;       // (the odd last column of an odd nc keeps the register path)
    BigFftArgs a;
    a.X = X; a.tw = tw; a.perm = perm; a.fs = fs; a.N = pl.n; a.h = pl.n / 2; a.nc = nc; a.imoff = imoff; a.has_b = c0 + 1 < nc;
    a.wstride = (size_t)ncx; a.off = (size_t)tr * (a.h + 1) * ncx + c0;
    a.scale = scale; a.out = out + (size_t)tr * out_tstride + c0; a.out_nstride = out_nstride;
    a.stage = nullptr;
    if (stage_rows) {
        // Real spectra with both columns present: the h + 1 spectrum rows of this pair (16 bytes each at a row stride of ncx doubles) go to
        // shared memory as cp.async copies that are all in flight at once; the digit-reversed, mirrored gather of the first pass then
        // reads shared memory.  (Register-staged, the gather was 2 x 8-byte loads per point, every row fetched twice, and their latency
        // sat in front of the first butterflies: 1 TB/s.)
        double2 *stg = fs + fft_pad(a.N) + 1;
        const double *src = X + a.off;
        for (int w = threadIdx.x; w <= a.h; w += blockDim.x) cp_async16_zfill(stg + w, src + (size_t)w * a.wstride, 16);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        a.stage = stg;
    }
    int ns = 1;
    for (int p = 0; p < pl.npass; ++p) {
        const int R = pl.radix[p], lg = 31 - __clz(ns);
        const bool first = p == 0, last = p == pl.npass - 1;
        if (first && last) big_pass_r<true, true>(R, a, ns, lg);
        else if (first) big_pass_r<true, false>(R, a, ns, lg);
        else if (last) big_pass_r<false, true>(R, a, ns, lg);
        else big_pass_r<false, false>(R, a, ns, lg);
        ns *= R;
        if (!last) __syncthreads();
    }
}
// radix sequence of the big-radix kernel for N = 2^e >= 2: 16, 16, ..., rest (<= 32)
bool make_big_plan(int n, FftPlan &pl) {
    if (n < 2 || (n & (n - 1))) return false;
    pl.n = n;
    pl.npass = 0;
    int rem = n;
    while (rem > 32) { pl.radix[pl.npass++] = 16; rem /= 16; }
    pl.radix[pl.npass++] = rem;
    return true;
}

// one (trajectory, column pair) per CTA; CTAs of neighbouring pairs run side by side and touch the same rows of X and of the
// output table at the same time (the 16-byte accesses of a pair combine to full sectors / DRAM pages in L2)
__global__ void __launch_bounds__(512) k_fft_inplace(const double *__restrict__ X, FftPlan pl, const double2 *__restrict__ tw,
                                                      const int *__restrict__ perm, int ntraj_chunk,
                                                      int nc, int ncp, int ncx, int imoff, double scale, double *__restrict__ out,
                                                      size_t out_tstride, size_t out_nstride) {
    extern __shared__ double2 fs[];
    const int N = pl.n, h = N / 2;
    const int pair = blockIdx.x, tr = blockIdx.y;
    const int c0 = 2 * pair;
    const bool has_b = c0 + 1 < nc;
    const size_t wstride = (size_t)ncx, off = (size_t)tr * (h + 1) * ncx + c0;       // X is [traj][w][ncx]
    for (int pos = threadIdx.x; pos < N; pos += blockDim.x) fs[pos] = load_Z(X, __ldg(perm + pos), N, h, wstride, off, imoff, has_b);
    __syncthreads();
    fft_inplace_smem(fs, pl, tw);
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double *o = out + (size_t)n * out_nstride + (size_t)tr * out_tstride + c0;
        if (has_b && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
            *reinterpret_cast<double2 *>(o) = make_double2(fs[n].x * scale, fs[n].y * scale);
        } else {
            o[0] = fs[n].x * scale;
            if (has_b) o[1] = fs[n].y * scale;
        }
    }
}

// four-step, N = N1*N2:  step 1 = N2 transforms of length N1 over stride-N2 data (+ twiddle), tile of k2 per CTA
__global__ void __launch_bounds__(256) k_fft_step1(const double *__restrict__ X, FftPlan p1, int N, int N2, int tile, int ntraj_chunk,
                                                    int nc, int ncx, int imoff, double2 *__restrict__ scratch) {
    extern __shared__ double2 fs[];
    const int N1 = p1.n, h = N / 2;
    const int pair = blockIdx.x, tr = blockIdx.y, k20 = blockIdx.z * tile;
    const int nt = min(tile, N2 - k20);
    const int c0 = 2 * pair;
    const bool has_b = c0 + 1 < nc;
    double2 *a = fs, *b = fs + (size_t)tile * N1;
    const size_t wstride = (size_t)ncx, off = (size_t)tr * (h + 1) * ncx + c0;       // X is [traj][w][ncx]
    for (int e = threadIdx.x; e < nt * N1; e += blockDim.x) {
        const int f = e % nt, k1 = e / nt;     // consecutive threads -> consecutive k2 (adjacent frequencies)
        a[(size_t)f * N1 + k1] = load_Z(X, k1 * N2 + k20 + f, N, h, wstride, off, imoff, has_b);
    }
    __syncthreads();
    double2 *y = fft_smem(a, b, p1, nt);
    double2 *S = scratch + ((size_t)tr * gridDim.x + pair) * N;   // [n1][k2]
    for (int e = threadIdx.x; e < nt * N1; e += blockDim.x) {
        const int f = e % nt, n1 = e / nt;
        double sn, cs;
        sincospi(-2.0 * (double)((long long)n1 * (k20 + f) % N) / (double)N, &sn, &cs);
        S[(size_t)n1 * N2 + k20 + f] = cmul(y[(size_t)f * N1 + n1], make_double2(cs, sn));
    }
}
// step 2 = N1 transforms of length N2 over contiguous rows; Y[n1 + N1*n2]
__global__ void __launch_bounds__(256) k_fft_step2(const double2 *__restrict__ scratch, FftPlan p2, int N, int N1, int tile, int nc,
                                                    double scale, double *__restrict__ out, size_t out_tstride, size_t out_nstride) {
    extern __shared__ double2 fs[];
    const int N2 = p2.n;
    const int pair = blockIdx.x, tr = blockIdx.y, n10 = blockIdx.z * tile;
    const int nt = min(tile, N1 - n10);
    const int c0 = 2 * pair;
    const bool has_b = c0 + 1 < nc;
    double2 *a = fs, *b = fs + (size_t)tile * N2;
    const double2 *S = scratch + ((size_t)tr * gridDim.x + pair) * N + (size_t)n10 * N2;
    for (int e = threadIdx.x; e < nt * N2; e += blockDim.x) a[e] = S[e];
    __syncthreads();
    double2 *y = fft_smem(a, b, p2, nt);
    for (int e = threadIdx.x; e < nt * N2; e += blockDim.x) {
        const int f = e % nt, n2 = e / nt;
        const int n = n10 + f + N1 * n2;
        double *o = out + (size_t)n * out_nstride + (size_t)tr * out_tstride + c0;
        const double2 v = y[(size_t)f * N2 + n2];
        o[0] = v.x * scale;
        if (has_b) o[1] = v.y * scale;
    }
}

// ---- generic batched complex transform (functions.myfft, functions.py:11-53; power spectra, functions.py:203-236) ----
// Y[f][m] = scale * sum_k X[f][k] exp(sgn * 2 pi i k m / N); planar input/output [batch][N] (im may be NULL).
// sgn = +1 is computed as the conjugate of the forward transform of the conjugate.
__device__ __forceinline__ double2 load_c(const double *__restrict__ re, const double *__restrict__ im, size_t i, double cj) {
    return make_double2(re[i], im ? cj * im[i] : 0.0);
}
__global__ void __launch_bounds__(256) k_c2c_direct(const double *__restrict__ re, const double *__restrict__ im, FftPlan pl, double cj,
                                                     double scale, double *__restrict__ ore, double *__restrict__ oim) {
    extern __shared__ double2 fs[];
    const int N = pl.n;
    const size_t base = (size_t)blockIdx.x * N;
    double2 *a = fs, *b = fs + N;
    for (int k = threadIdx.x; k < N; k += blockDim.x) a[k] = load_c(re, im, base + k, cj);
    __syncthreads();
    double2 *y = fft_smem(a, b, pl, 1);
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        ore[base + n] = y[n].x * scale;
        oim[base + n] = cj * y[n].y * scale;
    }
}
__global__ void __launch_bounds__(256) k_c2c_step1(const double *__restrict__ re, const double *__restrict__ im, FftPlan p1, int N, int N2,
                                                    int tile, double cj, double2 *__restrict__ scratch) {
    extern __shared__ double2 fs[];
    const int N1 = p1.n, k20 = blockIdx.y * tile, nt = min(tile, N2 - k20);
    const size_t base = (size_t)blockIdx.x * N;
    double2 *a = fs, *b = fs + (size_t)tile * N1;
    for (int e = threadIdx.x; e < nt * N1; e += blockDim.x) {
        const int f = e % nt, k1 = e / nt;
        a[(size_t)f * N1 + k1] = load_c(re, im, base + (size_t)k1 * N2 + k20 + f, cj);
    }
    __syncthreads();
    double2 *y = fft_smem(a, b, p1, nt);
    double2 *S = scratch + base;   // [n1][k2]
    for (int e = threadIdx.x; e < nt * N1; e += blockDim.x) {
        const int f = e % nt, n1 = e / nt;
        double sn, cs;
        sincospi(-2.0 * (double)((long long)n1 * (k20 + f) % N) / (double)N, &sn, &cs);
        S[(size_t)n1 * N2 + k20 + f] = cmul(y[(size_t)f * N1 + n1], make_double2(cs, sn));
    }
}
__global__ void __launch_bounds__(256) k_c2c_step2(const double2 *__restrict__ scratch, FftPlan p2, int N, int N1, int tile, double cj,
                                                    double scale, double *__restrict__ ore, double *__restrict__ oim) {
    extern __shared__ double2 fs[];
    const int N2 = p2.n, n10 = blockIdx.y * tile, nt = min(tile, N1 - n10);
    const size_t base = (size_t)blockIdx.x * N;
    double2 *a = fs, *b = fs + (size_t)tile * N2;
    const double2 *S = scratch + base + (size_t)n10 * N2;
    for (int e = threadIdx.x; e < nt * N2; e += blockDim.x) a[e] = S[e];
    __syncthreads();
    double2 *y = fft_smem(a, b, p2, nt);
    for (int e = threadIdx.x; e < nt * N2; e += blockDim.x) {
        const int f = e % nt, n2 = e / nt;
        const size_t n = base + n10 + f + (size_t)N1 * n2;
        const double2 v = y[(size_t)f * N2 + n2];
        ore[n] = v.x * scale;
        oim[n] = cj * v.y * scale;
    }
}

bool make_fft_plan(int n, FftPlan &pl) {
    pl.n = n;
    pl.npass = 0;
    int m = n;
    while (m % 4 == 0) { pl.radix[pl.npass++] = 4; m /= 4; }
    while (m % 2 == 0) { pl.radix[pl.npass++] = 2; m /= 2; }
    while (m % 5 == 0) { pl.radix[pl.npass++] = 5; m /= 5; }
    while (m % 3 == 0) { pl.radix[pl.npass++] = 3; m /= 3; }
    return m == 1 && pl.npass <= 24;
}

// gamt (baths.py:35-50): C[t][i] = cos(w_i t)                                                    (eta == 0, baths.py:40)
//                       or  Re[ w/(w - i eta) e^{-i w t - eta t} + w/(w + i eta) e^{+i w t - eta t} ]/2   (eta != 0, baths.py:48-49)
//                         = e^{-eta t} (w^2 cos wt + w eta sin wt)/(w^2 + eta^2)
__global__ void k_cos_table(const double *__restrict__ tl, const double *__restrict__ wl, int nt, int nw, int nwp, double eta,
                            double *__restrict__ C) {
    const size_t n = (size_t)nt * nwp;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % nwp), t = (int)(e / nwp);
        double v = 0.0;
        if (i < nw) {
            const double w = wl[i], tt = tl[t];
            if (eta == 0.0) v = cos(w * tt);
            else v = exp(-eta * tt) * (w * w * cos(w * tt) + w * eta * sin(w * tt)) / (w * w + eta * eta);
        }
        C[e] = v;
    }
}

}  // namespace

// ------------------------------------------------------------------ plan
struct sclmd_noise_plan {
    int device = 0, nmd = 0, nw = 0, nc = 0, ncp = 0, cplx = 0, nsm = 148;
    double dt = 0;
    cudaStream_t st = nullptr;
    DevBuf<double> L;      // [nw][nc*(1+cplx)][ncp]
    DevBuf<double> evals;  // [nw][nc]
    DevBuf<double> Ld;     // [nw][ncp]: the factors' diagonals when every L_w is diagonal (diagonal spectra: x = Ld * xi, no product)
    bool diagL = false;
    DevBuf<double2> tw;    // twiddle table exp(-2 pi i j / nmd) of the in-place transform
    DevBuf<int> perm;      // its digit-reversed load order
    TmaWorkspace tws;      // stream-K scratch of the batched x = L xi product
    int64_t launches = 0;
    int nchol = 0, njacobi = 0;     // frequencies factorised by the pivoted Cholesky kernel / handed to Jacobi
    // stage timing of the last generate call (CUDA events on the generating stream): 0 draws, 1 x = L xi, 2 transform
    double stage_ms[3] = {0, 0, 0};
    double factor_ms = 0;
};

namespace {

int factor_batch(sclmd_noise_plan *pl, int nw, int nbasis, const double *basis_h, int nterm, const int *idx_h, const double *cre_h,
                 const double *cim_h, bool cplx, double *Ldst, double *evdst, bool eigen_only) {
    const int nc = pl->nc, ncp = pl->ncp, E = cplx ? 2 : 1;
    DevBuf<double> basis, cre, cim, A, G;
    DevBuf<int> idx, status, wlist_d;
    SCLMD_CUDA(basis.alloc((size_t)nbasis * nc * nc));
    SCLMD_CUDA(cre.alloc((size_t)nw * nterm));
    SCLMD_CUDA(cim.alloc((size_t)nw * nterm));
    SCLMD_CUDA(idx.alloc((size_t)nw * nterm));
    SCLMD_CUDA(cudaMemcpy(basis.p, basis_h, basis.n * sizeof(double), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(cre.p, cre_h, cre.n * sizeof(double), cudaMemcpyHostToDevice));
    if (cim_h) SCLMD_CUDA(cudaMemcpy(cim.p, cim_h, cim.n * sizeof(double), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(idx.p, idx_h, idx.n * sizeof(int), cudaMemcpyHostToDevice));
    const size_t nn = (size_t)nc * nc * E;
    const size_t smem = nn * sizeof(double);
    const int use_smem = smem <= 200 * 1024;
    std::vector<int> wlist;
    if (eigen_only || getenv("SCLMD_NOISE_JACOBI")) {       // the single-basis shortcut needs eigenvectors; the switch is for A/B runs
        for (int w = 0; w < nw; ++w) wlist.push_back(w);
    } else {
        // pivoted Cholesky for every frequency; the ones that turn out not to be positive semi-definite go to Jacobi
        SCLMD_CUDA(status.alloc(nw));
        auto ck = cplx ? k_chol<true> : k_chol<false>;
        if (use_smem) SCLMD_CUDA(cudaFuncSetAttribute(ck, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int slab = use_smem ? nw : (int)std::max<size_t>(1, std::min<size_t>(nw, ((size_t)2 << 30) / (nn * sizeof(double))));
        DevBuf<double> Lcol;
        if (!use_smem) SCLMD_CUDA(Lcol.alloc(nn * slab));
        const int threads = std::max(64, round_up(nc, 32));
        for (int w0 = 0; w0 < nw; w0 += slab) {
            const int cnt = std::min(slab, nw - w0);
            ck<<<cnt, threads, use_smem ? smem : 0, pl->st>>>(basis.p, nc, ncp, nterm, idx.p + (size_t)w0 * nterm, cre.p + (size_t)w0 * nterm,
                                                           cim.p + (size_t)w0 * nterm, Lcol.p, use_smem, Ldst + (size_t)w0 * nc * E * ncp,
                                                           status.p + w0);
            SCLMD_CUDA(cudaGetLastError());
            ++pl->launches;
        }
        std::vector<int> st_h(nw);
        SCLMD_CUDA(cudaMemcpyAsync(st_h.data(), status.p, nw * sizeof(int), cudaMemcpyDeviceToHost, pl->st));
        SCLMD_CUDA(cudaStreamSynchronize(pl->st));
        for (int w = 0; w < nw; ++w)
            if (st_h[w]) wlist.push_back(w);
        pl->nchol += nw - (int)wlist.size();
    }
    pl->njacobi += (int)wlist.size();
    if (wlist.empty()) return 0;
    const int nj = (int)wlist.size();
    SCLMD_CUDA(wlist_d.alloc(nj));
    SCLMD_CUDA(cudaMemcpy(wlist_d.p, wlist.data(), nj * sizeof(int), cudaMemcpyHostToDevice));
    // frequencies are processed in slabs so the scratch (A, and G when it does not fit in smem) stays bounded
    const int slab = (int)std::max<size_t>(1, std::min<size_t>(nj, ((size_t)2 << 30) / (nn * sizeof(double) * (use_smem ? 1 : 2))));
    SCLMD_CUDA(A.alloc(nn * slab));
    if (!use_smem) SCLMD_CUDA(G.alloc(nn * slab));
    auto kern = cplx ? k_factor<true> : k_factor<false>;
    if (use_smem) SCLMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = nc >= 64 ? 512 : 128;
    for (int w0 = 0; w0 < nj; w0 += slab) {
        const int cnt = std::min(slab, nj - w0);
        kern<<<cnt, threads, use_smem ? smem : 0, pl->st>>>(basis.p, nc, ncp, nterm, idx.p, cre.p, cim.p, A.p, G.p, use_smem, Ldst, evdst, wlist_d.p, w0);
        SCLMD_CUDA(cudaGetLastError());
        ++pl->launches;
    }
    SCLMD_CUDA(cudaStreamSynchronize(pl->st));
    return 0;
}

// Scratch of the generator (draws, spectrum, four-step intermediate, injected draws), cached per device between calls: cudaMalloc /
// cudaFree of multi-GB buffers cost ~10 ms per GB, more than the kernels that use them.  One generation at a time per process.
struct NoiseScratch {
    int device = -1;
    DevBuf<double> xi, X, xih;
    DevBuf<double2> four;
};
std::mutex g_scratch_mutex;
std::vector<std::unique_ptr<NoiseScratch>> g_scratch;
NoiseScratch &noise_scratch(int device) {
    for (auto &s : g_scratch)
        if (s->device == device) return *s;
    g_scratch.emplace_back(new NoiseScratch());
    g_scratch.back()->device = device;
    return *g_scratch.back();
}
template <typename T>
cudaError_t reserve(DevBuf<T> &b, size_t count) { return b.n >= count ? cudaSuccess : b.alloc_raw(count); }

// series for `ntraj` trajectories written to out[(n*out_nstride) + traj*out_tstride + c]
int generate(sclmd_noise_plan *pl, int ntraj, const double *xi_host, uint64_t seed, long long traj0, double *out, size_t out_tstride,
             size_t out_nstride, cudaStream_t st) {
    const int nw = pl->nw, nc = pl->nc, ncp = pl->ncp, N = pl->nmd, E = pl->cplx ? 2 : 1;
    const int ncx = ncp * E;                 // x row: real block [ncp] then imaginary block [ncp]
    const int imoff = pl->cplx ? ncp : 0;
    FftPlan full, p1, p2;
    int N1 = 0, N2 = 0;
    if (!make_fft_plan(N, full)) {
        set_error("noise: nmd=%d is not of the form 2^a 3^b 5^c (needed by the in-house FFT)", N);
        return SCLMD_ERR_ARG;
    }
    const size_t smem_inplace = (size_t)N * sizeof(double2);
    const bool direct = smem_inplace <= 200 * 1024;          // one in-place buffer: N <= 12800 (nmd = 8192 of configs 1, 2, 5 included)
    if (!direct) {
        // balanced 5-smooth split N = N1*N2, both transforms small enough to tile several per CTA
        int best = 1;
        for (int d = 1; (long long)d * d <= N; ++d)
            if (N % d == 0) best = d;
        N1 = best; N2 = N / best;
        if (!make_fft_plan(N1, p1) || !make_fft_plan(N2, p2) || (size_t)2 * N2 * sizeof(double2) > 200 * 1024) {
            set_error("noise: cannot split nmd=%d for the four-step FFT", N);
            return SCLMD_ERR_ARG;
        }
    }
    FftPlan big;
    const bool use_big = direct && !getenv("SCLMD_FFT_RADIX4") && make_big_plan(N, big) && (size_t)fft_pad(N) * sizeof(double2) <= 200 * 1024;
    if (direct && !pl->tw.p) {
        SCLMD_CUDA(pl->tw.alloc(N));
        SCLMD_CUDA(pl->perm.alloc(N));
        k_twiddle<<<cdiv(N, 256), 256, 0, st>>>(pl->tw.p, N);
        k_digit_reverse<<<cdiv(N, 256), 256, 0, st>>>(pl->perm.p, use_big ? big : full);
        SCLMD_CUDA(cudaGetLastError());
        pl->launches += 2;
    }
    const int npair = (nc + 1) / 2;
    // trajectory chunk: the scratch (draws + spectrum, + the four-step intermediate) takes at most a quarter of the free device memory
    // and at most 6 GB (cached between calls); whole 128-row tiles of the batched product when possible
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    NoiseScratch &sc = noise_scratch(pl->device);
    size_t free_b = 0, total_b = 0;
    SCLMD_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t have = (sc.xi.n + sc.X.n + sc.xih.n) * sizeof(double) + sc.four.n * sizeof(double2);
    const size_t budget = std::min<size_t>((free_b + have) / 4, (size_t)6 << 30);
    const size_t per_traj = (size_t)nw * (ncp + ncx) * sizeof(double) + (direct ? 0 : (size_t)npair * N * sizeof(double2)) +
                            (xi_host ? (size_t)nw * nc * sizeof(double) : 0);
    int chunk = (int)std::max<size_t>(1, std::min<size_t>(ntraj, budget / per_traj));
    if (chunk >= 128 && chunk < ntraj) chunk -= chunk % 128;
    DevBuf<double> &xi = sc.xi, &X = sc.X, &xih = sc.xih;
    DevBuf<double2> &scratch = sc.four;
    SCLMD_CUDA(reserve(xi, (size_t)nw * chunk * ncp));
    SCLMD_CUDA(reserve(X, (size_t)nw * chunk * ncx));        // the pad column of an odd nc is neither written nor read
    if (!direct) SCLMD_CUDA(reserve(scratch, (size_t)chunk * npair * N));
    if (xi_host) SCLMD_CUDA(reserve(xih, (size_t)chunk * nw * nc));
    const double scale = 1.0 / (pl->dt * N);   // dw/2pi (functions.py:51)
    std::vector<cudaEvent_t> evs;
    pl->stage_ms[0] = pl->stage_ms[1] = pl->stage_ms[2] = 0.0;
    auto mark = [&]() {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        evs.push_back(e);
    };
    int rc = 0;
    for (int t0 = 0; t0 < ntraj && !rc; t0 += chunk) {
        const int cnt = std::min(chunk, ntraj - t0);
        const size_t nel = (size_t)nw * cnt * (ncp / 2);
        const int blocks = (int)std::min<size_t>((nel + 255) / 256, (size_t)pl->nsm * 16);
        mark();
        const bool diag = pl->diagL && !pl->cplx && !getenv("SCLMD_NOISE_NO_DIAG");
        if (xi_host) {
            SCLMD_CUDA(cudaMemcpyAsync(xih.p, xi_host + (size_t)t0 * nw * nc, (size_t)cnt * nw * nc * sizeof(double), cudaMemcpyHostToDevice, st));
            k_scatter_xi<<<blocks, 256, 0, st>>>(xih.p, xi.p, nw, cnt, nc, ncp);
        } else if (!diag) {
            k_fill_xi<<<blocks, 256, 0, st>>>(xi.p, nw, cnt, nc, ncp, seed, traj0 + t0);
        }
        SCLMD_CUDA(cudaGetLastError());
        ++pl->launches;
        if (diag) {          // the draws (or the injected ones) times the factors' diagonals, straight into the spectrum rows
            k_fill_x_diag<<<blocks, 256, 0, st>>>(X.p, pl->Ld.p, xi_host ? xi.p : nullptr, nw, cnt, nc, ncp, seed, traj0 + t0);
            SCLMD_CUDA(cudaGetLastError());
        }
        mark();
        // X[w] (cnt x ncx) = xi[w] (cnt x ncp) . L[w]^T   -- one product per frequency
        if (diag) {
        } else if (tma_usable(cnt, nc)) {
            // persistent TMA / stream-K kernel over the whole frequency batch (dgemm_tma.cuh): operands as rank-3 tensors [w][row][k]
            for (int part = 0; part < E && !rc; ++part) {        // imaginary rows of L -> imaginary block [ncp, ncp+nc) of x
                TmaGemm g{};
                g.M = cnt; g.N = nc; g.K = ncp; g.nbatch = nw; g.nseg = 1;
                g.A = TmaOperand{xi.p, {(unsigned long long)ncp, (unsigned long long)cnt, (unsigned long long)nw},
                                 {(unsigned long long)ncp, (unsigned long long)cnt * ncp}, 0, 0};
                g.B = TmaOperand{pl->L.p + (size_t)part * nc * ncp, {(unsigned long long)ncp, (unsigned long long)nc, (unsigned long long)nw},
                                 {(unsigned long long)ncp, (unsigned long long)nc * E * ncp}, 0, 0};
                g.C = X.p + (size_t)part * ncp; g.ldc = (long long)nw * ncx; g.c_batch_stride = ncx; g.alpha = 1.0;      // X[traj][w][ncx]
                rc = launch_dgemm_tma(g, pl->tws, pl->nsm, st);
                ++pl->launches;
            }
            if (rc) break;
        } else {
            GemmArgs g{};
            g.M = cnt; g.N = nc; g.Kseg = ncp; g.nseg = nw; g.segs_per_split = 1;
            g.A = xi.p; g.lda = ncp; g.a_seg_stride = (long long)cnt * ncp; g.a_mod = 0;
            g.B = pl->L.p; g.ldb = ncp; g.b_seg_stride = (long long)nc * E * ncp; g.b_seg0 = 0;
            g.C = X.p; g.ldc = (long long)nw * ncx; g.c_split_stride = ncx; g.alpha = 1.0;                               // X[traj][w][ncx]
            for (int w0 = 0; w0 < nw; w0 += 32768) {   // gridDim.z limit
                const int wc = std::min(32768, nw - w0);
                g.nseg = wc;
                g.A = xi.p + (size_t)w0 * cnt * ncp;
                g.B = pl->L.p + (size_t)w0 * nc * E * ncp; g.C = X.p + (size_t)w0 * ncx;
                SCLMD_CUDA(launch_dgemm(g, wc, st));
                ++pl->launches;
                if (pl->cplx) {   // imaginary rows of L -> imaginary block [ncp, ncp+nc) of x
                    g.B += (size_t)nc * ncp; g.C += ncp;
                    SCLMD_CUDA(launch_dgemm(g, wc, st));
                    ++pl->launches;
                }
            }
        }
        mark();
        double *o = out + (size_t)t0 * out_tstride;
        if (use_big) {
            size_t smem_big = (size_t)(N + (N >> 4) + 1) * sizeof(double2);
            // real spectra: room for the staged half spectrum of the pair behind the transform buffer (nmd = 8192: 139 + 66 KB)
            const size_t smem_staged = smem_big + (size_t)(N / 2 + 2) * sizeof(double2);
            const int stage_ok = !pl->cplx && smem_staged <= 220 * 1024 && !getenv("SCLMD_FFT_NO_STAGE") ? 1 : 0;
            if (stage_ok) smem_big = smem_staged;
            SCLMD_CUDA(cudaFuncSetAttribute(k_fft_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big));
            k_fft_big<<<dim3(npair, cnt), std::min(256, std::max(32, N / 16)), smem_big, st>>>(X.p, big, pl->tw.p, pl->perm.p, nc, ncp, ncx, imoff, scale, o,
                                                                                           out_tstride, out_nstride, stage_ok);
            SCLMD_CUDA(cudaGetLastError());
            ++pl->launches;
        } else if (direct) {
            SCLMD_CUDA(cudaFuncSetAttribute(k_fft_inplace, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_inplace));
            k_fft_inplace<<<dim3(npair, cnt), N >= 2048 ? 512 : 128, smem_inplace, st>>>(X.p, full, pl->tw.p, pl->perm.p, cnt, nc, ncp, ncx, imoff, scale, o, out_tstride,
                                                                                      out_nstride);
            SCLMD_CUDA(cudaGetLastError());
            ++pl->launches;
        } else {
            const int tile1 = std::max(1, std::min(N2, 2048 / N1)), tile2 = std::max(1, std::min(N1, 2048 / N2));
            const size_t sm1 = (size_t)2 * tile1 * N1 * sizeof(double2), sm2 = (size_t)2 * tile2 * N2 * sizeof(double2);
            SCLMD_CUDA(cudaFuncSetAttribute(k_fft_step1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
            SCLMD_CUDA(cudaFuncSetAttribute(k_fft_step2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
            k_fft_step1<<<dim3(npair, cnt, cdiv(N2, tile1)), 256, sm1, st>>>(X.p, p1, N, N2, tile1, cnt, nc, ncx, imoff, scratch.p);
            SCLMD_CUDA(cudaGetLastError());
            k_fft_step2<<<dim3(npair, cnt, cdiv(N1, tile2)), 256, sm2, st>>>(scratch.p, p2, N, N1, tile2, nc, scale, o, out_tstride, out_nstride);
            SCLMD_CUDA(cudaGetLastError());
            pl->launches += 2;
        }
        mark();
    }
    const cudaError_t se = cudaStreamSynchronize(st);       // the cached scratch may be reused by the next call
    for (size_t i = 0; i + 3 < evs.size(); i += 4)
        for (int k = 0; k < 3; ++k) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, evs[i + k], evs[i + k + 1]) == cudaSuccess) pl->stage_ms[k] += ms;
        }
    for (cudaEvent_t e : evs) cudaEventDestroy(e);
    if (rc) return rc;
    SCLMD_CUDA(se);
    return 0;
}

}  // namespace

void sclmd::release_noise_scratch() {
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    for (auto &s : g_scratch)
        if (cudaSetDevice(s->device) == cudaSuccess) { s->xi.release(); s->X.release(); s->xih.release(); s->four.release(); }
    g_scratch.clear();
}

extern "C" {

int sclmd_noise_plan_create(int device, int nmd, double dt, int nc, int nbasis, const double *basis, int nterm, const int32_t *idx,
                            const double *cre, const double *cim, sclmd_noise_plan **out) {
    SCLMD_REQUIRE(out, "sclmd_noise_plan_create: out is NULL");
    *out = nullptr;
    SCLMD_REQUIRE(nmd > 0 && nmd % 2 == 0, "MyFFT.iFourier1D: array length error! (nmd must be even, noise.py:64,93)");
    SCLMD_REQUIRE(nc > 0 && nc <= 1024 && dt > 0, "sclmd_noise_plan_create: bad nc/dt");
    SCLMD_REQUIRE(nbasis > 0 && basis && nterm > 0 && nterm <= 4 && idx && cre, "sclmd_noise_plan_create: bad basis/terms");
    if (int e = select_device(device)) return e;
    std::unique_ptr<sclmd_noise_plan> pl(new sclmd_noise_plan());
    pl->device = device; pl->nmd = nmd; pl->dt = dt; pl->nc = nc; pl->ncp = round_up(nc, 2); pl->nw = nmd / 2 + 1;
    pl->nsm = sm_count(device);
    const int nw = pl->nw;
    bool cplx = false;
    if (cim)
        for (size_t i = 0; i < (size_t)nw * nterm; ++i)
            if (cim[i] != 0.0 && idx[i] >= 0) { cplx = true; break; }
    pl->cplx = cplx;
    for (size_t i = 0; i < (size_t)nw * nterm; ++i)
        SCLMD_REQUIRE(idx[i] < nbasis, "sclmd_noise_plan_create: basis index out of range");
    SCLMD_CUDA(cudaStreamCreateWithFlags(&pl->st, cudaStreamNonBlocking));
    const int E = cplx ? 2 : 1;
    SCLMD_CUDA(pl->L.alloc((size_t)nw * nc * E * pl->ncp));
    SCLMD_CUDA(pl->evals.alloc((size_t)nw * nc));
    cudaEvent_t f0 = nullptr, f1 = nullptr;
    SCLMD_CUDA(cudaEventCreate(&f0));
    SCLMD_CUDA(cudaEventCreate(&f1));
    SCLMD_CUDA(cudaEventRecord(f0, pl->st));
    // single real basis matrix for every frequency: factor once, scale per frequency
    bool single = !cplx;
    int b0 = -1;
    std::vector<double> cw(nw, 0.0);
    for (int w = 0; w < nw && single; ++w)
        for (int m = 0; m < nterm; ++m) {
            const int bi = idx[w * nterm + m];
            if (bi < 0 || cre[w * nterm + m] == 0.0) continue;
            if (b0 < 0) b0 = bi;
            if (bi != b0) { single = false; break; }
            cw[w] += cre[w * nterm + m];
        }
    if (single && b0 >= 0 && nw > 2) {
        const int one_idx[2] = {0, 0};
        const double cpos[2] = {1.0, -1.0};
        DevBuf<double> L2, ev2, V, lam, cwd;
        SCLMD_CUDA(L2.alloc((size_t)2 * nc * pl->ncp)); SCLMD_CUDA(ev2.alloc((size_t)2 * nc));
        SCLMD_CUDA(V.alloc((size_t)nc * pl->ncp)); SCLMD_CUDA(lam.alloc(nc)); SCLMD_CUDA(cwd.alloc(nw));
        const double *B0 = basis + (size_t)b0 * nc * nc;
        bool diagonal = true;
        for (int i = 0; i < nc && diagonal; ++i)
            for (int j = 0; j < nc; ++j)
                if (i != j && B0[(size_t)i * nc + j] != 0.0) { diagonal = false; break; }
        if (diagonal) {      // Debye friction, efric = I/damp: the eigenvectors are the unit vectors
            std::vector<double> Vh((size_t)nc * pl->ncp, 0.0), lh(nc);
            for (int i = 0; i < nc; ++i) { Vh[(size_t)i * pl->ncp + i] = 1.0; lh[i] = B0[(size_t)i * nc + i]; }
            SCLMD_CUDA(cudaMemcpyAsync(V.p, Vh.data(), Vh.size() * sizeof(double), cudaMemcpyHostToDevice, pl->st));
            SCLMD_CUDA(cudaMemcpyAsync(lam.p, lh.data(), nc * sizeof(double), cudaMemcpyHostToDevice, pl->st));
            SCLMD_CUDA(cudaStreamSynchronize(pl->st));
            // every L_w is diagonal: keep the diagonals (same expression as k_scale_factor) for the product-free generation path
            std::vector<double> ldh((size_t)nw * pl->ncp, 0.0);
            for (int w = 0; w < nw; ++w)
                for (int k = 0; k < nc; ++k) {
                    const double sv = cw[w] * lh[k];
                    ldh[(size_t)w * pl->ncp + k] = sv > 0 ? 1.0 * sqrt(sv) : 0.0;
                }
            SCLMD_CUDA(pl->Ld.alloc_raw(ldh.size()));
            SCLMD_CUDA(cudaMemcpy(pl->Ld.p, ldh.data(), ldh.size() * sizeof(double), cudaMemcpyHostToDevice));
            pl->diagL = true;
        } else {
            if (int e = factor_batch(pl.get(), 2, 1, B0, 1, one_idx, cpos, nullptr, false, L2.p, ev2.p, true)) return e;
            k_unit_vectors<<<1, 256, 0, pl->st>>>(L2.p, L2.p + (size_t)nc * pl->ncp, ev2.p, ev2.p + nc, nc, pl->ncp, V.p, lam.p);
            SCLMD_CUDA(cudaGetLastError());
        }
        SCLMD_CUDA(cudaMemcpyAsync(cwd.p, cw.data(), nw * sizeof(double), cudaMemcpyHostToDevice, pl->st));
        k_scale_factor<<<nw, 256, 0, pl->st>>>(V.p, lam.p, cwd.p, nc, pl->ncp, pl->L.p);
        SCLMD_CUDA(cudaGetLastError());
        SCLMD_CUDA(cudaStreamSynchronize(pl->st));
        pl->launches += 2;
    } else {
        if (int e = factor_batch(pl.get(), nw, nbasis, basis, nterm, idx, cre, cim, cplx, pl->L.p, pl->evals.p, false)) return e;
    }
    SCLMD_CUDA(cudaEventRecord(f1, pl->st));
    SCLMD_CUDA(cudaStreamSynchronize(pl->st));
    float fms = 0;
    if (cudaEventElapsedTime(&fms, f0, f1) == cudaSuccess) pl->factor_ms = fms;
    cudaEventDestroy(f0);
    cudaEventDestroy(f1);
    *out = pl.release();
    return SCLMD_OK;
}

int sclmd_noise_plan_destroy(sclmd_noise_plan *pl) {
    if (!pl) return SCLMD_OK;
    cudaSetDevice(pl->device);
    if (pl->st) { cudaStreamSynchronize(pl->st); cudaStreamDestroy(pl->st); }
    delete pl;
    return SCLMD_OK;
}

int sclmd_noise_plan_is_complex(sclmd_noise_plan *pl) { return pl ? pl->cplx : SCLMD_ERR_ARG; }

// L: [nw][nc][nc] real, or interleaved complex when the plan is complex
int sclmd_noise_plan_get_factors(sclmd_noise_plan *pl, double *L) {
    SCLMD_REQUIRE(pl && L, "sclmd_noise_plan_get_factors: NULL argument");
    SCLMD_CUDA(cudaSetDevice(pl->device));
    const int nc = pl->nc, ncp = pl->ncp, E = pl->cplx ? 2 : 1;
    std::vector<double> tmp(pl->L.n);
    SCLMD_CUDA(cudaMemcpy(tmp.data(), pl->L.p, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int w = 0; w < pl->nw; ++w)
        for (int i = 0; i < nc; ++i)
            for (int k = 0; k < nc; ++k) {
                const double *b = tmp.data() + (size_t)w * nc * E * ncp;
                if (E == 1) L[((size_t)w * nc + i) * nc + k] = b[(size_t)i * ncp + k];
                else {
                    L[(((size_t)w * nc + i) * nc + k) * 2] = b[(size_t)i * ncp + k];
                    L[(((size_t)w * nc + i) * nc + k) * 2 + 1] = b[(size_t)(nc + i) * ncp + k];
                }
            }
    return SCLMD_OK;
}

// inject factors (e.g. the reference's own V sqrt(lambda+)) for deterministic parity runs
int sclmd_noise_plan_set_factors(sclmd_noise_plan *pl, const double *L, int is_complex) {
    SCLMD_REQUIRE(pl && L, "sclmd_noise_plan_set_factors: NULL argument");
    SCLMD_CUDA(cudaSetDevice(pl->device));
    const int nc = pl->nc, ncp = pl->ncp, E = is_complex ? 2 : 1;
    pl->cplx = is_complex ? 1 : 0;
    pl->diagL = false;          // injected factors are general
    SCLMD_CUDA(pl->L.alloc((size_t)pl->nw * nc * E * ncp));
    std::vector<double> tmp(pl->L.n, 0.0);
    for (int w = 0; w < pl->nw; ++w)
        for (int i = 0; i < nc; ++i)
            for (int k = 0; k < nc; ++k) {
                double *b = tmp.data() + (size_t)w * nc * E * ncp;
                if (E == 1) b[(size_t)i * ncp + k] = L[((size_t)w * nc + i) * nc + k];
                else {
                    b[(size_t)i * ncp + k] = L[(((size_t)w * nc + i) * nc + k) * 2];
                    b[(size_t)(nc + i) * ncp + k] = L[(((size_t)w * nc + i) * nc + k) * 2 + 1];
                }
            }
    SCLMD_CUDA(cudaMemcpy(pl->L.p, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice));
    return SCLMD_OK;
}

int sclmd_noise_plan_generate(sclmd_noise_plan *pl, int ntraj, const double *xi, uint64_t seed, int64_t traj0, double *out) {
    SCLMD_REQUIRE(pl && out && ntraj > 0, "sclmd_noise_plan_generate: bad arguments");
    SCLMD_CUDA(cudaSetDevice(pl->device));
    DevBuf<double> o;   // [ntraj][nmd][nc]
    SCLMD_CUDA(o.alloc((size_t)ntraj * pl->nmd * pl->nc));
    if (int e = generate(pl, ntraj, xi, seed, traj0, o.p, (size_t)pl->nmd * pl->nc, (size_t)pl->nc, pl->st)) return e;
    SCLMD_CUDA(cudaMemcpy(out, o.p, o.n * sizeof(double), cudaMemcpyDeviceToHost));
    return SCLMD_OK;
}

int64_t sclmd_noise_plan_launch_count(sclmd_noise_plan *pl) { return pl ? pl->launches : -1; }

// ms[6]: device milliseconds of the last generate call per stage (0 draws, 1 x = L xi, 2 transform) and of the factorisation
// at plan creation (3); number of frequencies factorised by pivoted Cholesky (4) and by Jacobi (5)
int sclmd_noise_plan_get_profile(sclmd_noise_plan *pl, double *ms) {
    SCLMD_REQUIRE(pl && ms, "sclmd_noise_plan_get_profile: NULL argument");
    for (int i = 0; i < 3; ++i) ms[i] = pl->stage_ms[i];
    ms[3] = pl->factor_ms;
    ms[4] = pl->nchol;
    ms[5] = pl->njacobi;
    return SCLMD_OK;
}

// device-to-device variant used by md.cu: writes into a trajectory-major table [ntraj_total][nmd][ncp_table]
int sclmd_noise_plan_generate_into(sclmd_noise_plan *pl, int ntraj, uint64_t seed, int64_t traj0, double *table, int ntraj_total,
                                   int ncp_table, int traj_offset) {
    SCLMD_REQUIRE(pl && table && traj_offset >= 0 && traj_offset + ntraj <= ntraj_total, "sclmd_noise_plan_generate_into: bad arguments");
    SCLMD_CUDA(cudaSetDevice(pl->device));
    return generate(pl, ntraj, nullptr, seed, traj0, table + (size_t)traj_offset * pl->nmd * ncp_table, (size_t)pl->nmd * ncp_table,
                    (size_t)ncp_table, pl->st);
}

int sclmd_noise_plan_dims(sclmd_noise_plan *pl, int *nmd, int *nc) {
    SCLMD_REQUIRE(pl, "sclmd_noise_plan_dims: NULL plan");
    if (nmd) *nmd = pl->nmd;
    if (nc) *nc = pl->nc;
    return SCLMD_OK;
}

// out[nt][m] = alpha * sum_i table(tl_t, wl_i; eta) * giT[m][i]
// functions.myfft / numpy.fft with the reference's conventions left to the caller: out[f][m] = scale * sum_k in[f][k] exp(sign 2 pi i k m / n)
int sclmd_fft(int device, int n, int batch, const double *re, const double *im, int sign, double scale, double *out_re, double *out_im) {
    SCLMD_REQUIRE(n > 0 && batch > 0 && re && out_re && out_im && (sign == 1 || sign == -1), "sclmd_fft: bad arguments");
    if (int e = select_device(device)) return e;
    FftPlan full, p1, p2;
    SCLMD_REQUIRE(make_fft_plan(n, full), "sclmd_fft: n=%d is not of the form 2^a 3^b 5^c (in-house radix-2/3/4/5 FFT)", n);
    const size_t tot = (size_t)n * batch;
    DevBuf<double> dre, dim, ore, oim;
    SCLMD_CUDA(dre.alloc(tot)); SCLMD_CUDA(ore.alloc(tot)); SCLMD_CUDA(oim.alloc(tot));
    SCLMD_CUDA(cudaMemcpy(dre.p, re, tot * sizeof(double), cudaMemcpyHostToDevice));
    if (im) {
        SCLMD_CUDA(dim.alloc(tot));
        SCLMD_CUDA(cudaMemcpy(dim.p, im, tot * sizeof(double), cudaMemcpyHostToDevice));
    }
    const double cj = sign < 0 ? 1.0 : -1.0;
    const size_t smem_direct = (size_t)2 * n * sizeof(double2);
    if (smem_direct <= 200 * 1024) {
        SCLMD_CUDA(cudaFuncSetAttribute(k_c2c_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_direct));
        k_c2c_direct<<<batch, 256, smem_direct>>>(dre.p, dim.p, full, cj, scale, ore.p, oim.p);
        SCLMD_CUDA(cudaGetLastError());
    } else {
        int best = 1;
        for (int d = 1; (long long)d * d <= n; ++d)
            if (n % d == 0) best = d;
        const int N1 = best, N2 = n / best;
        SCLMD_REQUIRE(make_fft_plan(N1, p1) && make_fft_plan(N2, p2) && (size_t)2 * N2 * sizeof(double2) <= 200 * 1024,
                      "sclmd_fft: cannot split n=%d for the four-step transform", n);
        DevBuf<double2> scratch;
        SCLMD_CUDA(scratch.alloc(tot));
        const int tile1 = std::max(1, std::min(N2, 2048 / N1)), tile2 = std::max(1, std::min(N1, 2048 / N2));
        const size_t sm1 = (size_t)2 * tile1 * N1 * sizeof(double2), sm2 = (size_t)2 * tile2 * N2 * sizeof(double2);
        SCLMD_CUDA(cudaFuncSetAttribute(k_c2c_step1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
        SCLMD_CUDA(cudaFuncSetAttribute(k_c2c_step2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        k_c2c_step1<<<dim3(batch, cdiv(N2, tile1)), 256, sm1>>>(dre.p, dim.p, p1, n, N2, tile1, cj, scratch.p);
        SCLMD_CUDA(cudaGetLastError());
        k_c2c_step2<<<dim3(batch, cdiv(N1, tile2)), 256, sm2>>>(scratch.p, p2, n, N1, tile2, cj, scale, ore.p, oim.p);
        SCLMD_CUDA(cudaGetLastError());
    }
    SCLMD_CUDA(cudaDeviceSynchronize());
    SCLMD_CUDA(cudaMemcpy(out_re, ore.p, tot * sizeof(double), cudaMemcpyDeviceToHost));
    SCLMD_CUDA(cudaMemcpy(out_im, oim.p, tot * sizeof(double), cudaMemcpyDeviceToHost));
    return SCLMD_OK;
}

int sclmd_cos_transform(int device, int nt, int nw, int m, const double *tl, const double *wl, const double *giT, double eta,
                        double alpha, double *out) {
    SCLMD_REQUIRE(nt > 0 && nw > 0 && m > 0 && tl && wl && giT && out, "sclmd_cos_transform: bad arguments");
    if (int e = select_device(device)) return e;
    const int nwp = round_up(nw, 2);
    DevBuf<double> dtl, dwl, C, B, O;
    SCLMD_CUDA(dtl.alloc(nt)); SCLMD_CUDA(dwl.alloc(nw)); SCLMD_CUDA(C.alloc((size_t)nt * nwp));
    const int mp = round_up(m, 2);
    SCLMD_CUDA(B.alloc((size_t)m * nwp)); SCLMD_CUDA(O.alloc((size_t)nt * mp));
    SCLMD_CUDA(cudaMemcpy(dtl.p, tl, nt * sizeof(double), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(dwl.p, wl, nw * sizeof(double), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy2D(B.p, nwp * sizeof(double), giT, nw * sizeof(double), nw * sizeof(double), m, cudaMemcpyHostToDevice));
    k_cos_table<<<std::min(4096, cdiv(nt * nwp, 256)), 256>>>(dtl.p, dwl.p, nt, nw, nwp, eta, C.p);
    SCLMD_CUDA(cudaGetLastError());
    GemmArgs g{};
    g.M = nt; g.N = m; g.Kseg = nwp; g.nseg = 1; g.segs_per_split = 1;
    g.A = C.p; g.lda = nwp; g.B = B.p; g.ldb = nwp; g.C = O.p; g.ldc = mp;
    g.alpha = alpha;
    SCLMD_CUDA(launch_dgemm(g, 1, 0));
    SCLMD_CUDA(cudaDeviceSynchronize());
    SCLMD_CUDA(cudaMemcpy2D(out, m * sizeof(double), O.p, mp * sizeof(double), m * sizeof(double), nt, cudaMemcpyDeviceToHost));
    return SCLMD_OK;
}

int sclmd_gamt(int device, int nt, int nw, int m, const double *tl, const double *wl, const double *giT, double *out) {
    SCLMD_REQUIRE(nw > 0 && wl, "sclmd_gamt: bad arguments");
    return sclmd_cos_transform(device, nt, nw, m, tl, wl, giT, 0.0, 2.0 / nw * wl[nw - 1] / 3.14159265358979323846, out);   // 2*mean(...)*wl[-1]/pi
}

}  // extern "C"
