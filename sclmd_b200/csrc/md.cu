// Ensemble velocity-Verlet integrator for the semi-classical generalized Langevin equation.
//
// Replaces md.vv / md.force / md.potforce(harmonic) / {ebath,phbath}.bforce of the reference
// (sclmd/md.py:367-474, sclmd/baths.py:224-255,448-458) for `ntraj` independent noise
// realisations.  Single-tail restatement: per bath the friction tail
//   S = dt * sum_{j>=1} kernel[j] . p_{t+1-j}[cids]
// is contracted ONCE per step from an HBM ring buffer of p[cids] and reused by force
// evaluations B, C of step t and A of step t+1.
//
// HBM layout (row-major, leading dims padded to even so every row is 16-byte aligned)
//   q,p,G,...   [ntraj][ld]            ld  = nph rounded up to 2, pads stay 0
//   K           [nph][ld]
//   ring_b      [ntraj][ml_b][ncp_b]   slot s holds p_{t'}[cids] with t' mod ml_b == s
//   kernel_b    diag [ml_b+1][ncp_b] (last row 0)   |   full [ml_b][nc_b][ncp_b]
//   noise_b     [ntraj][nmd][ncp_b]    trajectory-major: a trajectory's series is one contiguous block (the generator's transform
//                                      writes it row by row; a time-major table would put every row on another page)
//   cur_b,etot  [nmd][ntraj]
//   tailp_b     [nsplit_b][ntraj][ncp_b]   partial tails, summed in fixed order by consumers
#include <algorithm>
#include <cstdlib>
#include <memory>
#include <type_traits>

#include <cooperative_groups.h>

#include "common.cuh"
#include "dgemm.cuh"
#include "dgemm_tma.cuh"

using namespace sclmd;

namespace {

constexpr int MAXB = 8;

struct BathDev {  // POD view passed to kernels
    int nc, ncp, ml, nsplit, diag, has_lin, use_tail;
    double c0;
    const int *inv;          // [nph] -> position in bath or -1
    const int *cids;         // [nc]
    const double *k0;        // diag: kernel row 0 [ncp]
    const double *noise;     // [ntraj][nmd][ncp]
    int nmd;
    const double *tailp;     // [nsplit][ntraj][ncp]
    const double *lin;       // [ntraj][ncp]
    double *ring;            // [ntraj][ml][ncp]
    double *cur;             // [nmd][ntraj]
    double *fa, *fc;         // [ntraj][ncp] bath force of evaluation A (md.fhis) / of evaluation C (md.fbaths after vv), or NULL
    const double *wt;        // has_lin: W^T [Kw][ncp] (the ensemble kernel's matrix-vector product reads it lane-contiguous)
    int Kw;
};
struct BathSet {
    int nb;
    BathDev b[MAXB];
};

// segment of the tensor-pipe far pass (see k_tail_far_mma / plan_far)
constexpr int FM_SPC = 2;                        // segments per CTA (a CTA's contiguous range crosses at most one pair boundary)
struct FarSeg {
    int c0, traj0, a_lo, a_hi, slot, nlive;      // nlive: warps with real dofs (equal for the segments of one CTA)
};
struct Bath {
    int nc = 0, ncp = 0, ml = 1, kind = 0, nsplit = 1, Kw = 0, gemm_cfg = -1;
    bool has_lin = false, has_extra = false;
    double c0 = 1.0;
    DevBuf<int> cids, inv;
    DevBuf<double> kern, W, ring, xq, lin, tailp, noise, cur, far, fa, fc, WT, rowstage;
    DevBuf<double> kT;        // tensor-pipe far pass: transposed kernel table [ncp][ldk] and the tensor map of the ring
    int ldk = 0, far_used = 1, far_cap = 1;
    CUtensorMap ringmap;
    bool mma_ready = false;
    // tensor-pipe far pass, spread over the steps of a block: segment plan (FarPlan), two halves of partial far tails (the block in use /
    // the next block, accumulated one slice per step), per-pair partial counts for the near pass
    DevBuf<FarSeg> fsegs, fmid;
    DevBuf<int> fnslot;
    std::vector<int> slice0;          // first CTA of slice i in fsegs (FM_SLICES + 1 entries)
    int fchunks = 0, fnpairs = 0, fslots = 0, fcur = 0, fnmid = 0;
    size_t fhalf = 0;                 // doubles per half: fslots * 2 TB * ntraj * ncp
    long long nxt_block = -1;         // block start the other half is being filled for, and how many of its slices are done
    int nxt_done = 0;
    bool blocked = false;     // time-blocked tails (diagonal kernel, long memory)
    int far_nsplit = 1;
    long long far_t0 = -1;    // block start the far tails in `far` belong to
    std::vector<int> cids_h;  // host copy of the dof list (disjointness test of the modal mode)
};

// ---------------------------------------------------------------- kernels
__device__ __forceinline__ double bath_force(const BathDev &b, int traj, int ntraj, int a, int slab, double x) {
    double fb = b.noise[((size_t)traj * b.nmd + slab) * b.ncp + a];
    if (b.diag) fb -= b.c0 * b.k0[a] * x;
    if (b.has_lin) fb += b.lin[(size_t)traj * b.ncp + a];
    if (b.use_tail) {
        double s = 0.0;
        for (int z = 0; z < b.nsplit; ++z) s += b.tailp[((size_t)z * ntraj + traj) * b.ncp + a];
        fb -= s;
    }
    return fb;
}

// evaluation A (md.py:383-398): etot, ring push, f_A, p_half, q_next, heat current
template <int NBATH>      // baths the per-element loop is unrolled for (register pressure: 128 registers at 8, 1/4 occupancy)
__global__ void __launch_bounds__(256, NBATH <= 2 ? 4 : NBATH <= 4 ? 2 : 1) k_phase_a(BathSet bs, int nph, int ld, int ntraj, int nmd, long long t, double dt,
                                                  const double *__restrict__ q, const double *__restrict__ p,
                                                  const double *__restrict__ G, int gsplit, size_t gstride,
                                                  const double *__restrict__ Dcorr,
                                                  double *__restrict__ phalf, double *__restrict__ qn, double *__restrict__ etot) {
    __shared__ double red[32];
    const int traj = blockIdx.x;
    const size_t row = (size_t)traj * ld;
    const int slab = (int)(t % nmd);
    double ke = 0.0, cur[NBATH];
#pragma unroll
    for (int b = 0; b < NBATH; ++b) cur[b] = 0.0;
    // two consecutive elements per thread and trip, moved as 16-byte accesses (rows start 16-byte aligned: ld is even); the
    // fused kernel k_phase_bca walks the elements the same way, so both produce bit-identical partial sums
    for (int i0 = 2 * threadIdx.x; i0 < nph; i0 += 2 * blockDim.x) {
        const double2 p2 = *reinterpret_cast<const double2 *>(p + row + i0), q2 = *reinterpret_cast<const double2 *>(q + row + i0);
        double2 g2 = *reinterpret_cast<const double2 *>(G + row + i0);
        for (int z = 1; z < gsplit; ++z) {               // K-slices of K.q, fixed order
            const double2 v = *reinterpret_cast<const double2 *>(G + (size_t)z * gstride + row + i0);
            g2.x += v.x;
            g2.y += v.y;
        }
        if (Dcorr) {                                     // K.constrain(q') = K.q' - K[:,c].q'[c]
            const double2 v = *reinterpret_cast<const double2 *>(Dcorr + row + i0);
            g2.x -= v.x;
            g2.y -= v.y;
        }
        double2 ph2 = make_double2(0.0, 0.0), qn2 = make_double2(0.0, 0.0);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = i0 + e;
            if (i >= nph) break;
            const double pi = e ? p2.y : p2.x, qi = e ? q2.y : q2.x;
            double f = -(e ? g2.y : g2.x);
#pragma unroll
            for (int b = 0; b < NBATH; ++b) {
                if (b < bs.nb) {
                    const int a = bs.b[b].inv[i];
                    if (a >= 0) {
                        const double fb = bath_force(bs.b[b], traj, ntraj, a, slab, pi);
                        cur[b] += fb * pi;
                        f += fb;
                        if (bs.b[b].fa) bs.b[b].fa[(size_t)traj * bs.b[b].ncp + a] = fb;
                        bs.b[b].ring[((size_t)traj * bs.b[b].ml + (int)(t % bs.b[b].ml)) * bs.b[b].ncp + a] = pi;
                    }
                }
            }
            ke += 0.5 * pi * pi;
            const double phv = pi + f * dt / 2.0, qnv = qi + pi * dt + f * dt * dt / 2.0;
            if (e) { ph2.y = phv; qn2.y = qnv; } else { ph2.x = phv; qn2.x = qnv; }
        }
        *reinterpret_cast<double2 *>(phalf + row + i0) = ph2;
        *reinterpret_cast<double2 *>(qn + row + i0) = qn2;
    }
    ke = block_sum(ke, red);
    if (threadIdx.x == 0) etot[(size_t)slab * ntraj + traj] = ke;
#pragma unroll
    for (int b = 0; b < NBATH; ++b) {
        if (b < bs.nb) {
            const double c = block_sum(cur[b], red);
            if (threadIdx.x == 0) bs.b[b].cur[(size_t)slab * ntraj + traj] = c;
        }
    }
}

// evaluation B or C (md.py:401-404): pout = phalf + dt/2 * F(t+1; x, qn); `final` applies the
// constraint (md.py:407-408) and commits q.
template <int NBATH>
__global__ void __launch_bounds__(256, NBATH <= 2 ? 4 : NBATH <= 4 ? 2 : 1) k_phase_bc(BathSet bs, int nph, int ld, int ntraj, int nmd, long long t,
                                                                      const long long *__restrict__ tptr, double dt,
                                                   const double *__restrict__ x, const double *__restrict__ phalf,
                                                   const double *__restrict__ Gn, int gsplit, size_t gstride, double *__restrict__ pout,
                                                   const double *__restrict__ qn, double *__restrict__ qout,
                                                   const unsigned char *__restrict__ cons, int final, int fused,
                                                   double *__restrict__ fout) {
    const int traj = blockIdx.x;
    const size_t row = (size_t)traj * ld;
    if (tptr) t = *tptr;                    // captured in a CUDA graph: the step counter lives on the device
    const int slab = (int)((t + 1) % nmd);
    for (int i = threadIdx.x; i < nph; i += blockDim.x) {
        const double ph = phalf[row + i];
        double g = Gn[row + i];
        for (int z = 1; z < gsplit; ++z) g += Gn[(size_t)z * gstride + row + i];
        double xi = fused ? ph : x[row + i];
        double pnew = 0.0, f = 0.0;
        const int reps = fused ? 2 : 1;  // all baths time-local & diagonal: B and C in registers
        for (int r = 0; r < reps; ++r) {
            f = -g;
#pragma unroll
            for (int b = 0; b < NBATH; ++b) {
                if (b < bs.nb) {
                    const int a = bs.b[b].inv[i];
                    if (a >= 0) {
                        const double fb = bath_force(bs.b[b], traj, ntraj, a, slab, xi);
                        f += fb;
                        if (final && bs.b[b].fc) bs.b[b].fc[(size_t)traj * bs.b[b].ncp + a] = fb;   // the last evaluation wins
                    }
                }
            }
            pnew = ph + dt * f / 2.0;
            xi = pnew;
        }
        if (final) {
            double qv = qn[row + i];
            if (cons && cons[i]) {
                pnew = 0.0;
                qv = 0.0;
            }
            qout[row + i] = qv;
            if (fout) fout[row + i] = f;     // md.f: the force of evaluation C (md.py:403,411)
        }
        pout[row + i] = pnew;
    }
}

// Evaluations B and C of step t fused with evaluation A of step t+1 (no constraints, every bath force diagonal in x): the
// three evaluations see the same K.q', noise row and history tail and differ only in the momentum they are taken at, so the
// state never leaves registers between them.  Saves one launch and the re-read of p, q and the K-slices of K.q per step.
template <int NBATH>
__global__ void __launch_bounds__(256, NBATH <= 2 ? 4 : NBATH <= 4 ? 2 : 1) k_phase_bca(BathSet bs, int nph, int ld, int ntraj, int nmd, long long t, double dt,
                                                    double *__restrict__ phalf, const double *__restrict__ Gn, int gsplit, size_t gstride,
                                                    double *__restrict__ qn, double *__restrict__ etot) {
    __shared__ double red[32];
    const int traj = blockIdx.x;
    const size_t row = (size_t)traj * ld;
    const int slab = (int)((t + 1) % nmd);
    double ke = 0.0, cur[NBATH];
#pragma unroll
    for (int b = 0; b < NBATH; ++b) cur[b] = 0.0;
    for (int i0 = 2 * threadIdx.x; i0 < nph; i0 += 2 * blockDim.x) {      // element pairs, as in k_phase_a
        const double2 ph2 = *reinterpret_cast<const double2 *>(phalf + row + i0), qv2 = *reinterpret_cast<const double2 *>(qn + row + i0);
        double2 g2 = *reinterpret_cast<const double2 *>(Gn + row + i0);
        for (int z = 1; z < gsplit; ++z) {
            const double2 v = *reinterpret_cast<const double2 *>(Gn + (size_t)z * gstride + row + i0);
            g2.x += v.x;
            g2.y += v.y;
        }
        double2 pho = make_double2(0.0, 0.0), qno = make_double2(0.0, 0.0);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = i0 + e;
            if (i >= nph) break;
            const double ph = e ? ph2.y : ph2.x, g = e ? g2.y : g2.x, qv = e ? qv2.y : qv2.x;
            int a[NBATH];
#pragma unroll
            for (int b = 0; b < NBATH; ++b) a[b] = b < bs.nb ? bs.b[b].inv[i] : -1;
            double xi = ph, pnew = 0.0;
#pragma unroll
            for (int r = 0; r < 2; ++r) {                  // evaluations B and C (md.py:401-404)
                double f = -g;
#pragma unroll
                for (int b = 0; b < NBATH; ++b)
                    if (a[b] >= 0) f += bath_force(bs.b[b], traj, ntraj, a[b], slab, xi);
                pnew = ph + dt * f / 2.0;
                xi = pnew;
            }
            // (p_{t+1}, q_{t+1}) = (pnew, qv) are not stored: the only consumers of the state arrays are flush() and the unfused
            // kernels, and a flush recomputes the then-current state from phalf / qn.  Two array passes less per step.
            // evaluation A of step t+1 (md.py:383-398) at (p_{t+1}, q_{t+1}) = (pnew, qv); K.q_{t+1} = K.q' without constraints
            double f = -g;
#pragma unroll
            for (int b = 0; b < NBATH; ++b) {
                if (a[b] >= 0) {
                    const double fb = bath_force(bs.b[b], traj, ntraj, a[b], slab, pnew);
                    cur[b] += fb * pnew;
                    f += fb;
                    bs.b[b].ring[((size_t)traj * bs.b[b].ml + (int)((t + 1) % bs.b[b].ml)) * bs.b[b].ncp + a[b]] = pnew;
                }
            }
            ke += 0.5 * pnew * pnew;
            const double phv = pnew + f * dt / 2.0, qnv = qv + pnew * dt + f * dt * dt / 2.0;
            if (e) { pho.y = phv; qno.y = qnv; } else { pho.x = phv; qno.x = qnv; }
        }
        *reinterpret_cast<double2 *>(phalf + row + i0) = pho;
        *reinterpret_cast<double2 *>(qn + row + i0) = qno;
    }
    ke = block_sum(ke, red);
    if (threadIdx.x == 0) etot[(size_t)slab * ntraj + traj] = ke;
#pragma unroll
    for (int b = 0; b < NBATH; ++b) {
        if (b < bs.nb) {
            const double c = block_sum(cur[b], red);
            if (threadIdx.x == 0) bs.b[b].cur[(size_t)slab * ntraj + traj] = c;
        }
    }
}

// ---- persistent step kernel for small systems -------------------------------------------------------------------------
// A single trajectory (or a handful) of a few hundred dofs with time-local diagonal baths -- the reference's own example
// (examples/runmd.py: 603 dofs, two ml = 1 baths) -- is latency-bound as a chain of launches (~50 us per step).  Here ONE
// cooperative launch advances nsteps steps: every CTA keeps the whole state (q, p, K.q, p_half, q') of all trajectories in
// shared memory and does the elementwise parts of a step redundantly (identical arithmetic, so identical values everywhere);
// only the matrix-vector products K.q' and K.constrain(q') are distributed (warp per row of K, which stays in L2) and exchanged
// through a double-buffered global vector with ONE grid-wide barrier per step.
constexpr int PS_MAXT = 2;      // trajectories the persistent kernel keeps in registers
struct PersistArgs {
    BathSet bs;
    int nph, ld, ntraj, nmd, has_cons;
    long long t0, nsteps;
    double dt;
    const double *K;
    double *q, *p, *G;             // state in / out; G = K.q of the current q (slice 0 of the handle's buffer)
    const unsigned char *cons;
    double *etot;
    double *xbuf;                  // [2][2][ntraj][ld]: (K.q', K.constrain(q')) of the step, double-buffered over steps
};

// A thread owns the same <= PS_EPT elements of every trajectory for the whole run: their bath index, friction coefficient and
// constraint flag sit in registers, and the noise row of step t+1 (needed by evaluations B, C and by evaluation A of the next
// step) is requested before the matrix-vector product so that it arrives under it.
constexpr int PS_EPT = 4;          // elements per thread: nph <= 1024
template <int NBATH, int NT>
__global__ void __launch_bounds__(256, 1) k_md_persist(const PersistArgs a) {
    extern __shared__ double psm[];
    __shared__ double red[32];
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int nph = a.nph, ld = a.ld;
    double *sn = psm;                                   // q' of every trajectory: the operand of the matrix-vector product
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const double dt = a.dt;
    double q[NT][PS_EPT], p[NT][PS_EPT], g[NT][PS_EPT], nz[NT][PS_EPT][NBATH], kf[PS_EPT][NBATH];
    int bc[PS_EPT][NBATH];
    bool fix[PS_EPT];
    const int slab0 = (int)(a.t0 % a.nmd);
#pragma unroll
    for (int j = 0; j < PS_EPT; ++j) {
        const int i = tid + j * 256;
        const bool ok = i < nph;
        fix[j] = ok && a.has_cons && a.cons[i];
#pragma unroll
        for (int b = 0; b < NBATH; ++b) {
            bc[j][b] = (ok && b < a.bs.nb) ? a.bs.b[b].inv[i] : -1;
            kf[j][b] = bc[j][b] >= 0 ? a.bs.b[b].c0 * a.bs.b[b].k0[bc[j][b]] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            q[k][j] = ok ? a.q[(size_t)k * ld + i] : 0.0;
            p[k][j] = ok ? a.p[(size_t)k * ld + i] : 0.0;
            g[k][j] = ok ? a.G[(size_t)k * ld + i] : 0.0;
#pragma unroll
            for (int b = 0; b < NBATH; ++b)
                nz[k][j][b] = bc[j][b] >= 0 ? a.bs.b[b].noise[((size_t)k * a.nmd + slab0) * a.bs.b[b].ncp + bc[j][b]] : 0.0;
        }
    }
    for (long long s = 0; s < a.nsteps; ++s) {
        const long long t = a.t0 + s;
        const int slab = (int)(t % a.nmd), slab1 = (int)((t + 1) % a.nmd);
        double ph[NT][PS_EPT], qn[NT][PS_EPT], n1[NT][PS_EPT][NBATH];
        // ---- evaluation A (md.py:383-398), redundantly in every CTA; CTA 0 records the observables and pushes the history
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            double ke = 0.0, cur[NBATH];
#pragma unroll
            for (int b = 0; b < NBATH; ++b) cur[b] = 0.0;
#pragma unroll
            for (int j = 0; j < PS_EPT; ++j) {
                const int i = tid + j * 256;
                double f = -g[k][j];
#pragma unroll
                for (int b = 0; b < NBATH; ++b)
                    if (bc[j][b] >= 0) {
                        const double fb = nz[k][j][b] - kf[j][b] * p[k][j];
                        cur[b] += fb * p[k][j];
                        f += fb;
                        if (blockIdx.x == 0)
                            a.bs.b[b].ring[((size_t)k * a.bs.b[b].ml + (int)(t % a.bs.b[b].ml)) * a.bs.b[b].ncp + bc[j][b]] = p[k][j];
                    }
                ke += 0.5 * p[k][j] * p[k][j];
                ph[k][j] = p[k][j] + f * dt / 2.0;
                qn[k][j] = q[k][j] + p[k][j] * dt + f * dt * dt / 2.0;
                if (i < nph) sn[(size_t)k * ld + i] = qn[k][j];
                // the noise of step t+1: in flight during the matrix-vector product and the grid barrier
#pragma unroll
                for (int b = 0; b < NBATH; ++b)
                    n1[k][j][b] = bc[j][b] >= 0 ? a.bs.b[b].noise[((size_t)k * a.nmd + slab1) * a.bs.b[b].ncp + bc[j][b]] : 0.0;
            }
            if (blockIdx.x == 0) {
                ke = block_sum(ke, red);
                if (tid == 0) a.etot[(size_t)slab * NT + k] = ke;
#pragma unroll
                for (int b = 0; b < NBATH; ++b)
                    if (b < a.bs.nb) {
                        const double c = block_sum(cur[b], red);
                        if (tid == 0) a.bs.b[b].cur[(size_t)slab * NT + k] = c;
                    }
            }
        }
        __syncthreads();
        // ---- K.q' and K.constrain(q'): warp per row, rows dealt round-robin over the grid
        double *xb = a.xbuf + (size_t)(s & 1) * 2 * NT * ld;
        for (int r = blockIdx.x * nwarp + warp; r < nph; r += gridDim.x * nwarp) {
            const double *krow = a.K + (size_t)r * ld;
            double acc[NT], accc[NT];
#pragma unroll
            for (int k = 0; k < NT; ++k) acc[k] = accc[k] = 0.0;
            for (int c = lane; c < nph; c += 32) {
                const double kv = krow[c];
                const bool fx = a.has_cons && a.cons[c];
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    const double x = sn[(size_t)k * ld + c];
                    acc[k] = fma(kv, x, acc[k]);
                    if (!fx) accc[k] = fma(kv, x, accc[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                const double v = warp_sum(acc[k]), vc = warp_sum(accc[k]);
                if (lane == 0) {
                    xb[(size_t)k * ld + r] = v;
                    xb[(size_t)(NT + k) * ld + r] = vc;
                }
            }
        }
        grid.sync();
        // ---- evaluations B and C (md.py:401-404), constraint (md.py:407-408); K.q of the next step
#pragma unroll
        for (int k = 0; k < NT; ++k)
#pragma unroll
            for (int j = 0; j < PS_EPT; ++j) {
                const int i = tid + j * 256;
                if (i >= nph) continue;
                const double gn = __ldcg(xb + (size_t)k * ld + i), gc = __ldcg(xb + (size_t)(NT + k) * ld + i);
                double xi = ph[k][j], pnew = 0.0;
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
                    double f = -gn;
#pragma unroll
                    for (int b = 0; b < NBATH; ++b)
                        if (bc[j][b] >= 0) f += n1[k][j][b] - kf[j][b] * xi;
                    pnew = ph[k][j] + dt * f / 2.0;
                    xi = pnew;
                }
                p[k][j] = fix[j] ? 0.0 : pnew;
                q[k][j] = fix[j] ? 0.0 : qn[k][j];
                g[k][j] = a.has_cons ? gc : gn;
#pragma unroll
                for (int b = 0; b < NBATH; ++b) nz[k][j][b] = n1[k][j][b];
            }
        __syncthreads();      // everyone is done with sn before the next evaluation A rewrites it
    }
    if (blockIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NT; ++k)
#pragma unroll
            for (int j = 0; j < PS_EPT; ++j) {
                const int i = tid + j * 256;
                if (i < nph) {
                    a.q[(size_t)k * ld + i] = q[k][j];
                    a.p[(size_t)k * ld + i] = p[k][j];
                    a.G[(size_t)k * ld + i] = g[k][j];
                }
            }
    }
}

// ---- ensemble-persistent step kernel ----------------------------------------------------------------------------------------
// Many trajectories of a small system with time-local diagonal baths (BASELINE configs[1]: 1024 trajectories of the 603-dof
// junction of examples/runmd.py).  Trajectories never interact, so a CTA takes EN_T = 8 of them -- the row count of the FP64
// tensor tile -- through ALL steps of a run with no grid-wide synchronisation: state (p, q', K.q, K.q') in shared memory, warp w
// owns trajectory w in the elementwise phases (observables are warp reductions), and the products K.q' and K[:,fixed].q'[fixed]
// are DMMA.8x8x4 contractions whose K operand streams from L2 in fragment order (k_build_kfrag: one 512-byte contiguous
// warp load per two k-steps, no shared-memory staging, no bank or tag conflicts).  Two CTA barriers per step, one launch per run.
constexpr int EN_T = 8;
constexpr int EN_D = 4;     // K fragment pairs in flight per tile
struct EnsArgs {
    BathSet bs;
    int nph, ld, lds, ntraj, nmd, has_cons, nk8, nkc8, ncpmax, lin_bath;
    long long t0, nsteps;
    double dt;
    const double *kfrag;            // [ntile][nk8 + nkc8][32][2]: all dofs, then the constrained dofs again (nk8, nkc8 multiples of EN_D)
    const int *cidx8;               // [nkc8 * 8] constrained dof of contraction slot, -1 = padding
    double *q, *p, *G;
    const unsigned char *cons;
    double *etot;
};

// fragment-ordered copy of K: out[((j * nkt + s) * 32 + l) * 2 + h] = K[8 j + l / 4][col(8 s + 4 h + l % 4)], zero outside;
// col(slot) = slot for s < nk8 (all dofs), cols[slot - 8 nk8] after that (the constrained dofs, -1 = padding)
__global__ void k_build_kfrag(const double *__restrict__ K, int nph, int ld, int nk8, int nkc8, const int *__restrict__ cols,
                              double *__restrict__ out) {
    const int j = blockIdx.x, nkt = nk8 + nkc8;
    for (int e = threadIdx.x; e < nkt * 64; e += blockDim.x) {
        const int s = e >> 6, l = (e >> 1) & 31, h = e & 1;
        const int row = 8 * j + (l >> 2), slot = 8 * s + 4 * h + (l & 3);
        const int col = s < nk8 ? (slot < nph ? slot : -1) : cols[slot - 8 * nk8];
        out[((size_t)j * nkt) * 64 + e] = (row < nph && col >= 0) ? K[(size_t)row * ld + col] : 0.0;
    }
}

constexpr int EN_LC = 64;   // largest dense (has_lin) bath the ensemble kernel takes: two outputs per lane
// EN_W warps per CTA: the first EN_T own a trajectory each in the elementwise phases, all of them share the DMMA product (sixteen
// warps = four per scheduler hide the latency of the DMMA / operand chains that eight leave exposed; 128 registers per thread then)
template <int NBATH, int NTILE, int EN_G, bool CONS, bool LIN, int EN_W>
__global__ void __launch_bounds__(EN_W * 32, 1) k_md_ens(const EnsArgs a) {
    extern __shared__ __align__(16) double esm[];
    const int nph = a.nph, lds = a.lds;
    double *sp = esm, *sq = sp + EN_T * lds, *sg = sq + EN_T * lds, *sg1 = sg + EN_T * lds;
    double *snz = sg1 + EN_T * lds;                       // [nb][EN_T][ncpmax]: the noise rows of the current slab
    // dense bath (at most one): its matrix part W.[x | q] for evaluations A/B (ml0) and C (ml1), and p^1 on its dofs (mx1)
    double *slin = snz + (size_t)a.bs.nb * EN_T * a.ncpmax;
    const int tid = threadIdx.x, lane = tid & 31, wg = tid >> 5;       // wg: warp in the product
    // warps per trajectory in the elementwise phases: with sixteen warps and no dense bath two warps share a trajectory (elements
    // lane + 32 hf, stride 64; the observables of the two halves are combined through shared memory after the barrier)
    constexpr int NH = (!LIN && EN_W >= 2 * EN_T) ? 2 : 1;
    const bool own = wg < NH * EN_T;                                     // this warp works on trajectory w in the elementwise phases
    const int w = own ? wg % EN_T : 0, hf = own ? wg / EN_T : 0;
    const int e0 = lane + 32 * hf;
    constexpr int ES = 32 * NH;
    __shared__ double sred[2 * EN_T * (1 + NBATH)];
    // per-dof bath tables of the diagonal baths (position in the bath, c0 k0[c]) in shared memory: the elementwise phases read them three
    // times per step, and through L1 the position -> coefficient chain was two dependent global loads per element
    double *skx = slin + (LIN ? 3 * EN_T * EN_LC : 0);
    int *sinv = reinterpret_cast<int *>(skx + NBATH * lds);
    constexpr int ED = EN_W == 8 ? EN_D : 2;                             // fragment pairs in flight per tile
    double *ml0 = slin + w * EN_LC, *ml1 = slin + (EN_T + w) * EN_LC, *mx1 = slin + (2 * EN_T + w) * EN_LC;
    const int bl = a.lin_bath;                            // index of the dense bath or -1
    const int gtraj = blockIdx.x * EN_T + w, ltraj = min(gtraj, a.ntraj - 1);
    const bool live = gtraj < a.ntraj;
    const double dt = a.dt;
    double *mp = sp + w * lds, *mq = sq + w * lds, *mg = sg + w * lds, *mg1 = sg1 + w * lds;
    if constexpr (!LIN) {
        for (int e = tid; e < NBATH * lds; e += EN_W * 32) {
            const int b = e / lds, i = e % lds;
            const int c = (b < a.bs.nb && i < nph) ? a.bs.b[b].inv[i] : -1;
            sinv[e] = c;
            skx[e] = c >= 0 ? a.bs.b[b].c0 * a.bs.b[b].k0[c] : 0.0;
        }
    }
    for (int i = e0; own && i < lds; i += ES) {
        const bool ok = i < nph;
        mp[i] = ok ? a.p[(size_t)ltraj * a.ld + i] : 0.0;
        mq[i] = ok ? a.q[(size_t)ltraj * a.ld + i] : 0.0;
        mg[i] = ok ? a.G[(size_t)ltraj * a.ld + i] : 0.0;
        mg1[i] = 0.0;
    }
    auto load_noise = [&](int slab) {      // asynchronous: the rows of this warp's trajectory, 16 bytes at a time
#pragma unroll
        for (int b = 0; b < NBATH; ++b)
            if (b < a.bs.nb) {
                const double *src = a.bs.b[b].noise + ((size_t)ltraj * a.nmd + slab) * a.bs.b[b].ncp;
                double *dst = snz + ((size_t)b * EN_T + w) * a.ncpmax;
                for (int c = 2 * e0; c < a.bs.b[b].ncp; c += 2 * ES) cp_async16_zfill(dst + c, src + c, 16);
            }
        cp_async_commit();
    };
    // out[c] = sum_k W[c][k] [x | q][k] of the dense bath for this warp's trajectory (x: the full-length array, or the
    // bath-ordered copy mx1 when `compact`); lane l owns outputs l and l + 32, W^T rows are read lane-contiguous
    auto lin_eval = [&](double *out, const double *xsrc, bool compact) {
        if (!LIN || bl < 0) return;
        const BathDev &d = a.bs.b[bl];
        double o0 = 0.0, o1 = 0.0;
        for (int k = 0; k < d.Kw; ++k) {
            const int kk = k < d.ncp ? k : k - d.ncp;
            double xv = 0.0;
            if (kk < d.nc) {
                const int dof = d.cids[kk];
                xv = k < d.ncp ? (compact ? xsrc[kk] : xsrc[dof]) : mq[dof];
            }
            const double *wr = d.wt + (size_t)k * d.ncp;
            if (lane < d.nc) o0 = fma(wr[lane], xv, o0);
            if (lane + 32 < d.nc) o1 = fma(wr[lane + 32], xv, o1);
        }
        __syncwarp();
        if (lane < d.nc) out[lane] = o0;
        if (lane + 32 < d.nc) out[lane + 32] = o1;
        __syncwarp();
    };
    auto bforce = [&](int b, int c, double x, const double *lin) -> double {      // bath_force() of the launch chain, ml = 1
        double fb = snz[((size_t)b * EN_T + w) * a.ncpmax + c];
        if (!LIN || a.bs.b[b].diag) fb -= a.bs.b[b].c0 * a.bs.b[b].k0[c] * x;
        if (LIN && b == bl) fb += lin[c];
        return fb;
    };
    if (own) load_noise((int)(a.t0 % a.nmd));
    cp_async_wait<0>();
    __syncthreads();                                      // tables, state and noise rows of every warp are in place
    const int arow = lane >> 2, aslot = lane & 3;
    for (long long s = 0; s < a.nsteps; ++s) {
        const long long t = a.t0 + s;
        const int slab = (int)(t % a.nmd);
        // ---- evaluation A (md.py:383-398) of this warp's trajectory
        if (own) {
            double ke = 0.0, cur[NBATH];
#pragma unroll
            for (int b = 0; b < NBATH; ++b) cur[b] = 0.0;
            lin_eval(ml0, mp, false);                            // from (p_t, q_t), before they are overwritten below
#pragma unroll 4
            for (int i = e0; i < nph; i += ES) {        // unrolled: the table loads of four elements overlap
                const double pi = mp[i];
                double f = -mg[i];
#pragma unroll
                for (int b = 0; b < NBATH; ++b)
                    if (b < a.bs.nb) {
                        const int c = LIN ? a.bs.b[b].inv[i] : sinv[b * lds + i];
                        if (c >= 0) {
                            const double fb = LIN ? bforce(b, c, pi, ml0) : snz[((size_t)b * EN_T + w) * a.ncpmax + c] - skx[b * lds + i] * pi;
                            cur[b] += fb * pi;
                            f += fb;
                            if (live) a.bs.b[b].ring[((size_t)gtraj * a.bs.b[b].ml + (int)(t % a.bs.b[b].ml)) * a.bs.b[b].ncp + c] = pi;
                        }
                    }
                ke += 0.5 * pi * pi;
                mp[i] = pi + f * dt / 2.0;                       // p_half
                mq[i] = mq[i] + pi * dt + f * dt * dt / 2.0;     // q'
            }
            ke = warp_sum(ke);
            if (NH == 1) {
                if (live && lane == 0) a.etot[(size_t)slab * a.ntraj + gtraj] = ke;
            } else if (lane == 0) {
                sred[(hf * EN_T + w) * (1 + NBATH)] = ke;
            }
#pragma unroll
            for (int b = 0; b < NBATH; ++b)
                if (b < a.bs.nb) {
                    const double c = warp_sum(cur[b]);
                    if (NH == 1) {
                        if (live && lane == 0) a.bs.b[b].cur[(size_t)slab * a.ntraj + gtraj] = c;
                    } else if (lane == 0) {
                        sred[(hf * EN_T + w) * (1 + NBATH) + 1 + b] = c;
                    }
                }
        }
        __syncthreads();                                  // q' of all eight trajectories is in place; every warp is done with noise slab t
        if (own) load_noise((int)((t + 1) % a.nmd));      // evaluations B, C and the next evaluation A read slab t+1 (lands under the product)
        if (NH == 2 && own && hf == 0 && live && lane <= NBATH) {      // observables: first half + second half
            const double v = sred[w * (1 + NBATH) + lane] + sred[(EN_T + w) * (1 + NBATH) + lane];
            if (lane == 0) a.etot[(size_t)slab * a.ntraj + gtraj] = v;
            else if (lane - 1 < a.bs.nb) a.bs.b[lane - 1].cur[(size_t)slab * a.ntraj + gtraj] = v;
        }
        // ---- K.q' (and the part of it that comes from the constrained dofs): DMMA over fragment-ordered K
        {
            // (the fragment loads bypass L1 -- ld.global.nc.L1::no_allocate -- so that the 3.5 MB streamed per step do not evict the
            // bath index / friction / constraint tables the elementwise phases read through L1)
            // The K fragments of a tile are ONE stream of nk8 + nkc8 double2 per lane (all dofs, then the constrained ones again);
            // EN_D of them are in flight per tile (register ring, refilled right after use), for EN_G (4 or 5) tiles at a time:
            // the L2 round trip (~1000 cycles under load) is covered by 4 x 5 x 512 bytes per warp.
            const double *arowp = sq + arow * lds;
            const int nkt = a.nk8 + a.nkc8;
            const size_t tstride = (size_t)EN_W * nkt * 32;     // double2 elements between the tiles wg + EN_W j of this warp
#pragma unroll 1
            for (int g0 = 0; g0 < NTILE; g0 += EN_G) {
                // two accumulators per tile (even / odd half of a k-block): 2 EN_G independent DMMA chains per warp -- with two
                // warps per scheduler the tensor pipe needs that many to stay busy; the A operands are fetched one block ahead
                double acc[2][EN_G][2], accc[2][EN_G][2];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < EN_G; ++j) acc[h][j][0] = acc[h][j][1] = accc[h][j][0] = accc[h][j][1] = 0.0;
                const double2 *kb = reinterpret_cast<const double2 *>(a.kfrag) + ((size_t)(wg + EN_W * g0) * nkt) * 32 + lane;
                double2 bn[ED][EN_G];
#pragma unroll
                for (int d = 0; d < ED; ++d)
#pragma unroll
                    for (int j = 0; j < EN_G; ++j) bn[d][j] = ld_stream2(reinterpret_cast<const double *>(kb + j * tstride + (size_t)d * 32));
                double a0 = arowp[aslot], a1 = arowp[4 + aslot];
                for (int ks = 0; ks < a.nk8; ks += ED) {        // nk8, nkc8 are multiples of EN_D (and of ED)
#pragma unroll
                    for (int d = 0; d < ED; ++d) {
                        const int nx = min(ks + d + 1, a.nk8 - 1);
                        const double n0 = arowp[8 * nx + aslot], n1 = arowp[8 * nx + 4 + aslot];
#pragma unroll
                        for (int j = 0; j < EN_G; ++j) dmma884(acc[0][j][0], acc[0][j][1], a0, bn[d][j].x);
#pragma unroll
                        for (int j = 0; j < EN_G; ++j) dmma884(acc[1][j][0], acc[1][j][1], a1, bn[d][j].y);
                        if (ks + d + ED < nkt) {
#pragma unroll
                            for (int j = 0; j < EN_G; ++j) bn[d][j] = ld_stream2(reinterpret_cast<const double *>(kb + j * tstride + (size_t)(ks + d + ED) * 32));
                        }
                        a0 = n0;
                        a1 = n1;
                    }
                }
                if (CONS) {
                    for (int ks = 0; ks < a.nkc8; ks += ED) {
#pragma unroll
                        for (int d = 0; d < ED; ++d) {
                            const int c0 = a.cidx8[8 * (ks + d) + aslot], c1 = a.cidx8[8 * (ks + d) + 4 + aslot];
                            const double x0 = c0 >= 0 ? arowp[c0] : 0.0, x1 = c1 >= 0 ? arowp[c1] : 0.0;
#pragma unroll
                            for (int j = 0; j < EN_G; ++j) dmma884(accc[0][j][0], accc[0][j][1], x0, bn[d][j].x);
#pragma unroll
                            for (int j = 0; j < EN_G; ++j) dmma884(accc[1][j][0], accc[1][j][1], x1, bn[d][j].y);
                            if (a.nk8 + ks + d + ED < nkt) {
#pragma unroll
                                for (int j = 0; j < EN_G; ++j) bn[d][j] = ld_stream2(reinterpret_cast<const double *>(kb + j * tstride + (size_t)(a.nk8 + ks + d + ED) * 32));
                            }
                        }
                    }
                }
                // D[traj = lane / 4][n = 8 tile + 2 (lane % 4) + {0, 1}]; sg (K.q of the step just evaluated) is free again
#pragma unroll
                for (int j = 0; j < EN_G; ++j) {
                    const int n = 8 * (wg + EN_W * (g0 + j)) + 2 * aslot;
                    const double v0 = acc[0][j][0] + acc[1][j][0], v1 = acc[0][j][1] + acc[1][j][1];
                    const double c0 = accc[0][j][0] + accc[1][j][0], c1 = accc[0][j][1] + accc[1][j][1];
                    if (n < nph) {
                        sg1[arow * lds + n] = v0;
                        sg[arow * lds + n] = v0 - c0;
                    }
                    if (n + 1 < nph) {
                        sg1[arow * lds + n + 1] = v1;
                        sg[arow * lds + n + 1] = v1 - c1;
                    }
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        // ---- evaluations B and C (md.py:401-404), constraint (md.py:407-408)
        if (!own) {
        } else if (LIN && bl >= 0) {
            // a dense bath couples its dofs, so B and C cannot be chained per element: B gives p^1, kept only on the dense
            // bath's dofs (mx1); C recomputes p^1 of its own element and adds the matrix part taken at p^1
            lin_eval(ml0, mp, false);                            // from (p_half, q')
            auto p_one = [&](int i, double ph, double g1) {
                double f = -g1;
#pragma unroll
                for (int b = 0; b < NBATH; ++b)
                    if (b < a.bs.nb) {
                        const int c = a.bs.b[b].inv[i];
                        if (c >= 0) f += bforce(b, c, ph, ml0);
                    }
                return ph + dt * f / 2.0;
            };
            for (int i = lane; i < nph; i += 32) {
                const int c = a.bs.b[bl].inv[i];
                if (c >= 0) mx1[c] = p_one(i, mp[i], mg1[i]);
            }
            __syncwarp();
            lin_eval(ml1, mx1, true);                            // from (p^1, q')
            for (int i = lane; i < nph; i += 32) {
                const double ph = mp[i], g1 = mg1[i];
                const double p1 = p_one(i, ph, g1);
                double f = -g1;
#pragma unroll
                for (int b = 0; b < NBATH; ++b)
                    if (b < a.bs.nb) {
                        const int c = a.bs.b[b].inv[i];
                        if (c >= 0) f += bforce(b, c, p1, ml1);
                    }
                const double pnew = ph + dt * f / 2.0;
                const bool fx = CONS && a.cons[i];
                mp[i] = fx ? 0.0 : pnew;
                if (fx) mq[i] = 0.0;
            }
        } else {
#pragma unroll 4
            for (int i = e0; i < nph; i += ES) {
                const double ph = mp[i], g1 = mg1[i];
                double nz[NBATH], kx[NBATH];
                bool in[NBATH];
#pragma unroll
                for (int b = 0; b < NBATH; ++b) {
                    const int c = sinv[b * lds + i];
                    in[b] = c >= 0;
                    nz[b] = in[b] ? snz[((size_t)b * EN_T + w) * a.ncpmax + c] : 0.0;
                    kx[b] = skx[b * lds + i];
                }
                double xi = ph, pnew = 0.0;
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
                    double f = -g1;
#pragma unroll
                    for (int b = 0; b < NBATH; ++b)
                        if (in[b]) f += nz[b] - kx[b] * xi;
                    pnew = ph + dt * f / 2.0;
                    xi = pnew;
                }
                const bool fx = CONS && a.cons[i];
                mp[i] = fx ? 0.0 : pnew;
                if (fx) mq[i] = 0.0;
            }
        }
        __syncwarp();
    }
    if (own && live)
        for (int i = e0; i < nph; i += ES) {
            a.p[(size_t)gtraj * a.ld + i] = mp[i];
            a.q[(size_t)gtraj * a.ld + i] = mq[i];
            a.G[(size_t)gtraj * a.ld + i] = mg[i];
        }
}

// G[0] <- sum_z G[z] (- D): the persistent kernel works on one K.q vector per trajectory
__global__ void k_fold_g(double *G, int nsplit, size_t n, const double *D) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double v = G[e];
    for (int z = 1; z < nsplit; ++z) v += G[(size_t)z * n + e];
    if (D) v -= D[e];
    G[e] = v;
}

// device-resident step counter of the graph-captured part of a step
__global__ void k_set_t(long long *t, long long v) { *t = v; }
__global__ void k_advance_t(long long *t) { *t += 1; }

// xq[traj] = [ x[cids] | q[cids] ]  (operand of the time-local matrix product, baths.py:236-249)
__global__ void k_gather_xq(const int *__restrict__ cids, int nc, int ncp, int Kw, int ld,
                            const double *__restrict__ x, const double *__restrict__ q, double *__restrict__ xq) {
    const int traj = blockIdx.x;
    for (int a = threadIdx.x; a < nc; a += blockDim.x) {
        const int i = cids[a];
        xq[(size_t)traj * Kw + a] = x[(size_t)traj * ld + i];
        if (Kw > ncp) xq[(size_t)traj * Kw + ncp + a] = q[(size_t)traj * ld + i];
    }
}

// Diagonal-kernel friction tail (baths.py:453-457 restricted to j>=1), the HBM-streaming kernel:
//   out[z][traj][c] = dt * sum_{s in rows of split z} kern[((head - s) mod ml) + 1][c] * ring[traj][s][c]
// kern has ml+1 rows, row ml == 0, so the slot that would pair with j == ml drops out branch-free.
// Thread = one column pair (16-byte loads) x one row group; T trajectories share each kernel load.
template <int T>
__global__ void __launch_bounds__(512) k_tail_diag(const double *__restrict__ ring, const double *__restrict__ kern,
                                                    double *__restrict__ out, int ntraj, int ml, int ncp, int head,
                                                    int rows_per_split, int cpt, int rg_count, double dt) {
    extern __shared__ double red[];  // [rg_count][T][2*cpt]
    const int cp = threadIdx.x % cpt, rg = threadIdx.x / cpt;
    const int col = (blockIdx.z * cpt + cp) * 2;
    const int traj0 = blockIdx.x * T;
    const int r0 = blockIdx.y * rows_per_split, r1 = min(ml, r0 + rows_per_split);
    const bool active = rg < rg_count && col < ncp;
    double2 acc[T];
#pragma unroll
    for (int k = 0; k < T; ++k) acc[k] = make_double2(0.0, 0.0);
    if (active) {
        const size_t tstride = (size_t)ml * ncp;
        const double *rbase = ring + (size_t)traj0 * tstride + col;
        constexpr int U = 4;
        int s = r0 + rg;
        for (; s + (U - 1) * rg_count < r1; s += U * rg_count) {
            double2 kv[U], hv[U][T];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int ss = s + u * rg_count;
                int d = head - ss;
                if (d < 0) d += ml;
                kv[u] = *reinterpret_cast<const double2 *>(kern + (size_t)(d + 1) * ncp + col);
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    const int tr = min(traj0 + k, ntraj - 1) - traj0;
                    hv[u][k] = ld_stream2(rbase + (size_t)tr * tstride + (size_t)ss * ncp);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    acc[k].x = fma(kv[u].x, hv[u][k].x, acc[k].x);
                    acc[k].y = fma(kv[u].y, hv[u][k].y, acc[k].y);
                }
        }
        for (; s < r1; s += rg_count) {
            int d = head - s;
            if (d < 0) d += ml;
            const double2 kv = *reinterpret_cast<const double2 *>(kern + (size_t)(d + 1) * ncp + col);
#pragma unroll
            for (int k = 0; k < T; ++k) {
                const int tr = min(traj0 + k, ntraj - 1) - traj0;
                const double2 hv = ld_stream2(rbase + (size_t)tr * tstride + (size_t)s * ncp);
                acc[k].x = fma(kv.x, hv.x, acc[k].x);
                acc[k].y = fma(kv.y, hv.y, acc[k].y);
            }
        }
    }
    if (rg < rg_count) {
#pragma unroll
        for (int k = 0; k < T; ++k) {
            red[((size_t)rg * T + k) * 2 * cpt + 2 * cp] = acc[k].x;
            red[((size_t)rg * T + k) * 2 * cpt + 2 * cp + 1] = acc[k].y;
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < T * 2 * cpt; e += blockDim.x) {
        const int k = e / (2 * cpt), cc = e % (2 * cpt);
        const int c = blockIdx.z * cpt * 2 + cc;
        if (traj0 + k >= ntraj || c >= ncp) continue;
        double s = 0.0;
        for (int g = 0; g < rg_count; ++g) s += red[((size_t)g * T + k) * 2 * cpt + cc];
        out[((size_t)blockIdx.y * ntraj + traj0 + k) * ncp + c] = dt * s;
    }
}

// ---- time-blocked history tails (diagonal kernels, long memory) -------------------------------------
// The friction tail of step t = t0 + s (t0 = block start, s < TB)
//   S'(t) = dt sum_{j=1}^{ml-1} k[j] p_{t+1-j}  =  Near_s + Far_s
//   Near_s = dt sum_{j=1}^{s+1}  k[j] p_{t0+s+1-j}            (p_{t0} .. p_{t0+s}: at most TB fresh ring rows)
//   Far_s  = dt sum_{d>=0}       k[s+2+d] p_{t0-1-d}          (everything older than the block start)
// All TB far tails of a block are produced by ONE pass over the ring (each loaded p feeds TB accumulators), so
// the ring is streamed from HBM once per TB steps instead of once per step; the flops are unchanged.
// kern is zero-padded past row ml-1, which also retires the ring slots that are younger than the block start.
constexpr int TB = 16;

template <int T>
__global__ void __launch_bounds__(256, 1) k_tail_far(const double *__restrict__ ring, const double *__restrict__ kern,
                                                      double *__restrict__ out, int ntraj, int ml, int ncp, int base,
                                                      int ages_per_split, int ct, double dt) {
    const int c = blockIdx.z * ct + threadIdx.x;
    if ((int)threadIdx.x >= ct || c >= ncp) return;
    const int traj0 = blockIdx.x * T;
    const int d_lo = blockIdx.y * ages_per_split, d_hi = min(d_lo + ages_per_split, ml);
    double acc[T][TB];
#pragma unroll
    for (int k = 0; k < T; ++k)
#pragma unroll
        for (int s2 = 0; s2 < TB; ++s2) acc[k][s2] = 0.0;
    const size_t tstride = (size_t)ml * ncp;
    const double *rb[T];
#pragma unroll
    for (int k = 0; k < T; ++k) rb[k] = ring + (size_t)min(traj0 + k, ntraj - 1) * tstride + c;
    for (int d0 = d_lo; d0 < d_hi; d0 += TB) {
        double kw[2 * TB - 1];
#pragma unroll
        for (int i = 0; i < 2 * TB - 1; ++i) kw[i] = kern[(size_t)(d0 + 2 + i) * ncp + c];
        int slot = (base - d0) % ml;
        if (slot < 0) slot += ml;
#pragma unroll
        for (int u = 0; u < TB; ++u) {
            double pv[T];
#pragma unroll
            for (int k = 0; k < T; ++k) {
                double v;
                asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(rb[k] + (size_t)slot * ncp));
                pv[k] = v;
            }
#pragma unroll
            for (int s2 = 0; s2 < TB; ++s2)
#pragma unroll
                for (int k = 0; k < T; ++k) acc[k][s2] = fma(kw[u + s2], pv[k], acc[k][s2]);
            slot = slot == 0 ? ml - 1 : slot - 1;
        }
    }
#pragma unroll
    for (int s2 = 0; s2 < TB; ++s2)
#pragma unroll
        for (int k = 0; k < T; ++k)
            if (traj0 + k < ntraj) out[(((size_t)blockIdx.y * TB + s2) * ntraj + traj0 + k) * ncp + c] = dt * acc[k][s2];
}

// TMA-staged variant (ncp <= 320): one elected thread streams contiguous 16-row ring chunks of T trajectories into
// shared memory with cp.async.bulk (SASS UBLKCP) against an mbarrier, two stages deep, while all warps run the
// 16x16 Toeplitz update of the previous chunk out of shared memory.  Age u of a chunk sits in smem row 15-u.
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity));
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

template <int T, int STAGES>
__global__ void __launch_bounds__(320, 1) k_tail_far_tma(const double *__restrict__ ring, const double *__restrict__ kern,
                                                          double *__restrict__ out, int ntraj, int ml, int ncp, int base,
                                                          int ages_per_split, double dt) {
    extern __shared__ __align__(128) double stage_mem[];   // [STAGES][T][TB][ncp]
    __shared__ __align__(8) uint64_t full[STAGES];
    const int c = threadIdx.x;
    const int traj0 = blockIdx.x * T;
    const int d_lo = blockIdx.y * ages_per_split, d_hi = min(d_lo + ages_per_split, ml);
    const int nchunk = (d_hi - d_lo + TB - 1) / TB;
    const size_t tstride = (size_t)ml * ncp, stage_elems = (size_t)T * TB * ncp;
    const unsigned rowbytes = (unsigned)ncp * 8u;

    auto issue = [&](int ch) {   // one thread: ring rows of chunk ch -> stage ch % STAGES
        double *dst = stage_mem + (size_t)(ch % STAGES) * stage_elems;
        uint64_t *bar = &full[ch % STAGES];
        mbar_expect_tx(bar, (unsigned)(T * TB) * rowbytes);
        int lo = (base - (d_lo + ch * TB) - (TB - 1)) % ml;
        if (lo < 0) lo += ml;
        const int n1 = min(TB, ml - lo);           // rows lo .. lo+n1-1, then (wrap) rows 0 .. TB-n1-1
#pragma unroll
        for (int k = 0; k < T; ++k) {
            const double *src = ring + (size_t)min(traj0 + k, ntraj - 1) * tstride;
            bulk_g2s(dst + (size_t)k * TB * ncp, src + (size_t)lo * ncp, (unsigned)n1 * rowbytes, bar);
            if (n1 < TB) bulk_g2s(dst + ((size_t)k * TB + n1) * ncp, src, (unsigned)(TB - n1) * rowbytes, bar);
        }
    };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i)
            if (i < nchunk) issue(i);
    }
    const bool active = c < ncp;
    const int cc = active ? c : 0;
    double acc[T][TB];
#pragma unroll
    for (int k = 0; k < T; ++k)
#pragma unroll
        for (int s2 = 0; s2 < TB; ++s2) acc[k][s2] = 0.0;
    double kw[2 * TB - 1];
#pragma unroll
    for (int i = 0; i < 2 * TB - 1; ++i) kw[i] = kern[(size_t)(d_lo + 2 + i) * ncp + cc];
    for (int ch = 0; ch < nchunk; ++ch) {
        const int d0 = d_lo + ch * TB;
        double knew[TB];   // the 16 kernel rows the next chunk adds to the sliding window (prefetched under the math)
#pragma unroll
        for (int i = 0; i < TB; ++i) knew[i] = kern[(size_t)(d0 + 2 + 2 * TB - 1 + i) * ncp + cc];
        mbar_wait(&full[ch % STAGES], (unsigned)((ch / STAGES) & 1));
        const double *sp = stage_mem + (size_t)(ch % STAGES) * stage_elems + cc;
#pragma unroll
        for (int u = 0; u < TB; ++u) {
            double pv[T];
#pragma unroll
            for (int k = 0; k < T; ++k) pv[k] = sp[((size_t)k * TB + (TB - 1 - u)) * ncp];
#pragma unroll
            for (int s2 = 0; s2 < TB; ++s2)
#pragma unroll
                for (int k = 0; k < T; ++k) acc[k][s2] = fma(kw[u + s2], pv[k], acc[k][s2]);
        }
#pragma unroll
        for (int i = 0; i < TB - 1; ++i) kw[i] = kw[i + TB];
#pragma unroll
        for (int i = 0; i < TB; ++i) kw[TB - 1 + i] = knew[i];
        __syncthreads();                            // everyone is done with this stage
        if (threadIdx.x == 0 && ch + STAGES < nchunk) issue(ch + STAGES);
    }
    if (active) {
#pragma unroll
        for (int s2 = 0; s2 < TB; ++s2)
#pragma unroll
            for (int k = 0; k < T; ++k)
                if (traj0 + k < ntraj) out[(((size_t)blockIdx.y * TB + s2) * ntraj + traj0 + k) * ncp + c] = dt * acc[k][s2];
    }
}

// Warp-specialised variant: one producer warp keeps NST small stages (SR ring rows of T trajectories each) in flight with
// cp.async.bulk against full/empty mbarrier pairs while the consumer warps run the Toeplitz update; a stage is handed back
// as soon as its SR rows are consumed, so (NST-1)/NST of the staging memory is always in flight (the two-stage kernel above
// has at most half in flight and is bound by the bulk-copy latency: 3.2 us per 77 KB stage against 1.3 us of math).
// Measured at C5 (ms per ring pass): 2 x 16 rows 2.93 | 16 x 2 rows 2.92 | 8 x 4 rows 2.42-2.62 | 4 x 8 rows 2.39 | 5 x 8 rows 2.19;
// one trajectory per CTA (10 x 8 or 5 x 16 rows) 2.7: the kernel-window loads are no longer shared by two trajectories.
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <int T, int SR, int NST>
__global__ void __launch_bounds__(352, 1) k_tail_far_ws(const double *__restrict__ ring, const double *__restrict__ kern,
                                                         double *__restrict__ out, int ntraj, int ml, int ncp, int base,
                                                         int ages_per_split, double dt) {
    static_assert(TB % SR == 0, "stage rows must divide the time block");
    extern __shared__ __align__(128) double stage_mem[];   // [NST][T][SR][ncp]
    __shared__ __align__(8) uint64_t full[NST], empty[NST];
    const int ncons = blockDim.x - 32;                      // consumer threads (whole warps), then one producer warp
    const int traj0 = blockIdx.x * T;
    const int d_lo = blockIdx.y * ages_per_split, d_hi = min(d_lo + ages_per_split, ml);
    const int nchunk = (d_hi - d_lo + TB - 1) / TB, nsub = nchunk * (TB / SR);
    const size_t tstride = (size_t)ml * ncp, stage_elems = (size_t)T * SR * ncp;
    const unsigned rowbytes = (unsigned)ncp * 8u;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NST; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], (unsigned)(ncons / 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if ((int)threadIdx.x >= ncons) {                        // ---- producer warp (one elected lane)
        if ((int)threadIdx.x == ncons) {
            for (int sc = 0; sc < nsub; ++sc) {
                const int st = sc % NST;
                if (sc >= NST) mbar_wait(&empty[st], (unsigned)(((sc / NST) - 1) & 1));
                double *dst = stage_mem + (size_t)st * stage_elems;
                mbar_expect_tx(&full[st], (unsigned)(T * SR) * rowbytes);
                int lo = (base - (d_lo + sc * SR) - (SR - 1)) % ml;
                if (lo < 0) lo += ml;
                const int n1 = min(SR, ml - lo);            // rows lo .. lo+n1-1, then (wrap) rows 0 .. SR-n1-1
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    const double *src = ring + (size_t)min(traj0 + k, ntraj - 1) * tstride;
                    bulk_g2s(dst + (size_t)k * SR * ncp, src + (size_t)lo * ncp, (unsigned)n1 * rowbytes, &full[st]);
                    if (n1 < SR) bulk_g2s(dst + ((size_t)k * SR + n1) * ncp, src, (unsigned)(SR - n1) * rowbytes, &full[st]);
                }
            }
        }
        return;
    }
    const int c = threadIdx.x;                              // ---- consumers: thread per column
    const bool active = c < ncp;
    const int cc = active ? c : 0;
    double acc[T][TB];
#pragma unroll
    for (int k = 0; k < T; ++k)
#pragma unroll
        for (int s2 = 0; s2 < TB; ++s2) acc[k][s2] = 0.0;
    // kernel window as a 32-entry circular register file: in a chunk of phase PH the logical entry kw[i] (i < 31) lives in
    // W[(16 PH + i) & 31].  The 16 entries the next chunk adds are loaded straight into the slots that die during this chunk
    // (4 per consumed stage, one stage later than their last use), so the window never shifts and needs no staging copy.
    double W[32];
#pragma unroll
    for (int i = 0; i < 2 * TB - 1; ++i) W[i] = kern[(size_t)(d_lo + 2 + i) * ncp + cc];
    W[31] = 0.0;
    auto chunk = [&](auto ph, int ch) {
        constexpr int O = decltype(ph)::value * TB;
        const double *knext = kern + (size_t)(d_lo + (ch + 1) * TB + 2) * ncp + cc;     // logical entry i of the next chunk: knext[i * ncp]
#pragma unroll
        for (int q = 0; q < TB / SR; ++q) {
            const int sc = ch * (TB / SR) + q, st = sc % NST;
            mbar_wait(&full[st], (unsigned)((sc / NST) & 1));
            const double *sp = stage_mem + (size_t)st * stage_elems + cc;
#pragma unroll
            for (int u = 0; u < SR; ++u) {
                double pv[T];
#pragma unroll
                for (int k = 0; k < T; ++k) pv[k] = sp[((size_t)k * SR + (SR - 1 - u)) * ncp];
#pragma unroll
                for (int s2 = 0; s2 < TB; ++s2)
#pragma unroll
                    for (int k = 0; k < T; ++k) acc[k][s2] = fma(W[(O + q * SR + u + s2) & 31], pv[k], acc[k][s2]);
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[st]);     // this warp is done with the stage
            // slots O + q SR .. O + q SR + SR-1 are dead now: they receive the next chunk's logical entries 16 + q SR + j
#pragma unroll
            for (int j = 0; j < SR; ++j)
                if (q * SR + j + TB < 2 * TB - 1) W[(O + q * SR + j) & 31] = knext[(size_t)(TB + q * SR + j) * ncp];
            if (q == 0) W[(O + 31) & 31] = knext[(size_t)(TB - 1) * ncp];      // the slot this chunk never used
        }
    };
    for (int ch = 0; ch < nchunk; ch += 2) {
        chunk(std::integral_constant<int, 0>{}, ch);
        if (ch + 1 < nchunk) chunk(std::integral_constant<int, 1>{}, ch + 1);
    }
    if (active) {
#pragma unroll
        for (int s2 = 0; s2 < TB; ++s2)
#pragma unroll
            for (int k = 0; k < T; ++k)
                if (traj0 + k < ntraj) out[(((size_t)blockIdx.y * TB + s2) * ntraj + traj0 + k) * ncp + c] = dt * acc[k][s2];
    }
}

// Generalisation of k_tail_far_ws to a time block of TBK steps and a WS-entry circular kernel window (WS >= TBK + SR, WS a
// multiple of SR).  Sub-stage q consumes SR ring rows and needs the window entries [q SR, q SR + SR + TBK - 1).  Entry q SR + u is
// dead once row u has been processed; the entry the NEXT sub-stage needs at its row u is loaded right then into a slot that has
// just died, so at most TBK + SR entries are live and a load has a whole sub-stage (SR x TBK FMAs) to arrive.  All window indices
// are compile-time (the sub-stage body is instantiated for the WS / SR phases of the circle).
//   <T = 1, TBK = 32, SR = 4, NST = 20, WS = 36>: ONE ring pass per 32 steps (half the HBM traffic of the 16-step block) within the
//   168 registers that 352 threads allow (32 accumulators + 36 window entries per thread), one kernel-window load per 32 FMAs.
template <int I, int N, class F>
__device__ __forceinline__ void for_phases(F &f, int q0, int nsub) {      // f(phase I, sub-stage q0 + I) for the phases of one turn of the circle
    if constexpr (I < N) {
        if (q0 + I < nsub) {
            f(std::integral_constant<int, I>{}, q0 + I);
            for_phases<I + 1, N>(f, q0, nsub);
        }
    }
}
template <int T, int TBK, int SR, int NST, int WS>
__global__ void __launch_bounds__(352, 1) k_tail_far_wsx(const double *__restrict__ ring, const double *__restrict__ kern,
                                                          double *__restrict__ out, int ntraj, int ml, int ncp, int base,
                                                          int ages_per_split, double dt) {
    static_assert(WS % SR == 0 && WS >= TBK + SR, "window too small");
    constexpr int PH = WS / SR;
    extern __shared__ __align__(128) double stage_mem[];   // [NST][T][SR][ncp]
    __shared__ __align__(8) uint64_t full[NST], empty[NST];
    const int ncons = blockDim.x - 32;
    const int traj0 = blockIdx.x * T;
    const int d_lo = blockIdx.y * ages_per_split, d_hi = min(d_lo + ages_per_split, ml);
    const int nsub = (d_hi - d_lo + SR - 1) / SR;
    const size_t tstride = (size_t)ml * ncp, stage_elems = (size_t)T * SR * ncp;
    const unsigned rowbytes = (unsigned)ncp * 8u;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NST; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], (unsigned)(ncons / 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if ((int)threadIdx.x >= ncons) {                        // ---- producer warp (one elected lane)
        if ((int)threadIdx.x == ncons) {
            for (int sc = 0; sc < nsub; ++sc) {
                const int st = sc % NST;
                if (sc >= NST) mbar_wait(&empty[st], (unsigned)(((sc / NST) - 1) & 1));
                double *dst = stage_mem + (size_t)st * stage_elems;
                mbar_expect_tx(&full[st], (unsigned)(T * SR) * rowbytes);
                int lo = (base - (d_lo + sc * SR) - (SR - 1)) % ml;
                if (lo < 0) lo += ml;
                const int n1 = min(SR, ml - lo);
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    const double *src = ring + (size_t)min(traj0 + k, ntraj - 1) * tstride;
                    bulk_g2s(dst + (size_t)k * SR * ncp, src + (size_t)lo * ncp, (unsigned)n1 * rowbytes, &full[st]);
                    if (n1 < SR) bulk_g2s(dst + ((size_t)k * SR + n1) * ncp, src, (unsigned)(SR - n1) * rowbytes, &full[st]);
                }
            }
        }
        return;
    }
    const int c = threadIdx.x;
    const bool active = c < ncp;
    const int cc = active ? c : 0;
    double acc[T][TBK];
#pragma unroll
    for (int k = 0; k < T; ++k)
#pragma unroll
        for (int s2 = 0; s2 < TBK; ++s2) acc[k][s2] = 0.0;
    double W[WS];
    const double *kbase = kern + (size_t)(d_lo + 2) * ncp + cc;     // window entry e: kbase[e * ncp]
#pragma unroll
    for (int i = 0; i < WS; ++i) W[i] = i < TBK + SR - 1 ? kbase[(size_t)i * ncp] : 0.0;
    // running stage cursor (no divisions in the loop): stage index, its barrier parity, its rows, and the kernel rows the next
    // sub-stage adds; row offsets inside a stage are kept in registers
    int st = 0;
    unsigned par = 0;
    const double *sp = stage_mem + cc;
    const double *knew = kbase + (size_t)(SR + TBK - 1) * ncp;
    int roff[SR];
#pragma unroll
    for (int u = 0; u < SR; ++u) roff[u] = u * ncp;
    auto sub = [&](auto ph, int) {
        constexpr int O = decltype(ph)::value * SR;
        mbar_wait(&full[st], par);
#pragma unroll
        for (int u = 0; u < SR; ++u) {
            double pv[T];
#pragma unroll
            for (int k = 0; k < T; ++k) pv[k] = sp[k * SR * ncp + roff[SR - 1 - u]];
#pragma unroll
            for (int s2 = 0; s2 < TBK; ++s2)
#pragma unroll
                for (int k = 0; k < T; ++k) acc[k][s2] = fma(W[(O + u + s2) % WS], pv[k], acc[k][s2]);
            W[(O + SR + TBK - 1 + u) % WS] = knew[roff[u]];       // slot of an entry that died at row u - 1 (or earlier)
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[st]);
        knew += SR * ncp;
        if (++st == NST) {
            st = 0;
            par ^= 1u;
            sp = stage_mem + cc;
        } else {
            sp += stage_elems;
        }
    };
    for (int q0 = 0; q0 < nsub; q0 += PH) for_phases<0, PH>(sub, q0, nsub);
    if (active) {
#pragma unroll
        for (int s2 = 0; s2 < TBK; ++s2)
#pragma unroll
            for (int k = 0; k < T; ++k)
                if (traj0 + k < ntraj) out[(((size_t)blockIdx.y * TBK + s2) * ntraj + traj0 + k) * ncp + c] = dt * acc[k][s2];
    }
}

// ---- far pass on the tensor pipe ---------------------------------------------------------------------------------------------
// For one dof c the TBK far tails of a block are a matrix product over the ages d:
//   Far[traj][s] = dt sum_d  P_c[traj][d] H_c[d][s],   P_c[traj][d] = p_{t0-1-d}[traj][c],   H_c[d][s] = k_c[s + 2 + d]   (Hankel)
// i.e. DMMA.8x8x4 tiles with M = 8 trajectories, K = 4 ages, N = 8 steps.  Against the DFMA kernels above an instruction does 256
// instead of 32 multiply-adds, the accumulators of a (dof, 8 trajectories, 32 steps) block are 16 registers instead of 256, and the
// Hankel operand needs ONE new fragment per k-step and dof: the fragment of (step tile nt, ages d0..d0+3) is F(d0 + 8 nt) with
// F(x)[lane] = k_c[x + 2 + lane/4 + lane%4], so the four tiles of a k-step take every second entry of a ring of fragments that advances
// by one entry per k-step.  That makes a 32-step block (half the ring traffic of the 16-step block) cheaper than the 16-step DFMA pass.
//   CTA = 8 trajectories x 32 dofs x an age range; warp w owns the dof pair (2w, 2w+1) of the chunk, both fed by one LDS.128.
//   A producer lane streams the ring through shared memory with cp.async.bulk.tensor (boxes of 16 dofs x 4 slots x 8 trajectories,
//   128-byte swizzle: the eight lanes of a quarter-warp -- trajectories 2i, 2i+1 x 4 ages -- hit eight different 16-byte chunks),
//   FM_NST stages of two k-steps each against full / empty mbarriers.  The kernel rows come from a transposed copy kT[c][j]
//   (fragment loads are 11 consecutive doubles, L1 hits, fetched ten k-steps ahead of their first use).
constexpr int FM_W = 16;                         // consumer warps (one dof pair each)
constexpr int FM_DC = 2 * FM_W;                  // dofs per CTA: two 16-dof boxes per ring row
constexpr int FM_SR = 8;                         // ages per stage (two k-steps)
constexpr int FM_NST = 8;                        // stages in flight
constexpr int FM_BOX = 8 * 4 * 16 * 8;           // one box: 8 trajectories x 4 slots x 16 dofs
constexpr int FM_STAGE = 4 * FM_BOX;             // [slot group][dof half]: 16 KB
constexpr int FM_RB = 10;                        // ring of Hankel fragments per dof: 4 k-steps x 2 live + 2 ahead
constexpr size_t FM_SMEM = (size_t)FM_NST * FM_STAGE + 1024 + 2 * FM_NST * 8 + 64;
// The pass is cut into SEGMENTS: stages [a_lo, a_hi) (8 ages each, first age d0 + 8 a_lo) of one (32-dof chunk, 8-trajectory group) pair,
// whose 32 partial tails go to partial slot `slot` of that pair.  A CTA works through FM_SPC segments; the host plan (FarPlan) deals the
// segments of a whole pass out to 32 slices of about one CTA per SM each, so that one slice can be launched per MD step.
struct FarMmaArgs {
    const double *kT;      // [ncp][ldk], zero beyond row ml - 1
    double *out;           // [slot][TBK][ntraj][ncp]
    const FarSeg *segs;    // [gridDim.x][FM_SPC] of this launch
    int ldk, ntraj, ml, ncp, base, d0;
    double dt;
};
// kT[c][j] = kern[j][c] for j < ml, 0 up to ldk
__global__ void k_kernel_transpose(const double *__restrict__ kern, int ml, int ncp, int ldk, double *__restrict__ kT) {
    __shared__ double tile[32][33];
    const int j0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = j0 + r, c = c0 + threadIdx.x;
        tile[r][threadIdx.x] = (j < ml && c < ncp) ? kern[(size_t)j * ncp + c] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, j = j0 + threadIdx.x;
        if (c < ncp && j < ldk) kT[(size_t)c * ldk + j] = tile[threadIdx.x][r];
    }
}
template <int I, int N, class F>
__device__ __forceinline__ void fm_unroll(F &f, int q0, int nk) {
    if constexpr (I < N) {
        if (q0 + I < nk) {
            f(std::integral_constant<int, I>{}, q0 + I);
            fm_unroll<I + 1, N>(f, q0, nk);
        }
    }
}
template <int TBK>
__device__ __forceinline__ void far_mma_body(const CUtensorMap *ringmap, const FarMmaArgs &a, int cta) {
    constexpr int NT = TBK / 8;
    static_assert(NT >= 1 && 2 * (NT - 1) + 1 <= FM_RB - 2, "fragment ring too short for this block length");
    extern __shared__ unsigned char fm_raw[];
    unsigned char *sm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(fm_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(sm + (size_t)FM_NST * FM_STAGE), *empty = full + FM_NST;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const FarSeg *my = a.segs + (size_t)cta * FM_SPC;
    const int nlive = my[0].nlive;                           // warps with real dofs (the last chunk may be partly padding): only they consume
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < FM_NST; ++i) {
            tma_mbar_init(&full[i], 1);
            tma_mbar_init(&empty[i], nlive);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp < FM_W && warp >= nlive) return;                // (a spinning pad warp would only take issue slots: measured 3 % slower)
    if (warp == FM_W) {                                      // ---- producer (a warp of its own: merged into a consumer warp the refills
                                                             //      come late -- 3.44 instead of 2.83 ms -- even though 16 warps could have 128 registers)
        if (lane == 0) {
            int sc = 0;                                      // stages issued so far, over all segments of this CTA
            for (int g = 0; g < FM_SPC; ++g) {
                const FarSeg sg = my[g];
                for (int ai = sg.a_lo; ai < sg.a_hi; ++ai, ++sc) {
                    const int st = sc % FM_NST;
                    if (sc >= FM_NST) tma_mbar_wait(&empty[st], (unsigned)(((sc / FM_NST) - 1) & 1));
                    tma_mbar_expect(&full[st], (unsigned)FM_STAGE);
                    int lo = (a.base - (a.d0 + ai * FM_SR) - (FM_SR - 1)) % a.ml;      // slots lo .. lo+7 hold the ages d+7 .. d (no wrap: base+1, ml, d are multiples of 8)
                    if (lo < 0) lo += a.ml;
                    unsigned char *dst = sm + (size_t)st * FM_STAGE;
#pragma unroll
                    for (int gg = 0; gg < 2; ++gg)
#pragma unroll
                        for (int h = 0; h < 2; ++h) tma_load_3d(dst + (gg * 2 + h) * FM_BOX, ringmap, &full[st], sg.c0 + 16 * h, lo + 4 * gg, sg.traj0);
                }
            }
        }
        return;
    }
    // ---- consumers: warp = dof pair
    const int fr = lane >> 2, fk = lane & 3;                 // A: trajectory fr, age fk | B: age fk, step fr | C: trajectory fr, steps 2 fk, 2 fk + 1
    const int ri = fr * 4 + (3 - fk);                        // row of the box that holds (trajectory fr, age d0 + fk): slots run against the ages
    const unsigned offA = (unsigned)((warp >> 3) * FM_BOX + ri * 128 + (((warp & 7) ^ (ri & 7)) << 4));
    int st = 0;                                              // stage cursor: runs on across the segments, like the producer's
    unsigned par = 0;
    const unsigned char *sp = sm;
    const size_t sstride = (size_t)a.ntraj * a.ncp;
#pragma unroll 1
    for (int g = 0; g < FM_SPC; ++g) {
        const FarSeg sg = my[g];
        const int nk = 2 * (sg.a_hi - sg.a_lo);
        if (nk <= 0) continue;
        const int c = sg.c0 + 2 * warp;                      // dofs c, c + 1 (both < ncp: this warp is live)
        const int dlo = a.d0 + sg.a_lo * FM_SR;
        const double *kb0 = a.kT + (size_t)min(c, a.ncp - 1) * a.ldk + dlo + 2 + fr + fk, *kb1 = a.kT + (size_t)min(c + 1, a.ncp - 1) * a.ldk + dlo + 2 + fr + fk;
        double acc[2][NT][2];
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int n = 0; n < NT; ++n) acc[e][n][0] = acc[e][n][1] = 0.0;
        double f0[FM_RB], f1[FM_RB];                         // F(4 q) lives in slot q % FM_RB
#pragma unroll
        for (int i = 0; i < FM_RB; ++i) {
            f0[i] = __ldg(kb0 + 4 * i);
            f1[i] = __ldg(kb1 + 4 * i);
        }
        auto kstep = [&](auto ph, int q) {
            constexpr int I = decltype(ph)::value;           // q % FM_RB (FM_RB is even: I % 2 is the k-step inside the stage)
            if constexpr (I % 2 == 0) tma_mbar_wait(&full[st], par);
            // k-step 0 of a stage holds the younger ages = the upper slot group
            const double2 av = *reinterpret_cast<const double2 *>(sp + (I % 2 == 0 ? 2 * FM_BOX : 0) + offA);
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                dmma884(acc[0][n][0], acc[0][n][1], av.x, f0[(I + 2 * n) % FM_RB]);
                dmma884(acc[1][n][0], acc[1][n][1], av.y, f1[(I + 2 * n) % FM_RB]);
            }
            f0[I] = __ldg(kb0 + 4 * (q + FM_RB));            // first needed FM_RB - 2 (NT - 1) k-steps from now
            f1[I] = __ldg(kb1 + 4 * (q + FM_RB));
            if constexpr (I % 2 == 1) {
                __syncwarp();
                if (lane == 0) tma_mbar_arrive(&empty[st]);
                if (++st == FM_NST) {
                    st = 0;
                    par ^= 1u;
                    sp = sm;
                } else {
                    sp += FM_STAGE;
                }
            }
        };
        for (int q0 = 0; q0 < nk; q0 += FM_RB) fm_unroll<0, FM_RB>(kstep, q0, nk);
        if (c < a.ncp && sg.traj0 + fr < a.ntraj) {
            double *o = a.out + (((size_t)sg.slot * TBK) * a.ntraj + sg.traj0 + fr) * a.ncp + c;
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    *reinterpret_cast<double2 *>(o + (size_t)(8 * n + 2 * fk + i) * sstride) = make_double2(a.dt * acc[0][n][i], a.dt * acc[1][n][i]);
        }
    }
}

template <int TBK>
__global__ void __launch_bounds__((FM_W + 1) * 32, 1) k_tail_far_mma(const __grid_constant__ CUtensorMap ringmap, const FarMmaArgs a) {
    far_mma_body<TBK>(&ringmap, a, blockIdx.x);
}
// the slices of two baths in one launch (CTAs [0, n0) work for the first bath): one ramp and one drain per step instead of two
template <int TBK>
__global__ void __launch_bounds__((FM_W + 1) * 32, 1) k_tail_far_mma2(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                                                                         const FarMmaArgs a0, const FarMmaArgs a1, int n0) {
    if ((int)blockIdx.x < n0) far_mma_body<TBK>(&map0, a0, blockIdx.x);
    else far_mma_body<TBK>(&map1, a1, blockIdx.x - n0);
}

// tail[traj][c] = Near_s + sum_z Far[z][s]   (written where the phase kernels expect a single partial)
struct NearArgs {
    const double *ring, *kern, *far;
    double *out;
    const int *nslot;
    int ntraj, ml, ncp, head, s, nsplit, tb, chunks;
    double dt;
};
__device__ __forceinline__ void tail_near_body(const double *__restrict__ ring, const double *__restrict__ kern,
                                               const double *__restrict__ far, double *__restrict__ out, int ntraj, int ml,
                                               int ncp, int head, int s, int nsplit, double dt, int tb,
                                               const int *__restrict__ nslot, int chunks) {
    const int traj = blockIdx.x;
    const double *r = ring + (size_t)traj * ml * ncp;
    for (int c = threadIdx.x; c < ncp; c += blockDim.x) {
        double acc = 0.0;
        int slot = head;
        for (int j0 = 1; j0 <= s + 1; j0 += 8) {          // eight rows in flight per trip (the sum keeps its order)
            double kv[8], rv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool ok = j0 + u <= s + 1;
                kv[u] = ok ? kern[(size_t)(j0 + u) * ncp + c] : 0.0;
                rv[u] = ok ? r[(size_t)slot * ncp + c] : 0.0;
                slot = slot == 0 ? ml - 1 : slot - 1;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = fma(kv[u], rv[u], acc);
        }
        double f = 0.0;
        // partial far tails: a fixed count, or (tensor-pipe pass) the slots of this element's (32-dof chunk, 8-trajectory group) pair
        const int nz = nslot ? nslot[(traj >> 3) * chunks + (c >> 5)] : nsplit;
        for (int z = 0; z < nz; ++z) f += far[(((size_t)z * tb + s) * ntraj + traj) * ncp + c];
        out[(size_t)traj * ncp + c] = dt * acc + f;
    }
}
__global__ void __launch_bounds__(256) k_tail_near(const double *__restrict__ ring, const double *__restrict__ kern,
                                                    const double *__restrict__ far, double *__restrict__ out, int ntraj, int ml,
                                                    int ncp, int head, int s, int nsplit, double dt, int tb,
                                                    const int *__restrict__ nslot, int chunks) {
    tail_near_body(ring, kern, far, out, ntraj, ml, ncp, head, s, nsplit, dt, tb, nslot, chunks);
}
// the near passes of two baths in one launch (blockIdx.y = bath)
__global__ void __launch_bounds__(256) k_tail_near2(const NearArgs a0, const NearArgs a1) {
    const NearArgs &a = blockIdx.y ? a1 : a0;
    tail_near_body(a.ring, a.kern, a.far, a.out, a.ntraj, a.ml, a.ncp, a.head, a.s, a.nsplit, a.dt, a.tb, a.nslot, a.chunks);
}

// Kc[i][k] = K[i][cons_k]  (columns of K that multiply the constrained dofs)
__global__ void k_gather_kc(const double *__restrict__ K, int nph, int ld, const int *__restrict__ cidx, int ncons, int ldc, double *__restrict__ Kc) {
    const int i = blockIdx.x;
    for (int k = threadIdx.x; k < ncons; k += blockDim.x) Kc[(size_t)i * ldc + k] = K[(size_t)i * ld + cidx[k]];
}
// qc[traj][k] = q'[traj][cons_k]  (values the constraint is about to zero)
__global__ void k_gather_qc(const double *__restrict__ qn, int ld, const int *__restrict__ cidx, int ncons, int ldc, double *__restrict__ qc) {
    const int traj = blockIdx.x;
    for (int k = threadIdx.x; k < ncons; k += blockDim.x) qc[(size_t)traj * ldc + k] = qn[(size_t)traj * ld + cidx[k]];
}

// streamed noise rows: stage[i][traj][nc] (time slab slab0 + i, mod nmd) -> table[traj][nmd][ncp]
__global__ void k_scatter_rows(const double *__restrict__ stage, double *__restrict__ table, int slab0, int ntraj, int nc, int ncp, int nmd) {
    const int traj = blockIdx.x, i = blockIdx.y;
    const int slab = (slab0 + i) % nmd;
    const double *src = stage + ((size_t)i * ntraj + traj) * nc;
    double *dst = table + ((size_t)traj * nmd + slab) * ncp;
    for (int c = threadIdx.x; c < nc; c += blockDim.x) dst[c] = src[c];
}

__global__ void k_negate(double *__restrict__ x, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = -x[i];
}

__global__ void k_sum_slots(const double *__restrict__ cur, int nmd, int ntraj, double *__restrict__ sums) {
    const int traj = blockIdx.x * blockDim.x + threadIdx.x;
    if (traj >= ntraj) return;
    double s = 0.0;
    for (int k = 0; k < nmd; ++k) s += cur[(size_t)k * ntraj + traj];
    sums[traj] = s;
}


// ---- modal (eigenbasis) propagation ------------------------------------------------------------------------------------------
// md.setDyn keeps K = U diag(lam) U^T (md.py:266-281).  With Q = U^T q, Pi = U^T p the harmonic force is diagonal; only the
// bath dofs need real space.  Per step ONE gather product (K q')[cids] = Q' . (E lam)^T and ONE scatter product of the bath forces,
// W = (fC(t-1) + fA(t)) . E, E = U[cids,:]: 4 nph sum(nc) flops per trajectory instead of 2 nph^2 (specification: oracle ModalMD).
//   Pt_t    = Pi_t + h/2 E^T fA(t)                   momentum half-kicked by the bath forces of evaluation A
//   Q_{t+1} = Q_t + h Pt_t - h^2/2 lam Q_t           (md.py:392)
//   bath dofs, real space: evaluations A, B, C exactly as md.py:390-404 with g = (K q)[cids]
//   Pt_{t+1} = Pt_t - h/2 lam (Q_t + Q_{t+1}) + h/2 E^T (fC(t) + fA(t+1))
//   Pi.Pi   = Pt.Pt - h sum_b p_c.fA_b - h^2/4 sum_b |fA_b|^2        (rows of E are orthonormal, baths disjoint)
// As in the fused real-space path evaluations B, C of a step stay pending and run with evaluation A of the next one.
struct ModalArgs {
    BathSet bs;
    int off[MAXB];                 // first slot of bath b in the concatenated bath-dof arrays [ntraj][ncs]
    int ncs, ntraj, nmd, pending, doA, gsplit;
    long long t;                   // time of evaluation A (and of the noise slab / tail the pending B, C read)
    double dt;
    double *pc, *fA, *g, *sbuf, *ecorr;
    const double *gn;              // [gsplit][ntraj][ncs] K-slices of the gather product
};

template <int NBATH>
__global__ void __launch_bounds__(320) k_modal_bath(const ModalArgs a) {      // (launched with 256 threads by default)
    __shared__ double red[32];
    const int traj = blockIdx.x;
    const double h = a.dt;
    const int slab = (int)(a.t % a.nmd);
    const size_t row = (size_t)traj * a.ncs;
    double cur[NBATH], ec = 0.0;
#pragma unroll
    for (int b = 0; b < NBATH; ++b) cur[b] = 0.0;
    for (int e = threadIdx.x; e < a.ncs; e += blockDim.x) {
        int b = 0;
#pragma unroll
        for (int k = 1; k < NBATH; ++k)
            if (k < a.bs.nb && e >= a.off[k]) b = k;
        const int c = e - a.off[b];
        if (c >= a.bs.b[b].nc) {               // pad column
            a.sbuf[row + e] = 0.0;
            continue;
        }
        double x = a.pc[row + e], gold = a.g[row + e], s = 0.0;
        // the three evaluations of this element see the same noise row, friction coefficient and tail S'(t-1): fetched / summed once
        // (eigenbasis mode: diagonal time-local part, no dense matrices); the expression and its order are those of bath_force()
        const BathDev &bd = a.bs.b[b];
        const double nzv = bd.noise[((size_t)traj * bd.nmd + slab) * bd.ncp + c], kx = bd.c0 * bd.k0[c];
        double tail = 0.0;
        if (bd.use_tail)
            for (int z = 0; z < bd.nsplit; ++z) tail += bd.tailp[((size_t)z * a.ntraj + traj) * bd.ncp + c];
        auto force = [&](double xv) {
            double fb = nzv;
            fb -= kx * xv;
            if (bd.use_tail) fb -= tail;
            return fb;
        };
        if (a.pending) {                       // evaluations B and C of step t-1 (md.py:401-404): noise row t, tail S'(t-1)
            double gnew = a.gn[row + e];
            for (int z = 1; z < a.gsplit; ++z) gnew += a.gn[(size_t)z * a.ntraj * a.ncs + row + e];
            const double ph = x + (a.fA[row + e] - gold) * h / 2.0;
            double xi = ph, fb = 0.0;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                fb = force(xi);
                xi = ph + h * (fb - gnew) / 2.0;
            }
            x = xi;
            gold = gnew;
            s = fb;
            a.pc[row + e] = x;
            a.g[row + e] = gold;
        }
        if (a.doA) {                           // evaluation A of step t (md.py:383-398) on the bath dofs
            const double fa = force(x);
#pragma unroll
            for (int k = 0; k < NBATH; ++k)
                if (k == b) cur[k] += fa * x;
            a.bs.b[b].ring[((size_t)traj * a.bs.b[b].ml + (int)(a.t % a.bs.b[b].ml)) * a.bs.b[b].ncp + c] = x;
            ec -= h * x * fa + h * h / 4.0 * fa * fa;
            a.fA[row + e] = fa;
            s += fa;
        }
        a.sbuf[row + e] = s;
    }
    if (!a.doA) return;
    ec = block_sum(ec, red);
    if (threadIdx.x == 0) a.ecorr[traj] = ec;
#pragma unroll
    for (int b = 0; b < NBATH; ++b)
        if (b < a.bs.nb) {
            const double cc = block_sum(cur[b], red);
            if (threadIdx.x == 0) a.bs.b[b].cur[(size_t)slab * a.ntraj + traj] = cc;
        }
}

// Pt <- Pbase + h/2 W - [pending] h/2 lam (Qold + Q);  [advance] etot[t] = (Pt.Pt + ecorr)/2 and Q_{t+1} into Qold's buffer
__global__ void __launch_bounds__(256, 4) k_modal_pq(int nph, int ld, int ntraj, int nmd, long long t, double dt, int pending, int advance,
                                                      const double *__restrict__ lam, double *__restrict__ Pm, const double *__restrict__ Q,
                                                      double *__restrict__ Qo, const double *__restrict__ W, int wsplit, size_t wstride,
                                                      const double *__restrict__ ecorr, double *__restrict__ etot) {
    __shared__ double red[32];
    const int traj = blockIdx.x;
    const size_t row = (size_t)traj * ld;
    const double h = dt;
    double ke = 0.0;
    for (int i0 = 2 * threadIdx.x; i0 < nph; i0 += 2 * blockDim.x) {
        double2 w2 = *reinterpret_cast<const double2 *>(W + row + i0);
        for (int z = 1; z < wsplit; ++z) {
            const double2 v = *reinterpret_cast<const double2 *>(W + (size_t)z * wstride + row + i0);
            w2.x += v.x;
            w2.y += v.y;
        }
        const double2 p2 = *reinterpret_cast<const double2 *>(Pm + row + i0), q2 = *reinterpret_cast<const double2 *>(Q + row + i0);
        const double2 l2 = *reinterpret_cast<const double2 *>(lam + i0);
        double2 o2 = make_double2(0.0, 0.0);
        if (pending) o2 = *reinterpret_cast<const double2 *>(Qo + row + i0);
        double2 pn, qn;
        pn.x = p2.x + h / 2.0 * w2.x - (pending ? h / 2.0 * l2.x * (o2.x + q2.x) : 0.0);
        pn.y = p2.y + h / 2.0 * w2.y - (pending ? h / 2.0 * l2.y * (o2.y + q2.y) : 0.0);
        if (i0 + 1 >= nph) pn.y = 0.0;
        ke += pn.x * pn.x + pn.y * pn.y;
        *reinterpret_cast<double2 *>(Pm + row + i0) = pn;
        if (advance) {
            qn.x = q2.x + h * pn.x - h * h / 2.0 * l2.x * q2.x;
            qn.y = i0 + 1 < nph ? q2.y + h * pn.y - h * h / 2.0 * l2.y * q2.y : 0.0;
            *reinterpret_cast<double2 *>(Qo + row + i0) = qn;
        }
    }
    if (!advance) return;
    ke = block_sum(ke, red);
    if (threadIdx.x == 0) etot[(size_t)(t % nmd) * ntraj + traj] = 0.5 * (ke + ecorr[traj]);
}

// pc[traj][off + a] = p[traj][cids[a]]
__global__ void k_modal_gather_pc(const int *__restrict__ cids, int nc, int off, int ncs, int ld, const double *__restrict__ p, double *__restrict__ pc) {
    const int traj = blockIdx.x;
    for (int a = threadIdx.x; a < nc; a += blockDim.x) pc[(size_t)traj * ncs + off + a] = p[(size_t)traj * ld + cids[a]];
}
// g <- sum_z gn[z]
__global__ void k_modal_fold(const double *__restrict__ gn, int nsplit, size_t n, double *__restrict__ g) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double v = gn[e];
    for (int z = 1; z < nsplit; ++z) v += gn[(size_t)z * n + e];
    g[e] = v;
}
// tables from U [nph][ld] (columns = eigenvectors): UT[k][i] = U[i][k];  EL[off+a][k] = U[cid_a][k] lam_k;  ET[k][off+a] = U[cid_a][k]
__global__ void k_modal_transpose(const double *__restrict__ U, int nph, int ld, double *__restrict__ UT) {
    __shared__ double tile[32][33];
    const int i0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, k = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < nph && k < nph) ? U[(size_t)i * ld + k] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r, i = i0 + threadIdx.x;
        if (k < nph && i < nph) UT[(size_t)k * ld + i] = tile[threadIdx.x][r];
    }
}
__global__ void k_modal_tables(const double *__restrict__ U, const double *__restrict__ lam, int nph, int ld, const int *__restrict__ cids, int nc,
                               int off, int ncs, double *__restrict__ EL, double *__restrict__ ET) {
    const int a = blockIdx.x;
    const int i = cids[a];
    for (int k = threadIdx.x; k < nph; k += blockDim.x) {
        const double u = U[(size_t)i * ld + k];
        EL[(size_t)(off + a) * ld + k] = u * lam[k];
        ET[(size_t)k * ncs + off + a] = u;
    }
}
// max_ik |KU[i][k] - U[i][k] lam_k| and max |lam| as bit patterns of non-negative doubles (atomicMax on unsigned long long)
__global__ void k_modal_residual(const double *__restrict__ KU, const double *__restrict__ U, const double *__restrict__ lam, int nph, int ld,
                                 unsigned long long *__restrict__ out) {
    const int i = blockIdx.x;
    double m = 0.0;
    for (int k = threadIdx.x; k < nph; k += blockDim.x) m = fmax(m, fabs(KU[(size_t)i * ld + k] - U[(size_t)i * ld + k] * lam[k]));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
    if (i == 0) {
        double l = 0.0;
        for (int k = threadIdx.x; k < nph; k += blockDim.x) l = fmax(l, fabs(lam[k]));
        for (int o = 16; o > 0; o >>= 1) l = fmax(l, __shfl_xor_sync(0xffffffffu, l, o));
        if ((threadIdx.x & 31) == 0) atomicMax(out + 1, (unsigned long long)__double_as_longlong(l));
    }
}

}  // namespace

// ---------------------------------------------------------------- handle
struct sclmd_md {
    int nph = 0, ld = 0, ntraj = 0, nmd = 0, device = 0, nsm = 148;
    double dt = 0;
    long long t = 0;
    bool g_valid = false, have_dyn = false;
    cudaStream_t stc = nullptr;                  // copy stream: streamed noise rows overlap the running step
    cudaEvent_t evN = nullptr, evObs = nullptr;   // evObs: the observables of slab obs_slab (etot, currents) have been written
    cudaStream_t str = nullptr;                   // read-back stream of sclmd_md_get_step_observables
    long long obs_slab = -1;
    bool noise_pending = false;
    // time slabs of the pending streamed upload(s); np_mixed: uploads of different slab ranges are pending together
    int np_slab0 = 0, np_nslab = 0;
    bool np_mixed = false;
    bool pending_upload_covers(long long slab) const {
        return np_mixed || (int)(((slab % nmd) - np_slab0 + nmd) % nmd) < np_nslab;
    }
    cudaStream_t st = nullptr, st2 = nullptr;   // st2: the FP64-bound K.q GEMM overlaps the HBM-bound history tails
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evA = nullptr, evG = nullptr;
    bool overlap = true;
    DevBuf<double> K, q, p, G, Gn, phalf, p1, qn, etot, scratch;
    DevBuf<unsigned char> cons;
    bool has_cons = false, d_valid = false, kc_valid = false, use_corr = false;
    // md.AddPotential drivers (md.py:457-459, 481-485): the potential force comes from a host callback, one call per step
    // (two with constraints); everything else of the step stays on the device (sclmd_md_step_begin / sclmd_md_step_end)
    bool ext_force = false, ext_open = false;
    // md.f / md.fbaths (md.py:390-398, 411): force of evaluation C and bath forces of evaluation A of the last step, on request
    bool want_f = false;
    long long f_step = -1;      // the step (value of t after it) fC / fa belong to
    DevBuf<double> fC;
    bool corr() const { return has_cons && use_corr && !ext_force; }
    // CTA size of the per-trajectory evaluation kernels: small systems are latency-bound per CTA, so smaller CTAs (eight per SM
    // instead of four: the whole 1024-trajectory grid resident in one wave) finish sooner
    int phase_threads() const { return nph <= 1024 ? 128 : 256; }
    int ncons = 0, ldcons = 0;
    DevBuf<int> cidx;
    DevBuf<double> Kc, qc, Dc;   // constraint correction: D = q'[cons] . K[:,cons]^T
    std::vector<std::unique_ptr<Bath>> baths;
    int64_t launches = 0;
    // optional per-kernel timing (CUDA events on `st` around every tail / potential-force launch)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_kind;  // 0 = tail, 1 = potforce; events come in (start, stop) pairs
    size_t ev_used = 0, ev_open = 0;
    // 0 direct tail, 1 potforce, 2 far pass, 3 near; modal mode: 4 scatter GEMM, 5 gather GEMM, 6 bath-dof kernel, 7 modal update
    double prof_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool tail_block = true, far_tma = true, far_ws = true;
    SplitPlan gplan{0, 1, 0}, cplan{0, 1, 0};
    // Steps without history tails (every bath ml == 1: the reference's shipped examples) are launch-bound for small systems:
    // everything after evaluation A is captured once per G/Gn role into a CUDA graph and replayed; the step counter the
    // captured kernels need lives on the device (d_t).
    bool use_graphs = true, dt_synced = false;
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    int gnodes[2] = {0, 0}, gpar = 0;
    DevBuf<long long> d_t;
    void drop_graphs() {
        for (int i = 0; i < 2; ++i) {
            if (gexec[i]) cudaGraphExecDestroy(gexec[i]);
            gexec[i] = nullptr;
        }
        dt_synced = false;
    }

    void prof_begin(int kind, cudaStream_t stream = nullptr) {
        if (!profiling || kind < 0) return;
        if (!stream) stream = st;
        if (ev_used + 2 > ev_pool.size()) {
            for (int i = 0; i < 2; ++i) {
                cudaEvent_t e;
                cudaEventCreate(&e);
                ev_pool.push_back(e);
            }
        }
        ev_kind.push_back(kind);
        cudaEventRecord(ev_pool[ev_used], stream);
        ev_open = ev_used;
        ev_used += 2;
    }
    void prof_end(cudaStream_t stream = nullptr, int kind = 0) {
        if (!profiling || kind < 0) return;
        cudaEventRecord(ev_pool[ev_open + 1], stream ? stream : st);
    }
    void prof_collect() {  // call after the stream is synchronised
        for (size_t i = 0; i + 1 < ev_used; i += 2) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ev_pool[i], ev_pool[i + 1]) == cudaSuccess) {
                prof_ms[ev_kind[i / 2]] += ms;
                prof_n[ev_kind[i / 2]] += 1;
            }
        }
        ev_used = 0;
        ev_kind.clear();
    }

    BathSet view() const {
        BathSet s;
        memset(&s, 0, sizeof(s));
        s.nb = (int)baths.size();
        for (int i = 0; i < s.nb; ++i) {
            const Bath &b = *baths[i];
            BathDev &d = s.b[i];
            d.nc = b.nc; d.ncp = b.ncp; d.ml = b.ml; d.nsplit = (b.blocked && tail_block) ? 1 : b.nsplit;
            d.diag = b.kind == SCLMD_KERNEL_DIAG;
            d.has_lin = b.has_lin; d.use_tail = b.ml > 1; d.c0 = b.c0;
            d.inv = b.inv.p; d.cids = b.cids.p; d.k0 = b.kern.p; d.noise = b.noise.p; d.nmd = nmd;
            d.tailp = b.tailp.p; d.lin = b.lin.p; d.ring = b.ring.p; d.cur = b.cur.p;
            d.wt = b.WT.p; d.Kw = b.Kw;
            d.fa = want_f ? b.fa.p : nullptr;
            d.fc = want_f ? b.fc.p : nullptr;
        }
        return s;
    }
    bool any_lin() const {
        for (auto &b : baths) if (b->has_lin) return true;
        return false;
    }

    int potforce(const double *qsrc, double *dst, cudaStream_t stream = nullptr) {  // dst = qsrc . K^T   (md.py:467, sign applied by consumers)
        return gemm_nt(ntraj, nph, ld, qsrc, ld, K.p, ld, dst, ld, gplan, 1, stream ? stream : st);
    }
    int bath_lin(Bath &b, const double *x, const double *qq) {  // lin = [x|q][cids] . W^T
        k_gather_xq<<<ntraj, 128, 0, st>>>(b.cids.p, b.nc, b.ncp, b.Kw, ld, x, qq, b.xq.p);
        SCLMD_CUDA(cudaGetLastError());
        GemmArgs g{};
        g.M = ntraj; g.N = b.nc; g.Kseg = b.Kw; g.nseg = 1; g.segs_per_split = 1;
        g.A = b.xq.p; g.lda = b.Kw; g.B = b.W.p; g.ldb = b.Kw;
        g.C = b.lin.p; g.ldc = b.ncp; g.alpha = 1.0;
        SCLMD_CUDA(launch_dgemm(g, 1, st));
        launches += 2;
        return 0;
    }
    int tail_direct(Bath &b, int head) {  // partial tails from the ring with p_t already pushed at slot `head`
        if (b.ml <= 1) return 0;
        prof_begin(0);
        if (b.kind == SCLMD_KERNEL_DIAG) {
            const int cp_total = b.ncp / 2;
            const int cpt = std::min(cp_total, 256);
            const int rg = std::max(1, std::min(512 / cpt, b.ml));
            const int T = ntraj >= 4 ? 4 : 1;
            dim3 grid(cdiv(ntraj, T), b.nsplit, cdiv(cp_total, cpt));
            const int rps = cdiv(b.ml, b.nsplit);
            const int threads = round_up(cpt * rg, 32);
            const size_t sm = (size_t)rg * T * 2 * cpt * sizeof(double);
            if (T == 4)
                k_tail_diag<4><<<grid, threads, sm, st>>>(b.ring.p, b.kern.p, b.tailp.p, ntraj, b.ml, b.ncp, head, rps, cpt, rg, dt);
            else
                k_tail_diag<1><<<grid, threads, sm, st>>>(b.ring.p, b.kern.p, b.tailp.p, ntraj, b.ml, b.ncp, head, rps, cpt, rg, dt);
            SCLMD_CUDA(cudaGetLastError());
        } else {
            if (b.gemm_cfg == -2) {
                // persistent TMA / stream-K kernel over the rotating ring: A = ring as a rank-3 tensor [traj][slot][ncp] (box 16 x 1 x 128),
                // B = kernel [j][nc][ncp]; one K loop over (ml - 1) segments x ceil(ncp / 16) slabs, no K-slices of the tail in HBM
                TmaGemm g{};
                g.M = ntraj; g.N = b.nc; g.K = b.ncp; g.nbatch = 1; g.nseg = b.ml - 1;
                g.A = TmaOperand{b.ring.p, {(unsigned long long)b.ncp, (unsigned long long)b.ml, (unsigned long long)ntraj},
                                 {(unsigned long long)b.ncp, (unsigned long long)b.ml * b.ncp}, 0, 0};
                g.B = TmaOperand{b.kern.p, {(unsigned long long)b.ncp, (unsigned long long)b.nc, (unsigned long long)b.ml},
                                 {(unsigned long long)b.ncp, (unsigned long long)b.nc * b.ncp}, 0, 0};
                g.a_mode = 1; g.a_head = head; g.a_mod = b.ml; g.b_mode = 1; g.b_seg0 = 1;
                g.C = b.tailp.p; g.ldc = b.ncp; g.c_batch_stride = 0; g.alpha = dt;
                if (int e = launch_dgemm_tma(g, tws[0], nsm, st)) return e;
                prof_end();
                ++launches;
                return 0;
            }
            GemmArgs g{};
            g.M = ntraj; g.N = b.nc; g.Kseg = b.ncp; g.nseg = b.ml - 1;
            g.segs_per_split = cdiv(g.nseg, b.nsplit);
            g.A = b.ring.p; g.lda = (long long)b.ml * b.ncp; g.a_seg_stride = b.ncp; g.a_head = head; g.a_mod = b.ml;
            g.B = b.kern.p; g.ldb = b.ncp; g.b_seg_stride = (long long)b.nc * b.ncp; g.b_seg0 = 1;
            g.C = b.tailp.p; g.ldc = b.ncp; g.c_split_stride = (long long)ntraj * b.ncp; g.alpha = dt;
            SCLMD_CUDA(launch_dgemm(g, b.nsplit, st, b.gemm_cfg));
        }
        prof_end();
        ++launches;
        return 0;
    }

    // length of the time block of bath b: 16 steps; 32 with the windowed ring-pass kernel (one trajectory per CTA, A/B option)
    bool far32 = false;      // measured at the config-5 shape: 5.09 ms per 32-step pass (15.8 TFLOP/s) against 2.18 ms per 16-step pass (18.3): off by default
    bool far_mma = true;     // tensor-pipe far pass (k_tail_far_mma): 32-step blocks, ring streamed by TMA boxes
    bool mma_ok(const Bath &b) const { return far_mma && far_tma && far_ws && b.ml % FM_SR == 0 && !tma_disabled() && tma_encoder() != nullptr; }
    int block_len(const Bath &b) const { return (mma_ok(b) || (b.ncp <= 320 && far_tma && far_ws && far32)) ? 2 * TB : TB; }
    int prepare_mma(Bath &b) {
        b.ldk = round_up(b.ml + 64, 16);
        SCLMD_CUDA(b.kT.alloc_raw((size_t)b.ncp * b.ldk));
        k_kernel_transpose<<<dim3(cdiv(b.ldk, 32), cdiv(b.ncp, 32)), dim3(32, 8), 0, st>>>(b.kern.p, b.ml, b.ncp, b.ldk, b.kT.p);
        SCLMD_CUDA(cudaGetLastError());
        ++launches;
        const TmaOperand op{b.ring.p, {(unsigned long long)b.ncp, (unsigned long long)b.ml, (unsigned long long)ntraj},
                            {(unsigned long long)b.ncp, (unsigned long long)b.ml * b.ncp}, 4, 8};
        if (int e = tma_make_map(&b.ringmap, op)) return e;
        SCLMD_CUDA(cudaFuncSetAttribute(k_tail_far_mma<2 * TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FM_SMEM));
        SCLMD_CUDA(cudaFuncSetAttribute(k_tail_far_mma2<2 * TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FM_SMEM));
        if (int e = plan_far(b)) return e;
        b.mma_ready = true;
        return 0;
    }
    // FarPlan: the far-far part of a pass (ages >= 32 relative to the block start: ring rows that are final one block EARLIER) as
    // FM_SLICES slices of about one CTA per SM, so that the pass for block B + 1 is worked off one slice per step during block B; the
    // ages 0..31 (the rows of block B itself) follow as a short "mid" pass at the block boundary.  Every step then costs the same.
    // Linear order: (trajectory group, stage); a work item is a range of stages of one group (cut at most once by a group boundary) and
    // becomes one CTA per 32-dof chunk, side by side: the CTAs of an item run together and read neighbouring pieces of the same ring
    // rows (with the 256-byte L2 promotion of the tensor map every DRAM sector is then fetched once; chunk-major items measured 36 %
    // slower).  The k-th segment of a (chunk, group) pair writes partial slot k of that pair: a fixed order.
    static constexpr int FM_SLICES = 2 * TB;
    int plan_far(Bath &b) {
        const int chunks = cdiv(b.ncp, FM_DC), groups = cdiv(ntraj, 8), npairs = chunks * groups;
        const int na = (b.ml - 2 * TB) / FM_SR;                     // far-far stages per pair
        const long long total = (long long)groups * na;
        const long long per_slice = std::max<long long>(1, (total + FM_SLICES - 1) / FM_SLICES);
        // work items of a slice: about one CTA per SM.  The slices of two baths share a launch: when nsm / chunks leaves SMs over, every
        // second bath takes one item more, so that the two grids together come closer to two full waves (config 5: 150 + 140 CTAs on
        // 2 x 148 SMs instead of 140 + 140)
        size_t bidx = 0;
        for (size_t i = 0; i < baths.size(); ++i) if (baths[i].get() == &b) bidx = i;
        static const int items_alt = getenv("SCLMD_FAR_ITEMS_ALT") ? atoi(getenv("SCLMD_FAR_ITEMS_ALT")) : 1;      // A/B switch
        const int base_items = std::max(1, nsm / chunks);
        const bool spare = 2 * nsm - (2 * base_items + 1) * chunks >= 0;
        const int items = base_items + ((items_alt && spare && bidx % 2 == 0 && baths.size() > 1) ? 1 : 0);
        const long long per_item = std::max<long long>(8, (per_slice + items - 1) / items);
        std::vector<FarSeg> segs;
        std::vector<int> cnt(npairs, 0);
        b.slice0.assign(FM_SLICES + 1, 0);
        auto nlive_of = [&](int chunk) { return std::min(FM_W, (b.ncp - chunk * FM_DC) / 2); };
        long long pos = 0;
        for (int sl = 0; sl < FM_SLICES; ++sl) {
            b.slice0[sl] = (int)(segs.size() / FM_SPC);
            const long long send = std::min(total, (sl + 1) * per_slice);
            while (pos < send) {
                long long left = std::min(per_item, send - pos);
                int nseg = 0, gg[FM_SPC], lo[FM_SPC], hi[FM_SPC];
                while (left > 0 && nseg < FM_SPC) {
                    gg[nseg] = (int)(pos / na);
                    lo[nseg] = (int)(pos % na);
                    hi[nseg] = (int)std::min<long long>(na, lo[nseg] + left);
                    left -= hi[nseg] - lo[nseg];
                    pos += hi[nseg] - lo[nseg];
                    ++nseg;
                }
                for (int ch = 0; ch < chunks; ++ch)
                    for (int k = 0; k < FM_SPC; ++k) {
                        if (k < nseg) segs.push_back(FarSeg{ch * FM_DC, gg[k] * 8, lo[k], hi[k], cnt[gg[k] * chunks + ch]++, nlive_of(ch)});
                        else segs.push_back(FarSeg{0, 0, 0, 0, 0, nlive_of(ch)});
                    }
            }
        }
        b.slice0[FM_SLICES] = (int)(segs.size() / FM_SPC);
        // mid pass: ages 0..31, one CTA per pair (chunks side by side), slot right behind the pair's far-far slots
        std::vector<FarSeg> mid;
        int maxslot = 0;
        int nmid = 0;
        for (int gr = 0; gr < groups; gr += FM_SPC)                   // FM_SPC groups of one chunk per CTA: four stages each, so the fixed
            for (int ch = 0; ch < chunks; ++ch) {                      // cost of a CTA (pipeline fill, epilogue) is shared
                for (int k = 0; k < FM_SPC; ++k) {
                    if (gr + k < groups) {
                        const int pid = (gr + k) * chunks + ch;
                        mid.push_back(FarSeg{ch * FM_DC, (gr + k) * 8, 0, 2 * TB / FM_SR, cnt[pid]++, nlive_of(ch)});
                        maxslot = std::max(maxslot, cnt[pid]);
                    } else {
                        mid.push_back(FarSeg{0, 0, 0, 0, 0, nlive_of(ch)});
                    }
                }
                ++nmid;
            }
        b.fchunks = chunks; b.fnpairs = npairs; b.fslots = maxslot; b.fnmid = nmid;
        b.fhalf = (size_t)maxslot * 2 * TB * ntraj * b.ncp;
        if (segs.empty()) segs.push_back(FarSeg{0, 0, 0, 0, 0, FM_W});
        SCLMD_CUDA(b.fsegs.alloc_raw(segs.size()));
        SCLMD_CUDA(cudaMemcpy(b.fsegs.p, segs.data(), segs.size() * sizeof(FarSeg), cudaMemcpyHostToDevice));
        SCLMD_CUDA(b.fmid.alloc_raw(mid.size()));
        SCLMD_CUDA(cudaMemcpy(b.fmid.p, mid.data(), mid.size() * sizeof(FarSeg), cudaMemcpyHostToDevice));
        SCLMD_CUDA(b.fnslot.alloc_raw(npairs));
        SCLMD_CUDA(cudaMemcpy(b.fnslot.p, cnt.data(), npairs * sizeof(int), cudaMemcpyHostToDevice));
        if (b.far.n < 2 * b.fhalf) SCLMD_CUDA(b.far.alloc_raw(2 * b.fhalf));
        b.fcur = 0;
        b.far_t0 = -1;
        b.nxt_block = -1;
        return 0;
    }
    // 256 threads per trajectory (one column per thread -- 320 threads at config 5 -- measured slower: 0.063 against 0.057 ms per step for the
    // two baths, fewer CTAs per SM)
    static int near_threads(int ncp) { return std::min(256, std::max(64, round_up(ncp, 32))); }
    FarMmaArgs far_args(Bath &b, long long t0, int half, int d0) {
        auto fmod_ll = [](long long x, long long m) { long long r = x % m; return r < 0 ? r + m : r; };
        FarMmaArgs fa{};
        fa.kT = b.kT.p; fa.out = b.far.p + (size_t)half * b.fhalf; fa.ldk = b.ldk; fa.ntraj = ntraj; fa.ml = b.ml; fa.ncp = b.ncp;
        fa.base = (int)fmod_ll(t0 - 1, b.ml); fa.d0 = d0; fa.dt = dt;
        return fa;
    }
    // background work of a step, collected by tail_step_mma and launched by flush_far(): slice `sl` of the next block's pass for
    // every bath (two baths per launch), then -- after the last step of a block -- the mid pass and the change of halves
    struct FarPending { Bath *b; long long t1; int s0, s1; NearArgs near; };
    std::vector<FarPending> far_pending;
    int flush_far() {
        std::vector<FarPending> pend;
        pend.swap(far_pending);
        // the near passes S'(tt) first (the next step's bath-dof kernel waits for them), two baths per launch
        for (size_t k = 0; k < pend.size(); k += 2) {
            prof_begin(3);
            if (k + 1 < pend.size()) {
                k_tail_near2<<<dim3(ntraj, 2), near_threads(std::max(pend[k].near.ncp, pend[k + 1].near.ncp)), 0, st>>>(pend[k].near, pend[k + 1].near);
            } else {
                const NearArgs &a = pend[k].near;
                k_tail_near<<<ntraj, near_threads(a.ncp), 0, st>>>(a.ring, a.kern, a.far, a.out, a.ntraj, a.ml, a.ncp, a.head, a.s, a.nsplit, a.dt, a.tb, a.nslot, a.chunks);
            }
            prof_end();
            SCLMD_CUDA(cudaGetLastError());
            ++launches;
        }
        size_t i = 0;
        while (i < pend.size()) {
            FarPending &p0 = pend[i];
            // a second bath with the same slice range (the normal case) shares the launch
            if (p0.s0 > p0.s1) { ++i; continue; }          // (near pass only)
            if (i + 1 < pend.size() && pend[i + 1].s0 == p0.s0 && pend[i + 1].s1 == p0.s1 && p0.s0 == p0.s1 && pend[i + 1].b != p0.b) {
                FarPending &p1 = pend[i + 1];
                Bath &b0 = *p0.b, &b1 = *p1.b;
                const int sl = p0.s0;
                const int n0 = b0.slice0[sl + 1] - b0.slice0[sl], n1 = b1.slice0[sl + 1] - b1.slice0[sl];
                if (n0 > 0 && n1 > 0) {
                    FarMmaArgs a0 = far_args(b0, p0.t1, b0.fcur ^ 1, 2 * TB), a1 = far_args(b1, p1.t1, b1.fcur ^ 1, 2 * TB);
                    a0.segs = b0.fsegs.p + (size_t)b0.slice0[sl] * FM_SPC;
                    a1.segs = b1.fsegs.p + (size_t)b1.slice0[sl] * FM_SPC;
                    prof_begin(2);
                    k_tail_far_mma2<2 * TB><<<n0 + n1, (FM_W + 1) * 32, FM_SMEM, st>>>(b0.ringmap, b1.ringmap, a0, a1, n0);
                    prof_end();
                    SCLMD_CUDA(cudaGetLastError());
                    ++launches;
                    i += 2;
                    continue;
                }
            }
            if (int e = far_slices(*p0.b, p0.t1, p0.b->fcur ^ 1, p0.s0, p0.s1)) return e;
            ++i;
        }
        for (auto &p : pend)
            if (p.s0 <= p.s1 && p.s1 == FM_SLICES - 1) {
                Bath &b = *p.b;
                if (int e = far_mid(b, p.t1, b.fcur ^ 1)) return e;
                b.fcur ^= 1;
                b.far_t0 = p.t1;
                b.nxt_block = -1;
            }
        return 0;
    }
    // slices [s0, s1] of the far-far pass of the block starting at t0 into half `half`
    int far_slices(Bath &b, long long t0, int half, int s0, int s1) {
        auto fmod_ll = [](long long x, long long m) { long long r = x % m; return r < 0 ? r + m : r; };
        FarMmaArgs fa{};
        fa.kT = b.kT.p; fa.out = b.far.p + (size_t)half * b.fhalf; fa.ldk = b.ldk; fa.ntraj = ntraj; fa.ml = b.ml; fa.ncp = b.ncp;
        fa.base = (int)fmod_ll(t0 - 1, b.ml); fa.d0 = 2 * TB; fa.dt = dt;
        for (int sl = s0; sl <= s1; ++sl) {
            const int n = b.slice0[sl + 1] - b.slice0[sl];
            if (n <= 0) continue;
            fa.segs = b.fsegs.p + (size_t)b.slice0[sl] * FM_SPC;
            prof_begin(2);
            k_tail_far_mma<2 * TB><<<n, (FM_W + 1) * 32, FM_SMEM, st>>>(b.ringmap, fa);
            prof_end();
            SCLMD_CUDA(cudaGetLastError());
            ++launches;
        }
        return 0;
    }
    int far_mid(Bath &b, long long t0, int half) {
        auto fmod_ll = [](long long x, long long m) { long long r = x % m; return r < 0 ? r + m : r; };
        FarMmaArgs fa{};
        fa.kT = b.kT.p; fa.out = b.far.p + (size_t)half * b.fhalf; fa.ldk = b.ldk; fa.ntraj = ntraj; fa.ml = b.ml; fa.ncp = b.ncp;
        fa.base = (int)fmod_ll(t0 - 1, b.ml); fa.d0 = 0; fa.dt = dt; fa.segs = b.fmid.p;
        prof_begin(2);
        k_tail_far_mma<2 * TB><<<b.fnmid, (FM_W + 1) * 32, FM_SMEM, st>>>(b.ringmap, fa);
        prof_end();
        SCLMD_CUDA(cudaGetLastError());
        ++launches;
        return 0;
    }
    // tensor-pipe far pass, one slice per step: see plan_far
    int tail_step_mma(Bath &b, long long tt) {
        auto fmod_ll = [](long long x, long long m) { long long r = x % m; return r < 0 ? r + m : r; };
        if (!b.mma_ready) if (int e = prepare_mma(b)) return e;
        const int tb = 2 * TB, s = (int)fmod_ll(tt, tb);
        const long long t0 = tt - s, t1 = t0 + tb;
        if (b.far_t0 != t0) {            // nothing usable for this block (first step, state or history just set): the whole pass now
            if (int e = far_slices(b, t0, b.fcur, 0, FM_SLICES - 1)) return e;
            if (int e = far_mid(b, t0, b.fcur)) return e;
            b.far_t0 = t0;
            b.nxt_block = -1;
        }
        NearArgs na{};
        na.ring = b.ring.p; na.kern = b.kern.p; na.far = b.far.p + (size_t)b.fcur * b.fhalf; na.out = b.tailp.p; na.nslot = b.fnslot.p;
        na.ntraj = ntraj; na.ml = b.ml; na.ncp = b.ncp; na.head = (int)fmod_ll(tt, b.ml); na.s = s; na.nsplit = 0; na.tb = tb; na.chunks = b.fchunks;
        na.dt = dt;
        // the next block: its far-far ages are rows older than t0, all final -- slice s now (and any slices a late start has skipped);
        // after the last step of the block its own 32 rows are in the ring too: mid pass, and the halves change roles
        if (b.nxt_block != t1) {
            b.nxt_block = t1;
            b.nxt_done = 0;
        }
        // near pass and slice are launched by flush_far() once every bath has queued them (two baths per launch)
        if (b.nxt_done <= s) {
            far_pending.push_back(FarPending{&b, t1, b.nxt_done, s, na});
            b.nxt_done = s + 1;
        } else {
            far_pending.push_back(FarPending{&b, t1, 1, 0, na});          // the slices of this step are done already: near pass only
        }
        return 0;
    }
    // friction tail S'(tt) of step tt (ring already holds p_tt)
    int tail_step(Bath &b, long long tt) {
        if (b.ml <= 1) return 0;
        auto fmod_ll = [](long long a, long long m) { long long r = a % m; return r < 0 ? r + m : r; };
        if (!(b.blocked && tail_block)) return tail_direct(b, (int)fmod_ll(tt, b.ml));
        if (mma_ok(b)) return tail_step_mma(b, tt);
        const int tb = block_len(b);
        const long long t0 = tt - fmod_ll(tt, tb);
        const int T = ntraj >= 4 ? 4 : 1;
        if (b.far_t0 != t0) {
            const int base = (int)fmod_ll(t0 - 1, b.ml);
            const int ntiles = cdiv(b.ncp, 256), ct = cdiv(b.ncp, ntiles);
            const int aps = round_up(cdiv(b.ml, b.far_nsplit), tb);
            dim3 grid(cdiv(ntraj, T), b.far_nsplit, ntiles);
            prof_begin(2);
            if (tb == 2 * TB) {
                auto kern = k_tail_far_wsx<1, 2 * TB, 4, 20, 36>;
                static bool cfg = false;
                if (!cfg) {
                    SCLMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                    cfg = true;
                }
                const size_t smx = (size_t)20 * 1 * 4 * b.ncp * sizeof(double);    // twenty stages x 1 trajectory x 4 rows
                kern<<<dim3(ntraj, b.far_nsplit), round_up(b.ncp, 32) + 32, smx, st>>>(b.ring.p, b.kern.p, b.far.p, ntraj, b.ml, b.ncp, base, aps, dt);
            } else if (b.ncp <= 320 && ntraj >= 2 && far_tma) {
                static bool cfg = false;
                if (!cfg) {
                    SCLMD_CUDA(cudaFuncSetAttribute(k_tail_far_tma<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                    SCLMD_CUDA(cudaFuncSetAttribute(k_tail_far_ws<2, 8, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                    cfg = true;
                }
                const size_t sm = (size_t)2 * 2 * TB * b.ncp * sizeof(double);     // two stages x 2 trajectories x 16 rows
                const size_t smws = (size_t)5 * 2 * 8 * b.ncp * sizeof(double);    // five stages x 2 trajectories x 8 rows
                if (far_ws)
                    k_tail_far_ws<2, 8, 5><<<dim3(cdiv(ntraj, 2), b.far_nsplit), round_up(b.ncp, 32) + 32, smws, st>>>(b.ring.p, b.kern.p, b.far.p,
                                                                                                                 ntraj, b.ml, b.ncp, base, aps, dt);
                else
                    k_tail_far_tma<2, 2><<<dim3(cdiv(ntraj, 2), b.far_nsplit), round_up(b.ncp, 32), sm, st>>>(b.ring.p, b.kern.p, b.far.p, ntraj,
                                                                                                         b.ml, b.ncp, base, aps, dt);
            } else if (T == 4)
                k_tail_far<4><<<grid, round_up(ct, 32), 0, st>>>(b.ring.p, b.kern.p, b.far.p, ntraj, b.ml, b.ncp, base, aps, ct, dt);
            else
                k_tail_far<1><<<grid, round_up(ct, 32), 0, st>>>(b.ring.p, b.kern.p, b.far.p, ntraj, b.ml, b.ncp, base, aps, ct, dt);
            prof_end();
            SCLMD_CUDA(cudaGetLastError());
            b.far_t0 = t0;
            ++launches;
        }
        prof_begin(3);
        k_tail_near<<<ntraj, near_threads(b.ncp), 0, st>>>(b.ring.p, b.kern.p, b.far.p, b.tailp.p, ntraj, b.ml, b.ncp, (int)fmod_ll(tt, b.ml), (int)(tt - t0),
                                           b.far_nsplit, dt, tb, nullptr, 0);
        prof_end();
        SCLMD_CUDA(cudaGetLastError());
        ++launches;
        return 0;
    }

    // everything of a step after evaluation A and the history tails: K.q', evaluations B and C, constraint correction.
    // `tptr` != nullptr: enqueue for CUDA-graph capture (the step counter is read from the device).
    int enqueue_rest(const BathSet &bs, bool lin, const long long *tptr) {
        const unsigned char *cm = has_cons ? cons.p : nullptr;
        const size_t gs = (size_t)ntraj * ld;
        if (corr()) {
            k_gather_qc<<<ntraj, 128, 0, st>>>(qn.p, ld, cidx.p, ncons, ldcons, qc.p);
            SCLMD_CUDA(cudaGetLastError());
            ++launches;
        }
        auto bc = [&](const double *x, double *pout, int final, int fused) -> cudaError_t {
            if (bs.nb <= 2) k_phase_bc<2><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t, tptr, dt, x, phalf.p, Gn.p, gplan.nsplit, gs, pout, qn.p, q.p, cm, final, fused, want_f ? fC.p : nullptr);
            else if (bs.nb <= 4) k_phase_bc<4><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t, tptr, dt, x, phalf.p, Gn.p, gplan.nsplit, gs, pout, qn.p, q.p, cm, final, fused, want_f ? fC.p : nullptr);
            else k_phase_bc<MAXB><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t, tptr, dt, x, phalf.p, Gn.p, gplan.nsplit, gs, pout, qn.p, q.p, cm, final, fused, want_f ? fC.p : nullptr);
            ++launches;
            return cudaGetLastError();
        };
        if (!lin) {
            SCLMD_CUDA(bc(nullptr, p.p, 1, 1));
        } else {
            for (auto &b : baths)
                if (b->has_lin) if (int e = bath_lin(*b, phalf.p, qn.p)) return e;
            SCLMD_CUDA(bc(phalf.p, p1.p, 0, 0));
            for (auto &b : baths)
                if (b->has_lin) if (int e = bath_lin(*b, p1.p, qn.p)) return e;
            SCLMD_CUDA(bc(p1.p, p.p, 1, 0));
        }
        if (corr()) {
            // q_{t+1} = constrain(q') != q' (md.py:407-408), so the reference's force cache misses (md.py:449) and K.q is
            // evaluated again.  Here: K.q_{t+1} = K.q' - K[:,cons].q'[cons], a GEMM over the constrained columns only.
            GemmArgs g{};
            g.M = ntraj; g.N = nph; g.Kseg = ldcons; g.nseg = 1; g.segs_per_split = 1;
            g.A = qc.p; g.lda = ldcons; g.B = Kc.p; g.ldb = ldcons; g.C = Dc.p; g.ldc = ld; g.alpha = 1.0;
            SCLMD_CUDA(launch_dgemm(g, 1, st, cplan.cfg));
            ++launches;
        }
        return 0;
    }
    // whole run(n) in one cooperative launch (k_md_persist): few trajectories, time-local diagonal baths, state fits shared memory
    bool use_persist = true;
    DevBuf<double> xbuf;
    bool persist_ok() const {
        if (!use_persist || ext_force || want_f || profiling || ntraj > PS_MAXT || nph > 256 * PS_EPT || baths.size() > 2) return false;
        for (auto &b : baths)
            if (b->ml > 1 || b->has_lin || b->kind != SCLMD_KERNEL_DIAG) return false;
        return true;
    }
    int run_persist(long long nsteps) {
        BathSet bs = view();
        if (!g_valid) {
            if (int e = potforce(q.p, G.p)) return e;
            g_valid = true;
            d_valid = false;
        }
        const size_t n = (size_t)ntraj * ld;
        if (gplan.nsplit > 1 || d_valid) {     // fold the K-slices (and a pending constraint correction) into slice 0
            k_fold_g<<<cdiv((int)n, 256), 256, 0, st>>>(G.p, gplan.nsplit, n, d_valid ? Dc.p : nullptr);
            SCLMD_CUDA(cudaGetLastError());
            d_valid = false;
        }
        if (!xbuf.p) SCLMD_CUDA(xbuf.alloc(4 * n));
        if (noise_pending) {
            SCLMD_CUDA(cudaStreamWaitEvent(st, evN, 0));
            noise_pending = false;
        }
        PersistArgs a{};
        a.bs = bs; a.nph = nph; a.ld = ld; a.ntraj = ntraj; a.nmd = nmd; a.has_cons = has_cons ? 1 : 0;
        a.t0 = t; a.nsteps = nsteps; a.dt = dt; a.K = K.p; a.q = q.p; a.p = p.p; a.G = G.p; a.cons = cons.p; a.etot = etot.p; a.xbuf = xbuf.p;
        const size_t smem = n * sizeof(double);
        const int grid = std::max(1, std::min(nsm, cdiv(nph, 8)));
        void *args[] = {&a};
        const void *fn = ntraj == 1 ? (const void *)k_md_persist<2, 1> : (const void *)k_md_persist<2, 2>;
        SCLMD_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, smem, st));
        ++launches;
        t += nsteps;
        dt_synced = false;
        obs_slab = -1;               // every slab of the run is written by this one launch: read-backs take the synchronising path
        if (gplan.nsplit > 1) {      // the step kernels add the K-slices of G: the other slices must read as zero
            SCLMD_CUDA(cudaMemsetAsync(G.p + n, 0, (size_t)(gplan.nsplit - 1) * n * sizeof(double), st));
        }
        return 0;
    }
    // whole run(n) of an ensemble in one launch without grid synchronisation (k_md_ens): time-local diagonal baths, nph <= 768
    bool use_ens = true, kfrag_valid = false;
    DevBuf<double> kfrag;
    DevBuf<int> cidx8;
    int ens_nk8 = 0, ens_nkc8 = 0, ens_ntile = 0;
    bool ens_ok() const {
        if (!use_persist || !use_ens || ext_force || want_f || profiling || !have_dyn || ntraj <= PS_MAXT || nph > 760 || baths.size() > 4) return false;      // 12 tiles x 8 warps x 8 rows = 768
        int nlin = 0;
        for (auto &b : baths) {
            if (b->ml > 1) return false;
            if (b->has_lin) {          // one small dense bath (full friction matrix and / or the q-dependent matrices of an ebath)
                if (++nlin > 1 || b->ncp > EN_LC) return false;
            } else if (b->kind != SCLMD_KERNEL_DIAG) {
                return false;
            }
        }
        return true;
    }
    int ens_lin_bath() const {
        for (size_t i = 0; i < baths.size(); ++i) if (baths[i]->has_lin) return (int)i;
        return -1;
    }
    int build_kfrag() {
        ens_nk8 = round_up(cdiv(nph, 8), EN_D);
        const int need = cdiv(cdiv(nph, 8), 8);
        ens_ntile = need <= 4 ? 4 : need <= 5 ? 5 : need <= 8 ? 8 : need <= 10 ? 10 : 12;      // tiles per warp: groups of 4 or 5
        ens_nkc8 = 0;
        if (has_cons) {
            std::vector<unsigned char> m(nph);
            SCLMD_CUDA(cudaMemcpyAsync(m.data(), cons.p, nph, cudaMemcpyDeviceToHost, st));
            SCLMD_CUDA(cudaStreamSynchronize(st));
            std::vector<int> idx;
            for (int i = 0; i < nph; ++i) if (m[i]) idx.push_back(i);
            ens_nkc8 = round_up(cdiv((int)idx.size(), 8), EN_D);
            idx.resize((size_t)ens_nkc8 * 8, -1);
            SCLMD_CUDA(cidx8.alloc(idx.size()));
            SCLMD_CUDA(cudaMemcpy(cidx8.p, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
        SCLMD_CUDA(kfrag.alloc((size_t)8 * ens_ntile * (ens_nk8 + ens_nkc8) * 64));
        k_build_kfrag<<<8 * ens_ntile, 256, 0, st>>>(K.p, nph, ld, ens_nk8, ens_nkc8, cidx8.p, kfrag.p);
        SCLMD_CUDA(cudaGetLastError());
        ++launches;
        kfrag_valid = true;
        return 0;
    }
    template <int NTILE, int G, int W>
    int launch_ens(const EnsArgs &a, size_t smem) {
        auto go = [&](auto kernel) -> int {
            SCLMD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kernel<<<cdiv(ntraj, EN_T), W * 32, smem, st>>>(a);
            SCLMD_CUDA(cudaGetLastError());
            return 0;
        };
        if (baths.size() > 2 || ens_lin_bath() >= 0)       // up to four baths, one of them dense
            return has_cons ? go(k_md_ens<4, NTILE, G, true, true, W>) : go(k_md_ens<4, NTILE, G, false, true, W>);
        return has_cons ? go(k_md_ens<2, NTILE, G, true, false, W>) : go(k_md_ens<2, NTILE, G, false, false, W>);
    }
    int run_ens(long long nsteps) {
        if (!kfrag_valid) if (int e = build_kfrag()) return e;
        BathSet bs = view();
        if (!g_valid) {
            if (int e = potforce(q.p, G.p)) return e;
            g_valid = true;
            d_valid = false;
        }
        const size_t n = (size_t)ntraj * ld;
        if (gplan.nsplit > 1 || d_valid) {     // fold the K-slices (and a pending constraint correction) into slice 0
            k_fold_g<<<cdiv((int)n, 256), 256, 0, st>>>(G.p, gplan.nsplit, n, d_valid ? Dc.p : nullptr);
            SCLMD_CUDA(cudaGetLastError());
            d_valid = false;
            ++launches;
        }
        if (noise_pending) {
            SCLMD_CUDA(cudaStreamWaitEvent(st, evN, 0));
            noise_pending = false;
        }
        EnsArgs a{};
        a.bs = bs; a.nph = nph; a.ld = ld; a.ntraj = ntraj; a.nmd = nmd; a.has_cons = has_cons ? 1 : 0;
        a.nk8 = ens_nk8; a.nkc8 = ens_nkc8;
        a.lds = round_up(8 * ens_nk8, 16) + 4;            // == 4 (mod 16): the eight trajectory rows of an A fragment hit distinct banks
        a.ncpmax = 2;
        for (auto &b : baths) a.ncpmax = std::max(a.ncpmax, b->ncp);
        a.t0 = t; a.nsteps = nsteps; a.dt = dt; a.kfrag = kfrag.p; a.cidx8 = cidx8.p;
        a.q = q.p; a.p = p.p; a.G = G.p; a.cons = cons.p; a.etot = etot.p;
        a.lin_bath = ens_lin_bath();
        const bool wide = baths.size() > 2 || a.lin_bath >= 0;
        // state + noise rows + (dense bath scratch | per-dof bath tables of the two-bath kernel: coefficient and position)
        const size_t smem = ((size_t)4 * EN_T * a.lds + (size_t)std::max<size_t>(baths.size(), 1) * EN_T * a.ncpmax + (wide ? 3 * EN_T * EN_LC : 0)) * sizeof(double) +
                            (wide ? 0 : (size_t)2 * a.lds * (sizeof(double) + sizeof(int)));
        if (smem > 226 * 1024) return 1;                  // caller falls back to the launch chain
        // ens_ntile tiles per warp of an 8-warp CTA; even counts are split over 16 warps (SCLMD_ENS_WARPS=8 keeps eight)
        static const bool wide16 = !(getenv("SCLMD_ENS_WARPS") && atoi(getenv("SCLMD_ENS_WARPS")) == 8);
        int e;
        if (wide16 && ens_ntile == 8) e = launch_ens<4, 4, 16>(a, smem);
        else if (wide16 && ens_ntile == 10) e = launch_ens<5, 5, 16>(a, smem);
        else if (wide16 && ens_ntile == 12) e = launch_ens<6, 3, 16>(a, smem);
        else e = ens_ntile == 4 ? launch_ens<4, 4, 8>(a, smem) : ens_ntile == 5 ? launch_ens<5, 5, 8>(a, smem) : ens_ntile == 8 ? launch_ens<8, 4, 8>(a, smem)
               : ens_ntile == 10 ? launch_ens<10, 5, 8>(a, smem) : launch_ens<12, 4, 8>(a, smem);
        if (e) return e;
        ++launches;
        t += nsteps;
        dt_synced = false;
        obs_slab = -1;
        if (gplan.nsplit > 1) SCLMD_CUDA(cudaMemsetAsync(G.p + n, 0, (size_t)(gplan.nsplit - 1) * n * sizeof(double), st));
        return 0;
    }
    void finish_step() {      // host-side state after a step
        if (corr()) {
            d_valid = true;
            std::swap(G.p, Gn.p);
            gpar ^= 1;
        } else if (has_cons) {
            g_valid = false;
        } else {
            std::swap(G.p, Gn.p);
            gpar ^= 1;
        }
        ++t;
        if (want_f) f_step = t;
    }

    // dst (K-slice 0) = -f, the other K-slices zero: the phase kernels add the slices and apply the sign themselves
    int upload_force(double *dst, const double *f) {
        const size_t n = (size_t)ntraj * ld, w = nph * sizeof(double);
        SCLMD_CUDA(cudaMemcpy2DAsync(dst, ld * sizeof(double), f, w, w, ntraj, cudaMemcpyHostToDevice, st));
        k_negate<<<cdiv((int)n, 256), 256, 0, st>>>(dst, n);
        SCLMD_CUDA(cudaGetLastError());
        ++launches;
        if (gplan.nsplit > 1) SCLMD_CUDA(cudaMemsetAsync(dst + n, 0, (size_t)(gplan.nsplit - 1) * n * sizeof(double), st));
        return 0;
    }
    // first half of a step with a host force driver: evaluation A, q' and the history tails; q' goes back to the caller
    int ext_begin(double *q_trial) {
        BathSet bs = view();
        if (any_lin())
            for (auto &b : baths)
                if (b->has_lin) if (int e = bath_lin(*b, p.p, q.p)) return e;
        if (bs.nb <= 2) k_phase_a<2><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t, dt, q.p, p.p, G.p, gplan.nsplit, (size_t)ntraj * ld, nullptr, phalf.p, qn.p, etot.p);
        else if (bs.nb <= 4) k_phase_a<4><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t, dt, q.p, p.p, G.p, gplan.nsplit, (size_t)ntraj * ld, nullptr, phalf.p, qn.p, etot.p);
        else k_phase_a<MAXB><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t, dt, q.p, p.p, G.p, gplan.nsplit, (size_t)ntraj * ld, nullptr, phalf.p, qn.p, etot.p);
        SCLMD_CUDA(cudaGetLastError());
        SCLMD_CUDA(cudaEventRecord(evObs, st));
        obs_slab = t % nmd;
        ++launches;
        dt_synced = false;
        { for (auto &b : baths) if (int e = tail_step(*b, t)) return e; if (int e = flush_far()) return e; }
        const size_t w = nph * sizeof(double);
        SCLMD_CUDA(cudaMemcpy2DAsync(q_trial, w, qn.p, ld * sizeof(double), w, ntraj, cudaMemcpyDeviceToHost, st));
        SCLMD_CUDA(cudaStreamSynchronize(st));
        ext_open = true;
        return 0;
    }
    // second half: the driver's force at q' arrives, evaluations B and C, constraint
    int ext_end(const double *f_trial) {
        if (int e = upload_force(Gn.p, f_trial)) return e;
        if (noise_pending) {
            SCLMD_CUDA(cudaStreamWaitEvent(st, evN, 0));
            noise_pending = false;
        }
        if (int e = enqueue_rest(view(), any_lin(), nullptr)) return e;
        finish_step();
        ext_open = false;
        return 0;
    }

    // Lazy evaluations B, C: without constraints and with bath forces diagonal in x, a step ends after its K.q' GEMM and leaves B, C
    // pending (bc_pending; t and the G/Gn roles are already those of the next step).  If another step follows -- in the same
    // sclmd_md_run call or a later one -- they run fused with its evaluation A (k_phase_bca); anything that reads or changes the
    // state calls flush() first, which runs them alone.  Nothing is ever computed ahead of the reference's order.
    bool bc_pending = false, fuse_bca = true;
    bool lazy_ok(bool lin) const { return fuse_bca && !lin && !has_cons && !want_f && !ext_force; }
    int flush() {
        if (modal_state) return modal_flush();
        if (!bc_pending) return 0;
        if (noise_pending) {
            SCLMD_CUDA(cudaStreamWaitEvent(st, evN, 0));
            noise_pending = false;
        }
        BathSet bs = view();
        const size_t gs = (size_t)ntraj * ld;
        if (bs.nb <= 2) k_phase_bc<2><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t - 1, nullptr, dt, nullptr, phalf.p, G.p, gplan.nsplit, gs, p.p, qn.p, q.p, nullptr, 1, 1, nullptr);
        else if (bs.nb <= 4) k_phase_bc<4><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t - 1, nullptr, dt, nullptr, phalf.p, G.p, gplan.nsplit, gs, p.p, qn.p, q.p, nullptr, 1, 1, nullptr);
        else k_phase_bc<MAXB><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t - 1, nullptr, dt, nullptr, phalf.p, G.p, gplan.nsplit, gs, p.p, qn.p, q.p, nullptr, 1, 1, nullptr);
        SCLMD_CUDA(cudaGetLastError());
        ++launches;
        bc_pending = false;
        return 0;
    }

    // ---- modal (eigenbasis) mode: see k_modal_bath.  State lives in (Q, Pt / Pi, p_c, g) while modal_state is set; the
    // real-space arrays q, p are stale then and sync_real() brings them back.
    bool have_modes = false, modal_on = true, modal_state = false, m_pending = false, modal_tables = false;
    DevBuf<double> mU, mUT, mlam, mEL, mET, mQ[2], mP, mW, mpc, mfA, mg, mgn, msb, mec;
    int mqi = 0;               // mQ[mqi] = Q_t; mQ[mqi ^ 1] = Q_{t-1} while evaluations B, C are pending
    int ncs = 0, moff[MAXB] = {0, 0, 0, 0, 0, 0, 0, 0};
    SplitPlan splan{0, 1, 0}, gaplan{0, 1, 0};     // scatter (K = ncs) and gather (K = nph) products
    bool modal_ok() const {
        if (!have_modes || !modal_on || !have_dyn || has_cons || ext_force || want_f || baths.empty()) return false;
        std::vector<char> used(nph, 0);
        int tot = 0;
        for (auto &b : baths) {
            if (b->kind != SCLMD_KERNEL_DIAG || b->has_lin) return false;
            for (int c : b->cids_h) {
                if (used[c]) return false;       // E_b E_b'^T = delta_bb' needs disjoint dof sets
                used[c] = 1;
            }
            tot += b->ncp;
        }
        return 2 * tot <= nph;                    // gather + scatter (4 nph sum nc) against K.q (2 nph^2)
    }
    SplitPlan one_plan() const { return tma_usable(ntraj, nph) ? SplitPlan{-2, 1, ld} : SplitPlan{ntraj > 64 ? 0 : (ntraj > 16 ? 1 : 2), 1, ld}; }
    TmaWorkspace tws[2];       // stream-K scratch of the TMA GEMM: one per stream it is launched on (st, st2)
    // cfg -2 = persistent TMA / stream-K kernel (dgemm_tma.cuh): one C, no K-slices; otherwise the cp.async kernel with split-K
    SplitPlan plan_gemm(int M, int N, int K, int max_split) const {
        if (tma_usable(M, N)) return SplitPlan{-2, 1, K};
        return plan_split_k(M, N, K, nsm, max_split);
    }
    int gemm_nt(int M, int N, int K, const double *A, long long lda, const double *B, long long ldb, double *C, long long ldc,
                const SplitPlan &pl, int kind, cudaStream_t stream) {
        if (pl.cfg == -2) {
            prof_begin(kind, stream);
            if (int e = launch_dgemm_tma_plain(M, N, K, A, lda, B, ldb, C, ldc, 1.0, tws[stream == st2 ? 1 : 0], nsm, stream)) return e;
            prof_end(stream, kind);
            ++launches;
            return 0;
        }
        GemmArgs g{};
        g.M = M; g.N = N; g.Kseg = pl.kseg; g.nseg = pl.nsplit; g.segs_per_split = 1; g.Ktot = K;
        g.A = A; g.lda = lda; g.a_seg_stride = pl.kseg; g.B = B; g.ldb = ldb; g.b_seg_stride = pl.kseg;
        g.C = C; g.ldc = ldc; g.c_split_stride = (long long)M * ldc; g.alpha = 1.0;
        prof_begin(kind, stream);
        SCLMD_CUDA(launch_dgemm(g, pl.nsplit, stream, pl.cfg));
        prof_end(stream, kind);
        ++launches;
        return 0;
    }
    int build_modal_tables() {
        ncs = 0;
        for (size_t i = 0; i < baths.size(); ++i) {
            moff[i] = ncs;
            ncs += baths[i]->ncp;
        }
        const size_t n = (size_t)ntraj * ld, nb = (size_t)ntraj * ncs;
        SCLMD_CUDA(mEL.alloc((size_t)ncs * ld));          // zero-initialised: pad rows / columns stay 0
        SCLMD_CUDA(mET.alloc((size_t)nph * ncs));
        for (size_t i = 0; i < baths.size(); ++i) {
            k_modal_tables<<<baths[i]->nc, 256, 0, st>>>(mU.p, mlam.p, nph, ld, baths[i]->cids.p, baths[i]->nc, moff[i], ncs, mEL.p, mET.p);
            SCLMD_CUDA(cudaGetLastError());
            ++launches;
        }
        splan = plan_gemm(ntraj, nph, ncs, 4);
        gaplan = plan_gemm(ntraj, ncs, ld, 16);
        if (const char *e = getenv("SCLMD_MODAL_SPLITS")) {      // "scatter,gather" K-split counts for A/B measurements
            int a = 0, b = 0;
            if (sscanf(e, "%d,%d", &a, &b) == 2 && a > 0 && b > 0 && splan.cfg != -2) {
                splan.kseg = round_up(cdiv(ncs, a), 16); splan.nsplit = cdiv(ncs, splan.kseg);
                gaplan.kseg = round_up(cdiv(ld, b), 16); gaplan.nsplit = cdiv(ld, gaplan.kseg);
            }
        }
        SCLMD_CUDA(mQ[0].alloc(n)); SCLMD_CUDA(mQ[1].alloc(n)); SCLMD_CUDA(mP.alloc(n));
        SCLMD_CUDA(mW.alloc(n * splan.nsplit));
        SCLMD_CUDA(mpc.alloc(nb)); SCLMD_CUDA(mfA.alloc(nb)); SCLMD_CUDA(mg.alloc(nb)); SCLMD_CUDA(msb.alloc(nb));
        SCLMD_CUDA(mgn.alloc(nb * gaplan.nsplit));
        SCLMD_CUDA(mec.alloc(ntraj));
        modal_tables = true;
        return 0;
    }
    int modal_bath(const BathSet &bs, int pending, int doA) {
        ModalArgs a;
        memset(&a, 0, sizeof(a));
        a.bs = bs;
        for (int i = 0; i < MAXB; ++i) a.off[i] = moff[i];
        a.ncs = ncs; a.ntraj = ntraj; a.nmd = nmd; a.pending = pending; a.doA = doA; a.gsplit = gaplan.nsplit;
        a.t = t; a.dt = dt;
        a.pc = mpc.p; a.fA = mfA.p; a.g = mg.p; a.sbuf = msb.p; a.ecorr = mec.p; a.gn = mgn.p;
        prof_begin(6);
        // two or three elements per thread: the dependent loads of an element (state, K.q slices, noise row, tail) are the whole cost
        static const int bath_thr = getenv("SCLMD_BATH_THREADS") ? atoi(getenv("SCLMD_BATH_THREADS")) : 256;      // 1024 CTAs in ONE wave (8 per SM); ncu: 16.3 us against 24.4 (320) and 18.8 (192)
        const int nthr = std::min(bath_thr, std::max(64, round_up(cdiv(ncs, 2), 32)));
        if (bs.nb <= 2) k_modal_bath<2><<<ntraj, nthr, 0, st>>>(a);
        else if (bs.nb <= 4) k_modal_bath<4><<<ntraj, nthr, 0, st>>>(a);
        else k_modal_bath<MAXB><<<ntraj, nthr, 0, st>>>(a);
        prof_end();
        SCLMD_CUDA(cudaGetLastError());
        ++launches;
        return 0;
    }
    int modal_pq(int pending, int advance, cudaStream_t stream) {
        prof_begin(7, stream);
        k_modal_pq<<<ntraj, 256, 0, stream>>>(nph, ld, ntraj, nmd, t, dt, pending, advance, mlam.p, mP.p, mQ[mqi].p, mQ[mqi ^ 1].p, mW.p, splan.nsplit,
                                              (size_t)ntraj * ld, mec.p, etot.p);
        prof_end(stream);
        SCLMD_CUDA(cudaGetLastError());
        ++launches;
        return 0;
    }
    int enter_modal() {       // real-space (q, p) -> (Q, Pi, p_c, g); the real-space path has been flushed by the caller
        if (!modal_tables) if (int e = build_modal_tables()) return e;
        const SplitPlan one = one_plan();
        mqi = 0;
        if (int e = gemm_nt(ntraj, nph, ld, q.p, ld, mUT.p, ld, mQ[0].p, ld, one, -1, st)) return e;       // Q = q . U
        if (int e = gemm_nt(ntraj, nph, ld, p.p, ld, mUT.p, ld, mP.p, ld, one, -1, st)) return e;          // Pi = p . U
        for (size_t i = 0; i < baths.size(); ++i) {
            k_modal_gather_pc<<<ntraj, 128, 0, st>>>(baths[i]->cids.p, baths[i]->nc, moff[i], ncs, ld, p.p, mpc.p);
            SCLMD_CUDA(cudaGetLastError());
            ++launches;
        }
        if (int e = gemm_nt(ntraj, ncs, ld, mQ[0].p, ld, mEL.p, ld, mgn.p, ncs, gaplan, -1, st)) return e; // g = (K q)[cids]
        const size_t nb = (size_t)ntraj * ncs;
        k_modal_fold<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(mgn.p, gaplan.nsplit, nb, mg.p);
        SCLMD_CUDA(cudaGetLastError());
        ++launches;
        m_pending = false;
        modal_state = true;
        return 0;
    }
    int modal_flush() {       // pending evaluations B, C of step t-1; afterwards mP holds Pi_t (no half kick)
        if (!m_pending) return 0;
        if (noise_pending) {
            SCLMD_CUDA(cudaStreamWaitEvent(st, evN, 0));
            noise_pending = false;
        }
        if (int e = modal_bath(view(), 1, 0)) return e;
        if (int e = gemm_nt(ntraj, nph, ncs, msb.p, ncs, mET.p, ncs, mW.p, ld, splan, -1, st)) return e;
        if (int e = modal_pq(1, 0, st)) return e;
        m_pending = false;
        return 0;
    }
    int leave_modal() {       // (Q, Pi) -> real-space (q, p)
        if (int e = modal_flush()) return e;
        const SplitPlan one = one_plan();
        if (int e = gemm_nt(ntraj, nph, ld, mQ[mqi].p, ld, mU.p, ld, q.p, ld, one, -1, st)) return e;      // q = Q . U^T
        if (int e = gemm_nt(ntraj, nph, ld, mP.p, ld, mU.p, ld, p.p, ld, one, -1, st)) return e;           // p = Pi . U^T
        modal_state = false;
        g_valid = false;
        d_valid = false;
        bc_pending = false;
        return 0;
    }
    int sync_real() {         // whatever mode the state is in: finish pending evaluations and make q, p current
        if (modal_state) return leave_modal();
        return flush();
    }
    int modal_step() {
        // every kernel of this step reads noise slab t only: an upload of other slabs (the streaming path sends slab t + 1 while
        // step t runs) must not hold the step back -- the wait then goes to the end of the step, ahead of the next one
        const bool defer_wait = noise_pending && !pending_upload_covers(t);
        if (noise_pending && !defer_wait) {
            SCLMD_CUDA(cudaStreamWaitEvent(st, evN, 0));
            noise_pending = false;
        }
        if (int e = modal_bath(view(), m_pending ? 1 : 0, 1)) return e;      // [B, C of step t-1] + A of step t; pushes p_t[cids]
        // second stream: scatter -> modal update (etot[t], Q_{t+1}) -> gather;   st: the history tails S'(t)
        // (sclmd_md_set_overlap(h, 0): everything on st, no cross-stream events -- A/B)
        cudaStream_t ms = overlap ? st2 : st;
        if (overlap) {
            SCLMD_CUDA(cudaEventRecord(evA, st));
            SCLMD_CUDA(cudaStreamWaitEvent(st2, evA, 0));
        }
        if (int e = gemm_nt(ntraj, nph, ncs, msb.p, ncs, mET.p, ncs, mW.p, ld, splan, 4, ms)) return e;
        if (int e = modal_pq(m_pending ? 1 : 0, 1, ms)) return e;
        SCLMD_CUDA(cudaEventRecord(evObs, ms));
        obs_slab = t % nmd;
        if (int e = gemm_nt(ntraj, ncs, ld, mQ[mqi ^ 1].p, ld, mEL.p, ld, mgn.p, ncs, gaplan, 5, ms)) return e;
        if (overlap) SCLMD_CUDA(cudaEventRecord(evG, st2));
        { for (auto &b : baths) if (int e = tail_step(*b, t)) return e; if (int e = flush_far()) return e; }
        if (overlap) SCLMD_CUDA(cudaStreamWaitEvent(st, evG, 0));
        if (defer_wait) {
            SCLMD_CUDA(cudaStreamWaitEvent(st, evN, 0));
            noise_pending = false;
        }
        mqi ^= 1;
        ++t;
        dt_synced = false;
        m_pending = true;
        return 0;
    }

    int step() {
        BathSet bs = view();
        const bool lin = any_lin();
        if (!g_valid) {
            if (int e = potforce(q.p, G.p)) return e;
            g_valid = true;
            d_valid = false;
        }
        if (lin)
            for (auto &b : baths)
                if (b->has_lin) if (int e = bath_lin(*b, p.p, q.p)) return e;
        if (bc_pending) {     // evaluations B, C of step t-1 fused with evaluation A of step t; G = K.q' of step t-1 = K.q_t
            if (noise_pending) {
                SCLMD_CUDA(cudaStreamWaitEvent(st, evN, 0));
                noise_pending = false;
            }
            if (bs.nb <= 2) k_phase_bca<2><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t - 1, dt, phalf.p, G.p, gplan.nsplit, (size_t)ntraj * ld, qn.p, etot.p);
            else if (bs.nb <= 4) k_phase_bca<4><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t - 1, dt, phalf.p, G.p, gplan.nsplit, (size_t)ntraj * ld, qn.p, etot.p);
            else k_phase_bca<MAXB><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t - 1, dt, phalf.p, G.p, gplan.nsplit, (size_t)ntraj * ld, qn.p, etot.p);
            SCLMD_CUDA(cudaEventRecord(evObs, st));
            obs_slab = t % nmd;
            SCLMD_CUDA(cudaGetLastError());
            ++launches;
            bc_pending = false;
        } else {
            const double *dc = d_valid ? Dc.p : nullptr;
            if (bs.nb <= 2) k_phase_a<2><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t, dt, q.p, p.p, G.p, gplan.nsplit, (size_t)ntraj * ld, dc, phalf.p, qn.p, etot.p);
            else if (bs.nb <= 4) k_phase_a<4><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t, dt, q.p, p.p, G.p, gplan.nsplit, (size_t)ntraj * ld, dc, phalf.p, qn.p, etot.p);
            else k_phase_a<MAXB><<<ntraj, phase_threads(), 0, st>>>(bs, nph, ld, ntraj, nmd, t, dt, q.p, p.p, G.p, gplan.nsplit, (size_t)ntraj * ld, dc, phalf.p, qn.p, etot.p);
            SCLMD_CUDA(cudaEventRecord(evObs, st));      // evaluation A wrote etot / currents of this slab
            obs_slab = t % nmd;
            SCLMD_CUDA(cudaGetLastError());
            ++launches;
        }
        d_valid = false;
        bool any_tail = false;
        for (auto &b : baths) any_tail |= b->ml > 1;
        if (corr() && !kc_valid) {
            k_gather_kc<<<nph, 128, 0, st>>>(K.p, nph, ld, cidx.p, ncons, ldcons, Kc.p);
            SCLMD_CUDA(cudaGetLastError());
            kc_valid = true;
            ++launches;
        }
        if (use_graphs && !any_tail && !profiling) {
            // ---- graph path: K.q' + B + C + correction replayed as one launch
            if (noise_pending) {
                SCLMD_CUDA(cudaStreamWaitEvent(st, evN, 0));
                noise_pending = false;
            }
            if (!dt_synced) {
                k_set_t<<<1, 1, 0, st>>>(d_t.p, t);
                SCLMD_CUDA(cudaGetLastError());
                dt_synced = true;
            }
            if (!gexec[gpar]) {
                const int64_t l0 = launches;
                SCLMD_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                int e = potforce(qn.p, Gn.p);
                if (!e) e = enqueue_rest(bs, lin, d_t.p);
                if (!e) {
                    k_advance_t<<<1, 1, 0, st>>>(d_t.p);
                    if (cudaGetLastError() != cudaSuccess) e = SCLMD_ERR_CUDA;
                }
                cudaGraph_t graph = nullptr;
                const cudaError_t ce = cudaStreamEndCapture(st, &graph);
                if (e || ce != cudaSuccess) {
                    if (graph) cudaGraphDestroy(graph);
                    if (!e) set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
                    return e ? e : SCLMD_ERR_CUDA;
                }
                const cudaError_t ie = cudaGraphInstantiate(&gexec[gpar], graph, 0);
                cudaGraphDestroy(graph);
                if (ie != cudaSuccess) {
                    gexec[gpar] = nullptr;
                    set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
                    return SCLMD_ERR_CUDA;
                }
                gnodes[gpar] = (int)(launches - l0);
                launches = l0;
            }
            SCLMD_CUDA(cudaGraphLaunch(gexec[gpar], st));
            launches += gnodes[gpar];
            finish_step();
            return 0;
        }
        dt_synced = false;
        if (overlap && any_tail) {   // GEMM on st2 first (1 CTA/SM leaves room for the tail CTAs), tails on st
            SCLMD_CUDA(cudaEventRecord(evA, st));
            SCLMD_CUDA(cudaStreamWaitEvent(st2, evA, 0));
            if (int e = potforce(qn.p, Gn.p, st2)) return e;
            SCLMD_CUDA(cudaEventRecord(evG, st2));
            { for (auto &b : baths) if (int e = tail_step(*b, t)) return e; if (int e = flush_far()) return e; }
            SCLMD_CUDA(cudaStreamWaitEvent(st, evG, 0));
        } else {
            { for (auto &b : baths) if (int e = tail_step(*b, t)) return e; if (int e = flush_far()) return e; }
            if (int e = potforce(qn.p, Gn.p)) return e;
        }
        if (lazy_ok(lin)) {       // B, C stay pending: fused with the next step's evaluation A, or run by flush()
            finish_step();
            bc_pending = true;
            return 0;
        }
        if (noise_pending) {   // rows streamed by sclmd_md_set_noise_rows: evaluations B, C read slab t+1
            SCLMD_CUDA(cudaStreamWaitEvent(st, evN, 0));
            noise_pending = false;
        }
        if (int e = enqueue_rest(bs, lin, nullptr)) return e;
        finish_step();
        return 0;
    }
};

// ---------------------------------------------------------------- C ABI
extern "C" {

int sclmd_md_create(int nph, int ntraj, double dt, int nmd, int device, sclmd_md **out) {
    SCLMD_REQUIRE(out != nullptr, "sclmd_md_create: out is NULL");
    *out = nullptr;
    SCLMD_REQUIRE(nph > 0 && ntraj > 0 && nmd > 0 && dt > 0, "sclmd_md_create: nph, ntraj, nmd, dt must be positive");
    if (int e = select_device(device)) return e;
    std::unique_ptr<sclmd_md> h(new sclmd_md());
    h->nph = nph; h->ld = round_up(nph, 2); h->ntraj = ntraj; h->nmd = nmd; h->dt = dt; h->device = device;
    h->nsm = sm_count(device);
    SCLMD_CUDA(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    SCLMD_CUDA(cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking));
    SCLMD_CUDA(cudaStreamCreateWithFlags(&h->stc, cudaStreamNonBlocking));
    SCLMD_CUDA(cudaEventCreateWithFlags(&h->evN, cudaEventDisableTiming));
    SCLMD_CUDA(cudaEventCreateWithFlags(&h->evObs, cudaEventDisableTiming));
    SCLMD_CUDA(cudaStreamCreateWithFlags(&h->str, cudaStreamNonBlocking));
    SCLMD_CUDA(cudaEventCreateWithFlags(&h->evA, cudaEventDisableTiming));
    SCLMD_CUDA(cudaEventCreateWithFlags(&h->evG, cudaEventDisableTiming));
    SCLMD_CUDA(cudaEventCreate(&h->ev0));
    SCLMD_CUDA(cudaEventCreate(&h->ev1));
    const size_t n = (size_t)ntraj * h->ld;
    SCLMD_CUDA(h->K.alloc((size_t)nph * h->ld));
    h->gplan = h->plan_gemm(ntraj, nph, h->ld, 4);
    SCLMD_CUDA(h->q.alloc(n)); SCLMD_CUDA(h->p.alloc(n));
    SCLMD_CUDA(h->G.alloc(n * h->gplan.nsplit)); SCLMD_CUDA(h->Gn.alloc(n * h->gplan.nsplit));
    SCLMD_CUDA(h->d_t.alloc(1));
    h->use_graphs = getenv("SCLMD_NO_GRAPH") == nullptr;
    h->use_persist = getenv("SCLMD_NO_PERSIST") == nullptr;
    h->fuse_bca = getenv("SCLMD_NO_FUSE") == nullptr;
    h->use_ens = getenv("SCLMD_NO_ENS") == nullptr;
    h->far_mma = getenv("SCLMD_NO_FAR_MMA") == nullptr;
    SCLMD_CUDA(h->phalf.alloc(n)); SCLMD_CUDA(h->p1.alloc(n)); SCLMD_CUDA(h->qn.alloc(n));
    SCLMD_CUDA(h->etot.alloc((size_t)nmd * ntraj));
    SCLMD_CUDA(h->cons.alloc(nph));
    SCLMD_CUDA(h->scratch.alloc(ntraj));
    *out = h.release();
    return SCLMD_OK;
}

int sclmd_md_destroy(sclmd_md *h) {
    if (!h) return SCLMD_OK;
    cudaSetDevice(h->device);
    if (h->st) cudaStreamSynchronize(h->st);
    h->drop_graphs();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->evA) cudaEventDestroy(h->evA);
    if (h->evG) cudaEventDestroy(h->evG);
    if (h->evN) cudaEventDestroy(h->evN);
    if (h->evObs) cudaEventDestroy(h->evObs);
    if (h->str) cudaStreamDestroy(h->str);
    if (h->stc) cudaStreamDestroy(h->stc);
    if (h->st2) cudaStreamDestroy(h->st2);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
    return SCLMD_OK;
}

int sclmd_md_set_dyn(sclmd_md *h, const double *K) {
    if (h) h->drop_graphs();      // captured kernel arguments are about to change
    SCLMD_REQUIRE(h && K, "sclmd_md_set_dyn: NULL argument");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->sync_real()) return e;
    SCLMD_CUDA(cudaMemcpy2DAsync(h->K.p, h->ld * sizeof(double), K, h->nph * sizeof(double), h->nph * sizeof(double), h->nph,
                                 cudaMemcpyHostToDevice, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    h->have_dyn = true;
    h->have_modes = false;       // the eigen-decomposition belongs to the previous matrix: sclmd_md_set_modes again
    h->g_valid = false;
    h->kc_valid = false;
    h->kfrag_valid = false;
    h->d_valid = false;
    return SCLMD_OK;
}

int sclmd_md_set_constraint(sclmd_md *h, const int32_t *idx, int n) {
    if (h) h->drop_graphs();      // captured kernel arguments are about to change
    SCLMD_REQUIRE(h && (n == 0 || idx), "sclmd_md_set_constraint: NULL argument");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->sync_real()) return e;
    std::vector<unsigned char> m(h->nph, 0);
    for (int i = 0; i < n; ++i) {
        SCLMD_REQUIRE(idx[i] >= 0 && idx[i] < h->nph, "sclmd_md_set_constraint: index %d out of range", idx[i]);
        m[idx[i]] = 1;
    }
    SCLMD_CUDA(cudaMemcpy(h->cons.p, m.data(), h->nph, cudaMemcpyHostToDevice));
    h->has_cons = n > 0;
    std::vector<int> uniq;
    for (int i = 0; i < h->nph; ++i) if (m[i]) uniq.push_back(i);
    h->ncons = (int)uniq.size();
    h->ldcons = round_up(std::max(h->ncons, 2), 2);
    h->use_corr = h->has_cons && 2 * h->ncons <= h->nph;   // cheaper than a second full K.q
    h->kc_valid = false;
    h->kfrag_valid = false;
    h->d_valid = false;
    h->g_valid = false;
    if (h->use_corr) {
        SCLMD_CUDA(h->cidx.alloc(h->ncons));
        SCLMD_CUDA(cudaMemcpy(h->cidx.p, uniq.data(), h->ncons * sizeof(int), cudaMemcpyHostToDevice));
        SCLMD_CUDA(h->Kc.alloc((size_t)h->nph * h->ldcons));
        SCLMD_CUDA(h->qc.alloc((size_t)h->ntraj * h->ldcons));
        SCLMD_CUDA(h->Dc.alloc((size_t)h->ntraj * h->ld));
        h->cplan = plan_split_k(h->ntraj, h->nph, h->ldcons, h->nsm, 1);
    }
    return SCLMD_OK;
}

int sclmd_md_add_bath(sclmd_md *h, const int32_t *cids, int nc, int ml, const double *kernel, int kernel_kind,
                      const double *Mq, const double *Mp, int *bath_out) {
    if (h) h->drop_graphs();      // captured kernel arguments are about to change
    SCLMD_REQUIRE(h && cids && kernel, "sclmd_md_add_bath: NULL argument");
    SCLMD_REQUIRE(nc > 0 && ml >= 1, "sclmd_md_add_bath: nc and ml must be positive");
    SCLMD_REQUIRE(kernel_kind == SCLMD_KERNEL_FULL || kernel_kind == SCLMD_KERNEL_DIAG, "sclmd_md_add_bath: bad kernel_kind");
    SCLMD_REQUIRE((int)h->baths.size() < MAXB, "sclmd_md_add_bath: at most %d baths", MAXB);
    SCLMD_REQUIRE(!(Mq || Mp) || ml == 1, "sclmd_md_add_bath: Mq/Mp only act for time-local baths (baths.py:243-249)");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->sync_real()) return e;
    std::unique_ptr<Bath> b(new Bath());
    b->nc = nc; b->ncp = round_up(nc, 2); b->ml = ml; b->kind = kernel_kind;
    b->c0 = ml > 1 ? h->dt : 1.0;  // baths.py:454-457: the dt factor only exists for ml > 1
    b->has_extra = (Mq || Mp);
    b->has_lin = kernel_kind == SCLMD_KERNEL_FULL || b->has_extra;
    const int ncp = b->ncp, ntraj = h->ntraj;
    std::vector<int> inv(h->nph, -1);
    for (int a = 0; a < nc; ++a) {
        SCLMD_REQUIRE(cids[a] >= 0 && cids[a] < h->nph, "sclmd_md_add_bath: dof index %d out of range", cids[a]);
        SCLMD_REQUIRE(inv[cids[a]] < 0, "sclmd_md_add_bath: duplicate dof index %d", cids[a]);
        inv[cids[a]] = a;
    }
    b->cids_h.assign(cids, cids + nc);
    SCLMD_CUDA(b->cids.alloc(nc)); SCLMD_CUDA(b->inv.alloc(h->nph));
    SCLMD_CUDA(cudaMemcpy(b->cids.p, cids, nc * sizeof(int), cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy(b->inv.p, inv.data(), h->nph * sizeof(int), cudaMemcpyHostToDevice));
    // kernel
    if (kernel_kind == SCLMD_KERNEL_DIAG) {
        SCLMD_CUDA(b->kern.alloc((size_t)(ml + 8 * TB + 2) * ncp));   // rows >= ml stay 0 (branch-free drop-out / far-tail padding)
        SCLMD_CUDA(cudaMemcpy2D(b->kern.p, ncp * sizeof(double), kernel, nc * sizeof(double), nc * sizeof(double), ml, cudaMemcpyHostToDevice));
    } else {
        SCLMD_CUDA(b->kern.alloc((size_t)ml * nc * ncp));
        SCLMD_CUDA(cudaMemcpy2D(b->kern.p, ncp * sizeof(double), kernel, nc * sizeof(double), nc * sizeof(double), (size_t)ml * nc, cudaMemcpyHostToDevice));
    }
    if (b->has_lin) {  // W = [ -c0*K0 + Mp | Mq ]  (K0 = 0 here for diagonal kernels: handled elementwise)
        b->Kw = b->has_extra ? 2 * ncp : ncp;
        std::vector<double> W((size_t)nc * b->Kw, 0.0);
        for (int a = 0; a < nc; ++a)
            for (int c = 0; c < nc; ++c) {
                double v = 0.0;
                if (kernel_kind == SCLMD_KERNEL_FULL) v -= b->c0 * kernel[(size_t)a * nc + c];
                if (Mp) v += Mp[(size_t)a * nc + c];
                W[(size_t)a * b->Kw + c] = v;
                if (b->has_extra && Mq) W[(size_t)a * b->Kw + ncp + c] = Mq[(size_t)a * nc + c];
            }
        SCLMD_CUDA(b->W.alloc(W.size()));
        SCLMD_CUDA(cudaMemcpy(b->W.p, W.data(), W.size() * sizeof(double), cudaMemcpyHostToDevice));
        std::vector<double> WT((size_t)b->Kw * ncp, 0.0);
        for (int a = 0; a < nc; ++a)
            for (int k = 0; k < b->Kw; ++k) WT[(size_t)k * ncp + a] = W[(size_t)a * b->Kw + k];
        SCLMD_CUDA(b->WT.alloc(WT.size()));
        SCLMD_CUDA(cudaMemcpy(b->WT.p, WT.data(), WT.size() * sizeof(double), cudaMemcpyHostToDevice));
        SCLMD_CUDA(b->xq.alloc((size_t)ntraj * b->Kw));
        SCLMD_CUDA(b->lin.alloc((size_t)ntraj * ncp));
    }
    SCLMD_CUDA(b->ring.alloc((size_t)ntraj * ml * ncp));
    if (ml > 1) {
        // enough CTAs to fill the machine a few times over, deterministic partial sums
        int tiles = kernel_kind == SCLMD_KERNEL_DIAG ? cdiv(ntraj, ntraj >= 4 ? 4 : 1) * cdiv(ncp / 2, 256)
                                                      : cdiv(ntraj, 128) * cdiv(nc, 128);
        int want = kernel_kind == SCLMD_KERNEL_DIAG ? 4 * h->nsm : 2 * h->nsm;
        b->nsplit = std::max(1, std::min({cdiv(want, tiles), 32, std::max(1, (ml - 1) / 64)}));
        if (kernel_kind == SCLMD_KERNEL_FULL && tma_usable(ntraj, nc)) {
            b->gemm_cfg = -2;       // the ring-segment contraction runs on the TMA / stream-K kernel: one tail, no K-slices
            b->nsplit = 1;
        } else if (kernel_kind == SCLMD_KERNEL_FULL && ntraj > 64) {
            // ring-segment GEMM [ntraj x nc] with a very deep K: 64-wide tiles when 128-wide ones would be mostly padding, and a
            // split count that fills whole waves (6 tiles x 32 splits on 148 SMs ran two waves at 65 %)
            const bool narrow = round_up(nc, 128) - nc >= 64;
            b->gemm_cfg = narrow ? 3 : 0;
            tiles = cdiv(ntraj, 128) * cdiv(nc, narrow ? 64 : 128);
            b->nsplit = wave_fit_splits(tiles, h->nsm * (narrow ? 2 : 1), std::max(1, std::min(64, (ml - 1) / 64)));
        }
        SCLMD_CUDA(b->tailp.alloc((size_t)b->nsplit * ntraj * ncp));
        if (kernel_kind == SCLMD_KERNEL_DIAG && ml >= 8 * TB) {
            b->blocked = true;
            const int ctas = cdiv(ntraj, ntraj >= 4 ? 4 : 1) * cdiv(ncp, 256);
            b->far_nsplit = std::max(1, std::min({cdiv(4 * h->nsm, ctas), 16, ml / (4 * TB)}));
            b->far_cap = std::max(b->far_nsplit, 4);                                    // the tensor-pipe pass may use up to four age splits
            SCLMD_CUDA(b->far.alloc((size_t)b->far_cap * 2 * TB * ntraj * ncp));        // up to 32 far tails per pass and split
        }
    }
    SCLMD_CUDA(b->noise.alloc((size_t)h->nmd * ntraj * ncp));
    SCLMD_CUDA(b->cur.alloc((size_t)h->nmd * ntraj));
    if (bath_out) *bath_out = (int)h->baths.size();
    h->modal_tables = false;
    h->baths.push_back(std::move(b));
    return SCLMD_OK;
}

static int check_bath(sclmd_md *h, int bath, const char *fn) {
    SCLMD_REQUIRE(h, "%s: NULL handle", fn);
    SCLMD_REQUIRE(bath >= 0 && bath < (int)h->baths.size(), "%s: bath index %d out of range", fn, bath);
    return 0;
}

int sclmd_md_set_noise(sclmd_md *h, int bath, int traj0, int nsel, const double *noise) {
    if (int e = check_bath(h, bath, "sclmd_md_set_noise")) return e;
    SCLMD_REQUIRE(noise && traj0 >= 0 && nsel > 0 && traj0 + nsel <= h->ntraj, "sclmd_md_set_noise: bad trajectory range");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->flush()) return e;
    Bath &b = *h->baths[bath];
    // host [traj][nmd][nc] -> device [traj][nmd][ncp]: one pitched copy
    SCLMD_CUDA(cudaMemcpy2DAsync(b.noise.p + (size_t)traj0 * h->nmd * b.ncp, b.ncp * sizeof(double), noise, b.nc * sizeof(double),
                                 b.nc * sizeof(double), (size_t)nsel * h->nmd, cudaMemcpyHostToDevice, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    return SCLMD_OK;
}

int sclmd_md_get_noise(sclmd_md *h, int bath, int traj0, int nsel, double *noise) {
    if (int e = check_bath(h, bath, "sclmd_md_get_noise")) return e;
    SCLMD_REQUIRE(noise && traj0 >= 0 && nsel > 0 && traj0 + nsel <= h->ntraj, "sclmd_md_get_noise: bad trajectory range");
    SCLMD_CUDA(cudaSetDevice(h->device));
    Bath &b = *h->baths[bath];
    SCLMD_CUDA(cudaMemcpy2DAsync(noise, b.nc * sizeof(double), b.noise.p + (size_t)traj0 * h->nmd * b.ncp, b.ncp * sizeof(double),
                                 b.nc * sizeof(double), (size_t)nsel * h->nmd, cudaMemcpyDeviceToHost, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    return SCLMD_OK;
}

int sclmd_md_set_state(sclmd_md *h, const double *q, const double *p, int64_t t) {
    SCLMD_REQUIRE(h, "sclmd_md_set_state: NULL handle");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->sync_real()) return e;
    const size_t w = h->nph * sizeof(double), pitch = h->ld * sizeof(double);
    if (q) SCLMD_CUDA(cudaMemcpy2DAsync(h->q.p, pitch, q, w, w, h->ntraj, cudaMemcpyHostToDevice, h->st));
    if (p) SCLMD_CUDA(cudaMemcpy2DAsync(h->p.p, pitch, p, w, w, h->ntraj, cudaMemcpyHostToDevice, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    if (t >= 0 && t != h->t) {
        h->t = t;
        h->dt_synced = false;
        for (auto &b : h->baths) b->far_t0 = -1;
    }
    if (q) { h->g_valid = false; h->d_valid = false; }
    return SCLMD_OK;
}

int sclmd_md_get_state(sclmd_md *h, double *q, double *p, int64_t *t) {
    SCLMD_REQUIRE(h, "sclmd_md_get_state: NULL handle");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->sync_real()) return e;
    const size_t w = h->nph * sizeof(double), pitch = h->ld * sizeof(double);
    if (q) SCLMD_CUDA(cudaMemcpy2DAsync(q, w, h->q.p, pitch, w, h->ntraj, cudaMemcpyDeviceToHost, h->st));
    if (p) SCLMD_CUDA(cudaMemcpy2DAsync(p, w, h->p.p, pitch, w, h->ntraj, cudaMemcpyDeviceToHost, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    if (t) *t = h->t;
    return SCLMD_OK;
}

int sclmd_md_reset_history(sclmd_md *h) {
    SCLMD_REQUIRE(h, "sclmd_md_reset_history: NULL handle");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->flush()) return e;
    for (auto &b : h->baths) {
        SCLMD_CUDA(cudaMemsetAsync(b->ring.p, 0, b->ring.n * sizeof(double), h->st));
        if (b->tailp.p) SCLMD_CUDA(cudaMemsetAsync(b->tailp.p, 0, b->tailp.n * sizeof(double), h->st));
        b->far_t0 = -1;
    }
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    return SCLMD_OK;
}

// reference order: phis[i] = p_{t-1-i}[cids]  (md.py:387; row 0 newest) <-> ring slot (t-1-i) mod ml
int sclmd_md_get_history(sclmd_md *h, int bath, double *phis) {
    if (int e = check_bath(h, bath, "sclmd_md_get_history")) return e;
    SCLMD_REQUIRE(phis, "sclmd_md_get_history: NULL buffer");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->flush()) return e;
    Bath &b = *h->baths[bath];
    std::vector<double> tmp(b.ring.n);
    SCLMD_CUDA(cudaMemcpyAsync(tmp.data(), b.ring.p, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    for (int tr = 0; tr < h->ntraj; ++tr)
        for (int i = 0; i < b.ml; ++i) {
            long long s = (h->t - 1 - i) % b.ml;
            if (s < 0) s += b.ml;
            memcpy(phis + ((size_t)tr * b.ml + i) * b.nc, tmp.data() + ((size_t)tr * b.ml + s) * b.ncp, b.nc * sizeof(double));
        }
    return SCLMD_OK;
}

int sclmd_md_set_history(sclmd_md *h, int bath, const double *phis) {
    if (int e = check_bath(h, bath, "sclmd_md_set_history")) return e;
    SCLMD_REQUIRE(phis, "sclmd_md_set_history: NULL buffer");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->flush()) return e;
    Bath &b = *h->baths[bath];
    std::vector<double> tmp(b.ring.n, 0.0);
    for (int tr = 0; tr < h->ntraj; ++tr)
        for (int i = 0; i < b.ml; ++i) {
            long long s = (h->t - 1 - i) % b.ml;
            if (s < 0) s += b.ml;
            memcpy(tmp.data() + ((size_t)tr * b.ml + s) * b.ncp, phis + ((size_t)tr * b.ml + i) * b.nc, b.nc * sizeof(double));
        }
    SCLMD_CUDA(cudaMemcpyAsync(b.ring.p, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));      // `tmp` is pageable: the copy has landed before the tail kernel below is queued
    // the tail for evaluation A of the next step: ring as if p_{t-1} had just been pushed
    if (b.ml > 1) {
        b.far_t0 = -1;
        if (int e = h->tail_step(b, h->t - 1)) return e;
    if (int e = h->flush_far()) return e;
        if (int e = h->flush_far()) return e;
        SCLMD_CUDA(cudaStreamSynchronize(h->st));
    }
    return SCLMD_OK;
}

// 1 (default): diagonal-kernel baths with ml >= 128 stream their ring once per 16 steps (time-blocked far/near tails);
// 0: one full ring pass per step (the direct single-tail algorithm the HBM roofline of SURVEY 8d is defined on)
int sclmd_md_set_tail_block(sclmd_md *h, int on) {
    SCLMD_REQUIRE(h, "sclmd_md_set_tail_block: NULL handle");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->flush()) return e;
    h->tail_block = on != 0;
    h->far_tma = on != 2;    // 2 = time-blocked with the plain-load far kernel (for A/B measurements)
    h->far_ws = on != 3;     // 3 = time-blocked with the two-stage TMA kernel (no producer warp)
    h->far_mma = on == 1;    // 1 = the default: tensor-pipe far pass over 32-step blocks; 5 = the 16-step DFMA kernel with the producer warp (previous default)
    h->far32 = on == 4;      // 4 = 32-step blocks (k_tail_far_wsx, one trajectory per CTA): half the ring traffic, but FMA-issue bound
    for (auto &b : h->baths) {   // the partial-tail layout differs between the modes: rebuild S'(t-1)
        b->far_t0 = -1;
        if (b->ml > 1) if (int e = h->tail_step(*b, h->t - 1)) return e;
        if (int e = h->flush_far()) return e;
    }
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    return SCLMD_OK;
}

int sclmd_md_run(sclmd_md *h, int64_t nsteps, float *elapsed_ms) {
    SCLMD_REQUIRE(h && nsteps >= 0, "sclmd_md_run: bad arguments");
    if (h->ext_force) {
        set_error("sclmd_md_run: the handle takes its potential force from the host (sclmd_md_set_external_force): "
                  "step with sclmd_md_step_begin / sclmd_md_step_end");
        return SCLMD_ERR_STATE;
    }
    if (!h->have_dyn) {
        set_error("sclmd_md_run: no dynamical matrix set (md.py:469 'no driver, no md')");
        return SCLMD_ERR_STATE;
    }
    SCLMD_CUDA(cudaSetDevice(h->device));
    SCLMD_CUDA(cudaEventRecord(h->ev0, h->st));
    bool done = false;
    const bool ens = h->ens_ok(), persist = h->persist_ok();
    const bool modal = nsteps > 0 && !ens && !persist && h->modal_ok();
    if (h->modal_state && !modal && nsteps > 0)
        if (int e = h->sync_real()) return e;
    if (nsteps >= 8 && ens) {             // short runs (the per-step end-to-end path) stay on the launch chain
        if (int e = h->flush()) return e;
        const int e = h->run_ens(nsteps);
        if (e < 0) return e;
        done = e == 0;          // 1: the state does not fit shared memory, take the launch chain
    }
    if (done) {
    } else if (nsteps > 0 && persist) {
        if (int e = h->flush()) return e;
        if (int e = h->run_persist(nsteps)) return e;
    } else if (modal) {                   // eigenbasis propagation: gather + scatter products instead of K.q
        if (!h->modal_state) {
            if (int e = h->flush()) return e;
            if (int e = h->enter_modal()) return e;
        }
        for (int64_t s = 0; s < nsteps; ++s)
            if (int e = h->modal_step()) return e;
    } else {
        for (int64_t s = 0; s < nsteps; ++s)
            if (int e = h->step()) return e;
    }
    if (elapsed_ms || h->profiling)                     // a timed run includes its last evaluations B, C
        if (int e = h->flush()) return e;
    SCLMD_CUDA(cudaEventRecord(h->ev1, h->st));
    if (!elapsed_ms && !h->profiling) return SCLMD_OK;   // asynchronous: the next sclmd_md_get_* synchronises
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    if (elapsed_ms) SCLMD_CUDA(cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
    h->prof_collect();
    return SCLMD_OK;
}

// ---- force drivers (md.AddPotential, md.py:457-459, 481-485): potential force from a host callback
// on = 1: the handle never multiplies by K; the caller supplies f(q) (reference sign: the force itself, mass-weighted,
// [ntraj][nph]) through sclmd_md_set_force / sclmd_md_step_end.  Per step:
//     if (sclmd_md_force_needed(h)) sclmd_md_set_force(h, f(q_t));         // first step, after set_state, or with constraints
//     sclmd_md_step_begin(h, q_trial);  sclmd_md_step_end(h, f(q_trial));
int sclmd_md_set_external_force(sclmd_md *h, int on) {
    SCLMD_REQUIRE(h, "sclmd_md_set_external_force: NULL handle");
    SCLMD_REQUIRE(!h->ext_open, "sclmd_md_set_external_force: a step is open (sclmd_md_step_end missing)");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->sync_real()) return e;
    h->drop_graphs();
    h->ext_force = on != 0;
    h->g_valid = false;
    h->d_valid = false;
    return SCLMD_OK;
}

int sclmd_md_force_needed(sclmd_md *h) {
    SCLMD_REQUIRE(h, "sclmd_md_force_needed: NULL handle");
    return (h->ext_force && !h->g_valid) ? 1 : 0;
}

int sclmd_md_set_force(sclmd_md *h, const double *f) {
    SCLMD_REQUIRE(h && f, "sclmd_md_set_force: NULL argument");
    SCLMD_REQUIRE(h->ext_force, "sclmd_md_set_force: the handle is not in external-force mode");
    SCLMD_REQUIRE(!h->ext_open, "sclmd_md_set_force: a step is open (sclmd_md_step_end missing)");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->upload_force(h->G.p, f)) return e;
    SCLMD_CUDA(cudaStreamSynchronize(h->st));     // `f` may be reused by the caller
    h->g_valid = true;
    h->d_valid = false;
    return SCLMD_OK;
}

int sclmd_md_step_begin(sclmd_md *h, double *q_trial) {
    SCLMD_REQUIRE(h && q_trial, "sclmd_md_step_begin: NULL argument");
    SCLMD_REQUIRE(h->ext_force, "sclmd_md_step_begin: the handle is not in external-force mode");
    SCLMD_REQUIRE(!h->ext_open, "sclmd_md_step_begin: the previous step is still open (sclmd_md_step_end missing)");
    if (!h->g_valid) {
        set_error("sclmd_md_step_begin: the force at the current positions is missing (sclmd_md_set_force)");
        return SCLMD_ERR_STATE;
    }
    SCLMD_CUDA(cudaSetDevice(h->device));
    return h->ext_begin(q_trial);
}

int sclmd_md_step_end(sclmd_md *h, const double *f_trial) {
    SCLMD_REQUIRE(h && f_trial, "sclmd_md_step_end: NULL argument");
    SCLMD_REQUIRE(h->ext_force && h->ext_open, "sclmd_md_step_end: no open step (sclmd_md_step_begin first)");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->ext_end(f_trial)) return e;
    SCLMD_CUDA(cudaStreamSynchronize(h->st));     // `f_trial` may be reused by the caller
    return SCLMD_OK;
}

// md.f, md.fbaths and md.fhis (md.py:390-398, 403, 411): on = 1 makes every step also store the total force of evaluation C
// (md.f), each bath's force of evaluation C (md.fbaths after vv) and of evaluation A (the one the heat current is built from,
// md.fhis); read them with the two getters after a step.
int sclmd_md_set_force_output(sclmd_md *h, int on) {
    SCLMD_REQUIRE(h, "sclmd_md_set_force_output: NULL handle");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->sync_real()) return e;
    h->drop_graphs();
    if (on) {
        if (!h->fC.p) SCLMD_CUDA(h->fC.alloc((size_t)h->ntraj * h->ld));
        for (auto &b : h->baths)
            if (!b->fa.p) {
                SCLMD_CUDA(b->fa.alloc((size_t)h->ntraj * b->ncp));
                SCLMD_CUDA(b->fc.alloc((size_t)h->ntraj * b->ncp));
            }
    }
    h->want_f = on != 0;
    h->f_step = -1;
    return SCLMD_OK;
}

int sclmd_md_get_force(sclmd_md *h, double *f) {
    SCLMD_REQUIRE(h && f, "sclmd_md_get_force: NULL argument");
    SCLMD_REQUIRE(h->want_f && h->f_step == h->t, "sclmd_md_get_force: no step since sclmd_md_set_force_output(h, 1) / set_state");
    SCLMD_CUDA(cudaSetDevice(h->device));
    const size_t w = h->nph * sizeof(double);
    SCLMD_CUDA(cudaMemcpy2DAsync(f, w, h->fC.p, h->ld * sizeof(double), w, h->ntraj, cudaMemcpyDeviceToHost, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    return SCLMD_OK;
}

int sclmd_md_get_bath_force(sclmd_md *h, int bath, int evaluation, double *fb) {
    if (int e = check_bath(h, bath, "sclmd_md_get_bath_force")) return e;
    SCLMD_REQUIRE(fb && (evaluation == 0 || evaluation == 2), "sclmd_md_get_bath_force: NULL buffer or evaluation not 0 (A) / 2 (C)");
    SCLMD_REQUIRE(h->want_f && h->f_step == h->t && h->baths[bath]->fa.p,
                  "sclmd_md_get_bath_force: no step since sclmd_md_set_force_output(h, 1) / set_state");
    SCLMD_CUDA(cudaSetDevice(h->device));
    Bath &b = *h->baths[bath];
    const size_t w = b.nc * sizeof(double);
    SCLMD_CUDA(cudaMemcpy2DAsync(fb, w, evaluation == 0 ? b.fa.p : b.fc.p, b.ncp * sizeof(double), w, h->ntraj, cudaMemcpyDeviceToHost, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    return SCLMD_OK;
}

// 1 (default): run the K.q GEMM on a second stream concurrently with the history tails; 0: single stream
int sclmd_md_set_overlap(sclmd_md *h, int on) {
    SCLMD_REQUIRE(h, "sclmd_md_set_overlap: NULL handle");
    h->overlap = on != 0;
    return SCLMD_OK;
}

// 1 (default): runs of <= 2 trajectories of <= 1024 dofs with time-local diagonal baths use the persistent cooperative kernel
// (one launch per sclmd_md_run, one grid barrier per step); 0: always the per-step launch chain
int sclmd_md_set_persistent(sclmd_md *h, int on) {
    SCLMD_REQUIRE(h, "sclmd_md_set_persistent: NULL handle");
    h->use_persist = on != 0;
    return SCLMD_OK;
}

int sclmd_md_set_profiling(sclmd_md *h, int on) {
    SCLMD_REQUIRE(h, "sclmd_md_set_profiling: NULL handle");
    h->profiling = on != 0;
    for (int i = 0; i < 8; ++i) { h->prof_ms[i] = 0; h->prof_n[i] = 0; }
    return SCLMD_OK;
}

int sclmd_md_get_profile(sclmd_md *h, double *tail_ms, int64_t *tail_launches, double *potforce_ms, int64_t *potforce_launches) {
    SCLMD_REQUIRE(h, "sclmd_md_get_profile: NULL handle");
    if (tail_ms) *tail_ms = h->prof_ms[0];
    if (tail_launches) *tail_launches = h->prof_n[0];
    if (potforce_ms) *potforce_ms = h->prof_ms[1];
    if (potforce_launches) *potforce_launches = h->prof_n[1];
    return SCLMD_OK;
}

// ms[4], n[4]: 0 = direct history tail, 1 = potential force (when on the main stream), 2 = time-blocked far pass, 3 = near pass
int sclmd_md_get_profile_all(sclmd_md *h, double *ms, int64_t *n) {
    SCLMD_REQUIRE(h && ms && n, "sclmd_md_get_profile_all: NULL argument");
    for (int i = 0; i < 4; ++i) { ms[i] = h->prof_ms[i]; n[i] = h->prof_n[i]; }
    return SCLMD_OK;
}

// ms[8], n[8]: kinds 0-3 as above; modal mode: 4 = scatter product, 5 = gather product, 6 = bath-dof kernel, 7 = modal update
int sclmd_md_get_profile_ex(sclmd_md *h, double *ms, int64_t *n) {
    SCLMD_REQUIRE(h && ms && n, "sclmd_md_get_profile_ex: NULL argument");
    for (int i = 0; i < 8; ++i) { ms[i] = h->prof_ms[i]; n[i] = h->prof_n[i]; }
    return SCLMD_OK;
}

// md.setDyn keeps the eigen-decomposition it projects with (md.py:266-281: hw = sqrt(lam), U): handing it over lets the handle
// propagate in the eigenbasis (k_modal_bath) whenever that is cheaper and the problem allows it (no constraints, no force driver,
// diagonal-kernel baths on disjoint dofs).  lam[nph] (already clipped at 0), U[nph*nph] row-major with the eigenvectors as COLUMNS
// (numpy.linalg.eigh).  Checked against the matrix of sclmd_md_set_dyn: max |K U - U diag(lam)| <= 1e-9 max(lam).
int sclmd_md_set_modes(sclmd_md *h, const double *lam, const double *U) {
    SCLMD_REQUIRE(h && lam && U, "sclmd_md_set_modes: NULL argument");
    if (!h->have_dyn) {
        set_error("sclmd_md_set_modes: set the dynamical matrix first (sclmd_md_set_dyn)");
        return SCLMD_ERR_STATE;
    }
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->sync_real()) return e;
    const int nph = h->nph, ld = h->ld;
    h->have_modes = false;
    h->modal_tables = false;
    SCLMD_CUDA(h->mU.alloc((size_t)nph * ld)); SCLMD_CUDA(h->mUT.alloc((size_t)nph * ld)); SCLMD_CUDA(h->mlam.alloc(ld));
    SCLMD_CUDA(cudaMemcpy2DAsync(h->mU.p, ld * sizeof(double), U, nph * sizeof(double), nph * sizeof(double), nph, cudaMemcpyHostToDevice, h->st));
    SCLMD_CUDA(cudaMemcpyAsync(h->mlam.p, lam, nph * sizeof(double), cudaMemcpyHostToDevice, h->st));
    k_modal_transpose<<<dim3(cdiv(nph, 32), cdiv(nph, 32)), dim3(32, 8), 0, h->st>>>(h->mU.p, nph, ld, h->mUT.p);
    SCLMD_CUDA(cudaGetLastError());
    // residual of the decomposition against the matrix the real-space path multiplies with
    DevBuf<double> KU;
    DevBuf<unsigned long long> mx;
    SCLMD_CUDA(KU.alloc((size_t)nph * ld)); SCLMD_CUDA(mx.alloc(2));
    const SplitPlan one = tma_usable(nph, nph) ? SplitPlan{-2, 1, ld} : SplitPlan{nph > 64 ? 0 : (nph > 16 ? 1 : 2), 1, ld};
    if (int e = h->gemm_nt(nph, nph, ld, h->K.p, ld, h->mUT.p, ld, KU.p, ld, one, -1, h->st)) return e;
    k_modal_residual<<<nph, 128, 0, h->st>>>(KU.p, h->mU.p, h->mlam.p, nph, ld, mx.p);
    SCLMD_CUDA(cudaGetLastError());
    unsigned long long bits[2] = {0, 0};
    SCLMD_CUDA(cudaMemcpyAsync(bits, mx.p, sizeof(bits), cudaMemcpyDeviceToHost, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    h->launches += 3;
    double res, lmax;
    memcpy(&res, &bits[0], 8);
    memcpy(&lmax, &bits[1], 8);
    for (int k = 0; k < nph; ++k)
        SCLMD_REQUIRE(lam[k] >= 0.0, "sclmd_md_set_modes: negative eigenvalue %g (md.setDyn clips them to 0, md.py:268-274)", lam[k]);
    SCLMD_REQUIRE(res <= 1e-9 * std::max(lmax, 1e-300), "sclmd_md_set_modes: K U != U diag(lam): residual %.3e against max(lam) = %.3e", res, lmax);
    h->have_modes = true;
    return SCLMD_OK;
}

// 1 (default): use the eigenbasis propagation when modes are set and the problem allows it; 0: always real space
int sclmd_md_set_modal(sclmd_md *h, int on) {
    SCLMD_REQUIRE(h, "sclmd_md_set_modal: NULL handle");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->sync_real()) return e;
    h->modal_on = on != 0;
    return SCLMD_OK;
}
// 1 when the next sclmd_md_run would propagate in the eigenbasis
int sclmd_md_modal_active(sclmd_md *h) {
    SCLMD_REQUIRE(h, "sclmd_md_modal_active: NULL handle");
    return (!h->ens_ok() && !h->persist_ok() && h->modal_ok()) ? 1 : 0;
}

// rows[nslab][ntraj][nc] (time-major, as the device table) for time slabs [slab0, slab0+nslab) mod nmd
int sclmd_md_set_noise_rows(sclmd_md *h, int bath, int slab0, int nslab, const double *rows) {
    if (int e = check_bath(h, bath, "sclmd_md_set_noise_rows")) return e;
    SCLMD_REQUIRE(rows && nslab > 0 && nslab <= h->nmd && slab0 >= 0 && slab0 < h->nmd, "sclmd_md_set_noise_rows: bad slab range");
    SCLMD_CUDA(cudaSetDevice(h->device));
    Bath &b = *h->baths[bath];
    const size_t rowsz = (size_t)h->ntraj;
    if ((h->bc_pending || h->m_pending) && (int)(((h->t % h->nmd) - slab0 + h->nmd) % h->nmd) < nslab)     // pending evaluations B, C read slab t
        if (int e = h->flush()) return e;
    // ASYNCHRONOUS on the handle's copy stream: the upload overlaps the step that is running; the next
    // sclmd_md_run orders itself after it.  `rows` must stay valid until the next synchronising call
    // (sclmd_md_run with elapsed_ms != NULL, any sclmd_md_get_*).  Pinned host memory makes it a true DMA.
    // one linear DMA into a staging buffer, then a small kernel on the copy stream scatters the rows into the trajectory-major table
    const size_t cap_slabs = std::max<size_t>(1, ((size_t)64 << 20) / (rowsz * b.nc * sizeof(double)));
    int done = 0;
    while (done < nslab) {
        const int n = (int)std::min<size_t>(nslab - done, cap_slabs);
        const size_t cnt = (size_t)n * rowsz * b.nc;
        if (b.rowstage.n < cnt) SCLMD_CUDA(b.rowstage.alloc_raw(cnt));
        SCLMD_CUDA(cudaMemcpyAsync(b.rowstage.p, rows + (size_t)done * rowsz * b.nc, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stc));
        k_scatter_rows<<<dim3(h->ntraj, n), 128, 0, h->stc>>>(b.rowstage.p, b.noise.p, (slab0 + done) % h->nmd, h->ntraj, b.nc, b.ncp, h->nmd);
        SCLMD_CUDA(cudaGetLastError());
        ++h->launches;
        done += n;
    }
    SCLMD_CUDA(cudaEventRecord(h->evN, h->stc));
    if (h->noise_pending && (h->np_slab0 != slab0 || h->np_nslab != nslab)) h->np_mixed = true;
    if (!h->noise_pending) h->np_mixed = false;
    h->np_slab0 = slab0;
    h->np_nslab = nslab;
    h->noise_pending = true;
    return SCLMD_OK;
}

// observables recorded at time slab `slab` (= t % nmd of that step): out[(1+nbaths)][ntraj] = etot, cur_0, cur_1, ...
int sclmd_md_get_step_observables(sclmd_md *h, int slab, double *out) {
    SCLMD_REQUIRE(h && out && slab >= 0 && slab < h->nmd, "sclmd_md_get_step_observables: bad arguments");
    SCLMD_CUDA(cudaSetDevice(h->device));
    const size_t n = (size_t)h->ntraj;
    // The observables of a step are produced by its first kernel (evaluation A).  For the slab of the latest step the
    // read-back only waits for that kernel (event), on its own stream, so the rest of the step keeps running while the host
    // already prepares the next one; any other slab takes the fully synchronising path.
    const bool latest = (long long)slab == h->obs_slab;
    cudaStream_t s = latest ? h->str : h->st;
    if (latest) SCLMD_CUDA(cudaStreamWaitEvent(h->str, h->evObs, 0));
    SCLMD_CUDA(cudaMemcpyAsync(out, h->etot.p + (size_t)slab * n, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    for (size_t b = 0; b < h->baths.size(); ++b)
        SCLMD_CUDA(cudaMemcpyAsync(out + (b + 1) * n, h->baths[b]->cur.p + (size_t)slab * n, n * sizeof(double),
                                   cudaMemcpyDeviceToHost, s));
    SCLMD_CUDA(cudaStreamSynchronize(s));
    SCLMD_CUDA(cudaStreamSynchronize(h->stc));      // rows handed to sclmd_md_set_noise_rows have been consumed by the DMA
    return SCLMD_OK;
}

static int get_slots(sclmd_md *h, const double *dev, double *host) {  // device [nmd][ntraj] -> host [ntraj][nmd]
    std::vector<double> tmp((size_t)h->nmd * h->ntraj);
    // on the handle's own (non-blocking) stream: ordered after every queued step, e.g. after sclmd_md_run(h, n, NULL)
    SCLMD_CUDA(cudaMemcpyAsync(tmp.data(), dev, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    for (int k = 0; k < h->nmd; ++k)
        for (int tr = 0; tr < h->ntraj; ++tr) host[(size_t)tr * h->nmd + k] = tmp[(size_t)k * h->ntraj + tr];
    return SCLMD_OK;
}

static int set_slots(sclmd_md *h, double *dev, const double *host) {  // host [ntraj][nmd] -> device [nmd][ntraj]
    std::vector<double> tmp((size_t)h->nmd * h->ntraj);
    for (int k = 0; k < h->nmd; ++k)
        for (int tr = 0; tr < h->ntraj; ++tr) tmp[(size_t)k * h->ntraj + tr] = host[(size_t)tr * h->nmd + k];
    SCLMD_CUDA(cudaMemcpyAsync(dev, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    return SCLMD_OK;
}

// restart: the observables recorded so far by an earlier process (bath.cur, md.etot; [ntraj][nmd], index t % nmd)
int sclmd_md_set_current(sclmd_md *h, int bath, const double *cur) {
    if (int e = check_bath(h, bath, "sclmd_md_set_current")) return e;
    SCLMD_REQUIRE(cur, "sclmd_md_set_current: NULL buffer");
    SCLMD_CUDA(cudaSetDevice(h->device));
    return set_slots(h, h->baths[bath]->cur.p, cur);
}
int sclmd_md_set_etot(sclmd_md *h, const double *etot) {
    SCLMD_REQUIRE(h && etot, "sclmd_md_set_etot: NULL argument");
    SCLMD_CUDA(cudaSetDevice(h->device));
    return set_slots(h, h->etot.p, etot);
}

int sclmd_md_get_current(sclmd_md *h, int bath, double *cur) {
    if (int e = check_bath(h, bath, "sclmd_md_get_current")) return e;
    SCLMD_REQUIRE(cur, "sclmd_md_get_current: NULL buffer");
    SCLMD_CUDA(cudaSetDevice(h->device));
    return get_slots(h, h->baths[bath]->cur.p, cur);
}

int sclmd_md_get_etot(sclmd_md *h, double *etot) {
    SCLMD_REQUIRE(h && etot, "sclmd_md_get_etot: NULL argument");
    SCLMD_CUDA(cudaSetDevice(h->device));
    return get_slots(h, h->etot.p, etot);
}

int sclmd_md_get_current_sums(sclmd_md *h, int bath, double *sums) {
    if (int e = check_bath(h, bath, "sclmd_md_get_current_sums")) return e;
    SCLMD_REQUIRE(sums, "sclmd_md_get_current_sums: NULL buffer");
    SCLMD_CUDA(cudaSetDevice(h->device));
    k_sum_slots<<<cdiv(h->ntraj, 128), 128, 0, h->st>>>(h->baths[bath]->cur.p, h->nmd, h->ntraj, h->scratch.p);
    SCLMD_CUDA(cudaGetLastError());
    ++h->launches;
    SCLMD_CUDA(cudaMemcpyAsync(sums, h->scratch.p, h->ntraj * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    return SCLMD_OK;
}

int64_t sclmd_md_launch_count(sclmd_md *h) { return h ? h->launches : -1; }

// bath.gnoi() for the whole ensemble (baths.py:176-192, 397-409): fills the device noise table from a
// noise plan, trajectory k using the Philox stream of global trajectory traj0+k.  No host round trip.
int sclmd_md_generate_noise(sclmd_md *h, int bath, sclmd_noise_plan *plan, uint64_t seed, int64_t traj0) {
    if (int e = check_bath(h, bath, "sclmd_md_generate_noise")) return e;
    SCLMD_REQUIRE(plan, "sclmd_md_generate_noise: NULL plan");
    Bath &b = *h->baths[bath];
    int pn = 0, pc = 0;
    if (int e = sclmd_noise_plan_dims(plan, &pn, &pc)) return e;
    SCLMD_REQUIRE(pn == h->nmd && pc == b.nc, "sclmd_md_generate_noise: plan is for nmd=%d nc=%d, bath needs nmd=%d nc=%d", pn, pc, h->nmd, b.nc);
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->flush()) return e;
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    return sclmd_noise_plan_generate_into(plan, h->ntraj, seed, traj0, b.noise.p, h->ntraj, b.ncp, 0);      // table [ntraj][nmd][ncp]
}

// C[M x N] = alpha * A[M x K] . B[N x K]^T on the device with host buffers (the DMMA GEMM of dgemm.cuh).  The stand-alone force
// evaluations of the reference-shaped classes (md.potforce, md.force, bath.bforce: md.py:413-474, baths.py:224-255,448-458) are built on it.
int sclmd_dgemm_nt(int device, int M, int N, int K, const double *A, const double *B, double alpha, double *C) {
    SCLMD_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && C, "sclmd_dgemm_nt: bad arguments");
    if (int e = select_device(device)) return e;
    const int ldk = K + (K & 1), ldc = N + (N & 1);      // 16-byte aligned rows (pads are zero)
    DevBuf<double> dA, dB, dC;
    SCLMD_CUDA(dA.alloc((size_t)M * ldk)); SCLMD_CUDA(dB.alloc((size_t)N * ldk)); SCLMD_CUDA(dC.alloc((size_t)M * ldc));
    SCLMD_CUDA(cudaMemcpy2D(dA.p, ldk * sizeof(double), A, K * sizeof(double), K * sizeof(double), M, cudaMemcpyHostToDevice));
    SCLMD_CUDA(cudaMemcpy2D(dB.p, ldk * sizeof(double), B, K * sizeof(double), K * sizeof(double), N, cudaMemcpyHostToDevice));
    if (tma_usable(M, N)) {       // persistent TMA / stream-K kernel
        TmaWorkspace w;
        if (int e = launch_dgemm_tma_plain(M, N, K, dA.p, ldk, dB.p, ldk, dC.p, ldc, alpha, w, sm_count(device), nullptr)) return e;
        SCLMD_CUDA(cudaDeviceSynchronize());
    } else {
        GemmArgs g{};
        g.M = M; g.N = N; g.Kseg = ldk; g.nseg = 1; g.segs_per_split = 1;
        g.A = dA.p; g.lda = ldk; g.B = dB.p; g.ldb = ldk; g.C = dC.p; g.ldc = ldc; g.alpha = alpha;
        SCLMD_CUDA(launch_dgemm(g, 1, nullptr));
        SCLMD_CUDA(cudaDeviceSynchronize());
    }
    SCLMD_CUDA(cudaMemcpy2D(C, N * sizeof(double), dC.p, ldc * sizeof(double), N * sizeof(double), M, cudaMemcpyDeviceToHost));
    return SCLMD_OK;
}

int sclmd_md_time_tail(sclmd_md *h, int bath, int reps, float *avg_ms) {
    if (int e = check_bath(h, bath, "sclmd_md_time_tail")) return e;
    SCLMD_REQUIRE(reps > 0 && avg_ms, "sclmd_md_time_tail: bad arguments");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->flush()) return e;
    Bath &b = *h->baths[bath];
    SCLMD_REQUIRE(b.ml > 1, "sclmd_md_time_tail: bath %d is time-local (no history tail)", bath);
    // tailp is scratch between steps only in the sense that re-running with the same head is idempotent
    long long head = (h->t - 1) % b.ml;
    if (head < 0) head += b.ml;
    if (int e = h->tail_direct(b, (int)head)) return e;  // warm-up
    SCLMD_CUDA(cudaEventRecord(h->ev0, h->st));
    for (int r = 0; r < reps; ++r)
        if (int e = h->tail_direct(b, (int)head)) return e;
    SCLMD_CUDA(cudaEventRecord(h->ev1, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    float ms = 0;
    SCLMD_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *avg_ms = ms / reps;
    b.far_t0 = -1;   // restore the tail in the handle's own mode
    if (int e = h->tail_step(b, h->t - 1)) return e;
    if (int e = h->flush_far()) return e;
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    return SCLMD_OK;
}

int sclmd_md_time_potforce(sclmd_md *h, int reps, float *avg_ms) {
    SCLMD_REQUIRE(h && reps > 0 && avg_ms, "sclmd_md_time_potforce: bad arguments");
    SCLMD_CUDA(cudaSetDevice(h->device));
    if (int e = h->sync_real()) return e;
    if (int e = h->potforce(h->q.p, h->Gn.p)) return e;
    SCLMD_CUDA(cudaEventRecord(h->ev0, h->st));
    for (int r = 0; r < reps; ++r)
        if (int e = h->potforce(h->q.p, h->Gn.p)) return e;
    SCLMD_CUDA(cudaEventRecord(h->ev1, h->st));
    SCLMD_CUDA(cudaStreamSynchronize(h->st));
    float ms = 0;
    SCLMD_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *avg_ms = ms / reps;
    return SCLMD_OK;
}

}  // extern "C"
