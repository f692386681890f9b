// FP64 "NT" GEMM on the DMMA pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4, the only FP64
// tensor shape sm_100a has; tcgen05 has no FP64 kind).
//
//   C_z[M x N] = alpha * sum_{seg in split z}  A_seg[M x Kseg] . B_seg[N x Kseg]^T
//
// Both operands are K-contiguous.  "Segments" let one launch contract over a history
// ring buffer: segment s of A lives at slot (a_head - s) mod a_mod of the ring, segment
// s of B at (b_seg0 + s).  gridDim.z = number of K-splits; each split writes its own
// partial C (deterministic; the consumer adds them in order).
//
// Used for: the harmonic force  G = Q K^T  (md.py:467), full memory-kernel tails
// (baths.py:453-457), time-local friction / exim / zeta matrices (baths.py:236-249).
#pragma once
#include "common.cuh"

namespace sclmd {

struct GemmArgs {
    int M, N, Kseg, nseg, segs_per_split;
    int Ktot;  // > 0: the segments are consecutive K-slices of ONE operand pair; slice s covers [s*Kseg, min(Ktot,(s+1)*Kseg))
    const double *A;
    long long lda, a_seg_stride;
    int a_head, a_mod;  // a_mod > 0: ring addressing slot = (a_head - seg) mod a_mod
    const double *B;
    long long ldb, b_seg_stride;
    int b_seg0;
    double *C;
    long long ldc, c_split_stride;
    double alpha;
};

#ifdef __CUDACC__

__device__ __forceinline__ void cp_async16_zfill(void *smem, const void *gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int BM, int BN, int WM, int WN, int STAGES, int BK_ = 16>
struct GemmCfg {
    static constexpr int BK = BK_;
    static constexpr int LDS = BK + 4;  // 160 B row stride: conflict-free 64-bit fragment loads
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
    static constexpr int THREADS = WARPS_M * WARPS_N * 32;
    static constexpr int FM = WM / 8, FN = WN / 8;
    static constexpr size_t SMEM = (size_t)STAGES * (BM + BN) * LDS * sizeof(double);
};

template <int BM, int BN, int WM, int WN, int STAGES, int BK_ = 16>
__global__ void __launch_bounds__(GemmCfg<BM, BN, WM, WN, STAGES, BK_>::THREADS)
dgemm_nt_seg_kernel(const GemmArgs g) {
    using Cfg = GemmCfg<BM, BN, WM, WN, STAGES, BK_>;
    constexpr int BK = Cfg::BK, LDS = Cfg::LDS, THREADS = Cfg::THREADS, FM = Cfg::FM, FN = Cfg::FN;
    extern __shared__ __align__(16) double smem[];
    double *As = smem;                               // [STAGES][BM][LDS]
    double *Bs = smem + (size_t)STAGES * BM * LDS;   // [STAGES][BN][LDS]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp / Cfg::WARPS_N) * WM, wn = (warp % Cfg::WARPS_N) * WN;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int seg_begin = blockIdx.z * g.segs_per_split;
    const int seg_end = min(g.nseg, seg_begin + g.segs_per_split);
    const int ktiles = (g.Kseg + BK - 1) / BK;
    const int niter = max(0, seg_end - seg_begin) * ktiles;

    // per-thread invariants of the tile loads: a thread always copies the same 16-byte column chunk (kc) of rows
    // r0 + u * RSTEP, so the row offsets are computed once and an iteration only adds the K offset
    constexpr int CPR = BK / 2;                 // 16-byte chunks per row
    static_assert(THREADS % CPR == 0, "thread count must be a multiple of the chunks per row");
    constexpr int RSTEP = THREADS / CPR, NA = (BM + RSTEP - 1) / RSTEP, NB_ = (BN + RSTEP - 1) / RSTEP;
    const int kc = (tid % CPR) * 2, r0 = tid / CPR;
    long long a_off[NA], b_off[NB_];
#pragma unroll
    for (int u = 0; u < NA; ++u) a_off[u] = (long long)min(m0 + r0 + u * RSTEP, g.M - 1) * g.lda + kc;
#pragma unroll
    for (int u = 0; u < NB_; ++u) b_off[u] = (long long)min(n0 + r0 + u * RSTEP, g.N - 1) * g.ldb + kc;

    auto load_tile = [&](int it, int stage) {
        const int seg = seg_begin + it / ktiles;
        const int k0 = (it % ktiles) * BK;
        const int kvalid = g.Ktot > 0 ? min(g.Kseg, g.Ktot - seg * g.Kseg) : g.Kseg;
        long long aoff, boff;
        if (g.a_mod > 0) {
            int slot = (g.a_head - seg) % g.a_mod;
            if (slot < 0) slot += g.a_mod;
            aoff = (long long)slot * g.a_seg_stride;
        } else {
            aoff = (long long)seg * g.a_seg_stride;
        }
        boff = (long long)(g.b_seg0 + seg) * g.b_seg_stride;
        const int rem = kvalid - (k0 + kc);
        const int nb = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
        const long long kofs = nb ? k0 : -kc;      // nothing is read when nb == 0; keep the address inside the row
        double *as = As + (size_t)stage * BM * LDS + r0 * LDS + kc;
        double *bs = Bs + (size_t)stage * BN * LDS + r0 * LDS + kc;
        const double *ag = g.A + aoff + kofs, *bg = g.B + boff + kofs;
#pragma unroll
        for (int u = 0; u < NA; ++u)
            if (BM % RSTEP == 0 || r0 + u * RSTEP < BM) cp_async16_zfill(as + u * RSTEP * LDS, ag + a_off[u], nb);
#pragma unroll
        for (int u = 0; u < NB_; ++u)
            if (BN % RSTEP == 0 || r0 + u * RSTEP < BN) cp_async16_zfill(bs + u * RSTEP * LDS, bg + b_off[u], nb);
    };

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < niter) load_tile(s, s);
        cp_async_commit();
    }
    const int frow = lane >> 2, fcol = lane & 3;
    for (int it = 0; it < niter; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        // The prefetch of tile it+STAGES-1 (into the stage consumed at iteration it-1) is issued by warp w after its k-step
        // (w mod 4): right after the barrier every warp would otherwise run the same address arithmetic at the same time and
        // the tensor pipe would idle; staggered, some warp of every scheduler is always issuing DMMAs.
        const int nx = it + STAGES - 1;
        const double *as = As + (size_t)(it % STAGES) * BM * LDS + (wm + frow) * LDS + fcol;
        const double *bs = Bs + (size_t)(it % STAGES) * BN * LDS + (wn + frow) * LDS + fcol;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double a[FM], b[FN];
#pragma unroll
            for (int i = 0; i < FM; ++i) a[i] = as[i * 8 * LDS + kk];
#pragma unroll
            for (int j = 0; j < FN; ++j) b[j] = bs[j * 8 * LDS + kk];
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            if (kk / 4 == (((warp & 3) + 2 * (warp >> 2)) & 3)) {
                if (nx < niter) load_tile(nx, nx % STAGES);
                cp_async_commit();
            }
        }
    }
    cp_async_wait<0>();

    double *C = g.C + (long long)blockIdx.z * g.c_split_stride;
#pragma unroll
    for (int i = 0; i < FM; ++i) {
        const int r = m0 + wm + i * 8 + frow;
        if (r >= g.M) continue;
#pragma unroll
        for (int j = 0; j < FN; ++j) {
            const int c = n0 + wn + j * 8 + 2 * fcol;
            double *dst = C + (long long)r * g.ldc + c;
            if (c + 1 < g.N) {
                *reinterpret_cast<double2 *>(dst) = make_double2(g.alpha * acc[i][j][0], g.alpha * acc[i][j][1]);
            } else if (c < g.N) {
                dst[0] = g.alpha * acc[i][j][0];
            }
        }
    }
}

template <int BM, int BN, int WM, int WN, int STAGES, int BK_ = 16>
inline cudaError_t launch_dgemm_cfg(const GemmArgs &g, int nsplit, cudaStream_t st) {
    using Cfg = GemmCfg<BM, BN, WM, WN, STAGES, BK_>;
    auto kern = dgemm_nt_seg_kernel<BM, BN, WM, WN, STAGES, BK_>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), nsplit);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(g);
    return cudaGetLastError();
}

// Tile choice: big tiles when the problem fills the machine, small ones for skinny M (few
// trajectories) so that more CTAs exist.
inline cudaError_t launch_dgemm(const GemmArgs &g, int nsplit, cudaStream_t st, int force_cfg = -1) {
    const int cfg = force_cfg >= 0 ? force_cfg : (g.M > 64 ? 0 : (g.M > 16 ? 1 : 2));
    // 128x128 tiles, 32-deep K slabs (half the barriers of 16-deep ones: 29.6 -> 31.0 TFLOP/s on the K.q shape), 3 stages = 221 KB
    if (cfg == 0) return launch_dgemm_cfg<128, 128, 64, 32, 3, 32>(g, nsplit, st);
    if (cfg == 3) return launch_dgemm_cfg<128, 64, 32, 32, 3, 16>(g, nsplit, st);      // narrow N: 64-wide tiles, 2 CTAs/SM
    if (cfg == 1) return launch_dgemm_cfg<64, 64, 32, 32, 4>(g, nsplit, st);
    return launch_dgemm_cfg<16, 128, 16, 32, 4>(g, nsplit, st);
}

// Split-K plan for a single big product C = A.B^T on `sms` SMs: pick the tile config and the number of K-slices
// that minimise  rounds x per-unit work  (wave quantisation: 192 tiles of 128x128 on 148 SMs waste 35 %).
// Each slice writes its own partial C; consumers add the slices in a fixed order (deterministic).
struct SplitPlan {
    int cfg, nsplit, kseg;
};
// number of K-splits (<= max_split) that fills whole waves best: `tiles` output tiles, `slots` CTAs resident on the device
inline int wave_fit_splits(int tiles, int slots, int max_split) {
    int best = 1;
    double best_eff = 0.0;
    for (int ns = 1; ns <= max_split; ++ns) {
        const long long units = (long long)tiles * ns;
        const long long rounds = (units + slots - 1) / slots;
        const double eff = (double)units / (double)(rounds * slots);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = ns; }
    }
    return best;
}
inline SplitPlan plan_split_k(int M, int N, int K, int sms, int max_split) {
    SplitPlan best{M > 64 ? 0 : (M > 16 ? 1 : 2), 1, K};
    double best_cost = 1e300;
    const int bm[3] = {128, 64, 16}, bn[3] = {128, 64, 128}, per_sm[3] = {1, 3, 4};
    const double eff[3] = {1.0, 0.80, 0.45};
    for (int c = 0; c < 3; ++c) {
        if (c == 0 && M <= 64) continue;
        if (c == 2 && M > 16) continue;
        for (int ns = 1; ns <= max_split; ++ns) {
            const int kseg = round_up(cdiv(K, ns), 16);
            if (ns > 1 && kseg < 192) break;
            const int real_ns = cdiv(K, kseg);
            const long long units = (long long)cdiv(M, bm[c]) * cdiv(N, bn[c]) * real_ns;
            const long long slots = (long long)sms * per_sm[c];
            const double rounds = (double)((units + slots - 1) / slots);
            const double cost = rounds * per_sm[c] * bm[c] * bn[c] * (double)kseg / eff[c] + 2e5 * real_ns;
            if (cost < best_cost) {
                best_cost = cost;
                best = SplitPlan{c, real_ns, kseg};
            }
        }
    }
    return best;
}
#endif

}  // namespace sclmd
