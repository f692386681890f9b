// Persistent, warp-specialised FP64 "NT" GEMM on the DMMA pipe with TMA operand staging and stream-K scheduling.
//
//   C_z[M x N] = alpha * sum_{seg < nseg} A_{z,seg}[M x K] . B_{z,seg}[N x K]^T          (both operands K-contiguous)
//
// * One producer warp (one elected lane) moves 16-double (128-byte) K-slabs of the A and B tiles into shared memory with
//   cp.async.bulk.tensor (SASS UTMALDG) against full/empty mbarrier pairs, STAGES deep; it runs ahead across tile boundaries.
//   The tensor maps use the 128-byte swizzle: chunk c of row r lands in chunk c ^ (r & 7), so shared memory is dense (no row
//   padding) and still conflict-free for the fragment loads below.  Ragged M, N, K need no code: TMA zero-fills out of bounds.
// * Eight consumer warps (two warpgroups, 232 registers per thread after setmaxnreg; the producer warpgroup keeps 40) issue mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; tcgen05 has no FP64 kind).  A lane fetches 16 bytes = two
//   consecutive k of one row per LDS.128 and feeds .x / .y to two DMMAs: k-slot s of the first is physical k = 2s, of the second
//   2s+1 (A and B agree, so the contraction is merely reordered).  Fragment row j reads tile row pi(j) = (j >> 1) | ((j & 1) << 2)
//   of its 8-row group: the eight lanes of a quarter-warp then touch rows with swizzle keys r and r ^ 4, i.e. all eight 16-byte
//   chunks of a 128-byte line -- no bank conflict.  The same permutation maps accumulator rows / columns back on the way out.
// * Stream-K: the grid is one CTA per SM; the iteration space (tiles x K-slabs) is cut into equal contiguous ranges, so there is
//   no wave quantisation and no split-K partial C in HBM.  A CTA works through its range from the LAST tile to the first: the piece
//   of a tile that does not reach the tile's end is stored to a small workspace right away (at most one per CTA) and a flag is
//   raised; the CTA that holds the END of the tile -- its first tile, which it does last -- adds the partials of the earlier pieces
//   in descending CTA order (deterministic) and writes C once.  A CTA therefore waits only for CTAs with a LOWER logical index, and
//   the logical index is a ticket drawn when the CTA starts: every CTA that is waited for is already running, so the kernel needs no
//   co-residency guarantee (plain launch: its CTAs start as SMs become free; a cooperative launch cost 1.9 % of the config-5 step).
//
// Operands are described as rank-3 tensors so that one kernel covers plain products (K.q, the eigenbasis gather / scatter),
// frequency-batched products (noise x = L xi) and contractions over a rotating history ring (full memory-kernel tails).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "dgemm.cuh"

namespace sclmd {

struct TmaGemmParams {
    int M, N;                 // extents of one C_z
    int mt, nt, ntiles;       // tiles per C_z and in total (nbatch * mt * nt)
    int KI, kslabs;           // K iterations per tile (nseg * kslabs) and per segment
    int a_mode;               // 0: A coordinates (k, m, z);  1: history ring (k, slot(seg), m), slot = (a_head - seg) mod a_mod
    int a_head, a_mod;
    int b_mode;               // 0: B coordinates (k, n, z);  1: segmented (k, n, b_seg0 + seg)
    int b_seg0;
    double *C;
    long long ldc, c_batch_stride;
    double alpha;
    double *ws;               // [gridDim.x][BM * BN] partial accumulator tiles (stream-K fix-up)
    unsigned *flags;          // [gridDim.x]
    unsigned epoch;
    unsigned *ticket;         // [2]: CTAs started / CTAs finished in this launch; the last CTA to finish sets both back to zero, so the
                              // scheme survives CUDA-graph replays (nothing launch-specific is baked into the arguments)
};

#ifdef __CUDACC__

template <int BM, int BN, int WM, int WN, int STAGES>
struct TmaCfg {
    static constexpr int BK = 16;
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN, NW = WARPS_M * WARPS_N;
    // two consumer warpgroups + one producer warpgroup (one working lane): register budgets are moved between them with setmaxnreg
    static constexpr int CTHREADS = NW * 32, THREADS = CTHREADS + 128;
    static constexpr int FM = WM / 8, FN = WN / 8;
    static constexpr int STAGE_BYTES = (BM + BN) * BK * 8;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 2 * STAGES * 8 + 64;
};

__device__ __forceinline__ void tma_mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void tma_mbar_expect(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tTW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TD_%=;\n\tbra TW_%=;\n\tTD_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)),
                 "l"((unsigned long long)map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void consumer_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

template <int BM, int BN, int WM, int WN, int STAGES>
__global__ void __launch_bounds__(TmaCfg<BM, BN, WM, WN, STAGES>::THREADS, 1)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TmaGemmParams p) {
    using Cfg = TmaCfg<BM, BN, WM, WN, STAGES>;
    constexpr int NW = Cfg::NW, FM = Cfg::FM, FN = Cfg::FN, SB = Cfg::STAGE_BYTES, CT = Cfg::CTHREADS;
    extern __shared__ unsigned char tma_smem_raw[];
    unsigned char *sm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tma_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(sm + (size_t)STAGES * SB), *empty = full + STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            tma_mbar_init(&full[s], 1);
            tma_mbar_init(&empty[s], NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __shared__ unsigned s_bid;
    if (tid == 0) s_bid = atomicAdd(p.ticket, 1u);                          // logical index in the order in which the CTAs start
    __syncthreads();
    const unsigned bid = s_bid;
    const long long total = (long long)p.ntiles * p.KI;
    const long long beg = total * bid / gridDim.x, end = total * (bid + 1) / gridDim.x;
    const int per_z = p.mt * p.nt;

    if (warp >= NW) {   // ------------------------------------------------ producer warpgroup: one lane feeds the ring of stages
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == NW && lane == 0) {
            long long hi = end;
            unsigned fill = 0;
            while (hi > beg) {                                   // pieces of the range, last tile first
                const int tile = (int)((hi - 1) / p.KI);
                const long long tstart = (long long)tile * p.KI, lo = max(beg, tstart);
                const int kbeg = (int)(lo - tstart), kend = (int)(hi - tstart);
                const int z = tile / per_z, rem = tile % per_z;
                const int m0 = (rem % p.mt) * BM, n0 = (rem / p.mt) * BN;
                for (int kk = kbeg; kk < kend; ++kk, ++fill) {
                    const unsigned st = fill % STAGES;
                    if (fill >= (unsigned)STAGES) tma_mbar_wait(&empty[st], ((fill / STAGES) - 1) & 1);
                    tma_mbar_expect(&full[st], (unsigned)SB);
                    const int seg = kk / p.kslabs, k0 = (kk % p.kslabs) * Cfg::BK;
                    int a1 = m0, a2 = z, b2 = z;
                    if (p.a_mode == 1) {
                        int slot = (p.a_head - seg) % p.a_mod;
                        if (slot < 0) slot += p.a_mod;
                        a1 = slot;
                        a2 = m0;
                    }
                    if (p.b_mode == 1) b2 = p.b_seg0 + seg;
                    unsigned char *dst = sm + (size_t)st * SB;
                    tma_load_3d(dst, &mapA, &full[st], k0, a1, a2);
                    tma_load_3d(dst + BM * 128, &mapB, &full[st], k0, n0, b2);
                }
                hi = lo;
            }
        }
        return;
    }

    // -------------------------------------------------------------------- consumer warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int wm = (warp / Cfg::WARPS_N) * WM, wn = (warp % Cfg::WARPS_N) * WN;
    const int frow = lane >> 2, fcol = lane & 3;
    const int pr = (frow >> 1) | ((frow & 1) << 2);                 // tile row (within its group of 8) of fragment row `frow`
    const int ch0 = ((fcol) ^ pr) * 16, ch1 = ((4 + fcol) ^ pr) * 16;   // swizzled byte offsets of this lane's two 16-byte chunks
    const int arow = (wm + pr) * 128, brow = (wn + pr) * 128;
    double acc[FM][FN][2];
    long long hi = end;
    unsigned use = 0;
    while (hi > beg) {
        const int tile = (int)((hi - 1) / p.KI);
        const long long tstart = (long long)tile * p.KI, lo = max(beg, tstart);
        const int kbeg = (int)(lo - tstart), kend = (int)(hi - tstart);
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
            for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kk = kbeg; kk < kend; ++kk, ++use) {
            const unsigned st = use % STAGES;
            tma_mbar_wait(&full[st], (use / STAGES) & 1);
            const unsigned char *As = sm + (size_t)st * SB + arow, *Bs = sm + (size_t)st * SB + BM * 128 + brow;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ch = h ? ch1 : ch0;
                double2 b2[FN];
#pragma unroll
                for (int j = 0; j < FN; ++j) b2[j] = *reinterpret_cast<const double2 *>(Bs + j * 1024 + ch);
#pragma unroll
                for (int i = 0; i < FM; ++i) {      // one A fragment pair at a time: 64 accumulators leave room for little else
                    const double2 a2 = *reinterpret_cast<const double2 *>(As + i * 1024 + ch);
#pragma unroll
                    for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a2.x, b2[j].x);
#pragma unroll
                    for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a2.y, b2[j].y);
                }
            }
            __syncwarp();
            if (lane == 0) tma_mbar_arrive(&empty[st]);
        }
        if (kend < p.KI) {
            // this piece does not reach the end of the tile (it is the last tile of the range, done first): hand the partial
            // accumulators to the CTA that holds the end of the tile
            double *w = p.ws + (size_t)bid * (BM * BN) + tid;
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j) {
                    __stcg(w + (size_t)((i * FN + j) * 2) * CT, acc[i][j][0]);
                    __stcg(w + (size_t)((i * FN + j) * 2 + 1) * CT, acc[i][j][1]);
                }
            __threadfence();
            consumer_sync(CT);
            if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.flags + bid), "r"(p.epoch) : "memory");
        } else {
            if (kbeg > 0) {      // the start of the tile lives in the preceding CTAs (they did it first): add their partials, nearest first
                for (int cc = (int)bid - 1; cc >= 0 && total * (cc + 1) / gridDim.x > tstart; --cc) {
                    if (tid == 0) {
                        unsigned v;
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.flags + cc) : "memory");
                        } while (v != p.epoch);
                    }
                    consumer_sync(CT);
                    const double *w = p.ws + (size_t)cc * (BM * BN) + tid;
#pragma unroll
                    for (int i = 0; i < FM; ++i)
#pragma unroll
                        for (int j = 0; j < FN; ++j) {
                            acc[i][j][0] += __ldcg(w + (size_t)((i * FN + j) * 2) * CT);
                            acc[i][j][1] += __ldcg(w + (size_t)((i * FN + j) * 2 + 1) * CT);
                        }
                }
            }
            const int z = tile / per_z, rem = tile % per_z;
            const int m0 = (rem % p.mt) * BM, n0 = (rem / p.mt) * BN;
            double *C = p.C + (long long)z * p.c_batch_stride;
            // accumulator (row frow, columns 2 fcol, 2 fcol + 1) of a fragment belongs to tile row pi(frow), columns fcol and fcol + 4
#pragma unroll
            for (int i = 0; i < FM; ++i) {
                const int r = m0 + wm + i * 8 + pr;
                if (r >= p.M) continue;
#pragma unroll
                for (int j = 0; j < FN; ++j) {
                    const int c = n0 + wn + j * 8 + fcol;
                    double *dst = C + (long long)r * p.ldc + c;
                    if (c < p.N) dst[0] = p.alpha * acc[i][j][0];
                    if (c + 4 < p.N) dst[4] = p.alpha * acc[i][j][1];
                }
            }
        }
        hi = lo;
    }
    consumer_sync(CT);
    if (tid == 0 && atomicAdd(p.ticket + 1, 1u) == gridDim.x - 1) {             // last CTA out: counters and flags ready for the next launch on
        for (unsigned i = 0; i < gridDim.x; ++i) p.flags[i] = 0;                 // this workspace (also when a CUDA graph replays this very launch)
        p.ticket[0] = 0;
        p.ticket[1] = 0;
        __threadfence();
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct TmaOperand {       // rank-3 tensor, innermost dimension = K (contiguous)
    const double *base;
    unsigned long long dim[3];      // extents (elements): K, mid, outer
    unsigned long long stride[2];   // element strides of dims 1 and 2 (multiples of 2: 16-byte rows)
    unsigned box1, box2;            // box extents of dims 1 and 2 (one of them is the tile height, the other 1)
};

typedef CUresult (*sclmd_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline sclmd_encode_tiled_fn tma_encoder() {
    static sclmd_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<sclmd_encode_tiled_fn>(ptr);
    }
    return fn;
}
inline int tma_make_map(CUtensorMap *map, const TmaOperand &o) {
    sclmd_encode_tiled_fn enc = tma_encoder();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return SCLMD_ERR_CUDA;
    }
    const cuuint64_t dims[3] = {o.dim[0], o.dim[1], o.dim[2]};
    const cuuint64_t strides[2] = {o.stride[0] * 8, o.stride[1] * 8};
    const cuuint32_t box[3] = {16, o.box1, o.box2}, es[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(o.base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): base %p dims %llu %llu %llu strides %llu %llu box 16 %u %u", (int)r, (const void *)o.base,
                  o.dim[0], o.dim[1], o.dim[2], o.stride[0], o.stride[1], o.box1, o.box2);
        return SCLMD_ERR_CUDA;
    }
    return 0;
}

// per-stream scratch of the stream-K fix-up (launches that share it must be ordered on one stream)
struct TmaWorkspace {
    DevBuf<double> ws;
    DevBuf<unsigned> flags, ticket;
    unsigned epoch = 0;
    int grid = 0;
};

constexpr int TMA_BM = 128, TMA_BN_MAX = 160;
inline bool tma_disabled() {
    static const bool off = getenv("SCLMD_NO_TMA") != nullptr;
    return off;
}
// the TMA kernel pays off once a tile row is full; skinny products stay on the cp.async kernel of dgemm.cuh
inline bool tma_usable(int M, int N) { return !tma_disabled() && M >= 96 && N >= 32; }

struct TmaGemm {
    int M, N, K;              // one product: C[M x N], contraction length K per segment
    int nbatch, nseg;
    TmaOperand A, B;          // A.box1/box2 and B.box1/box2 are filled in by the launcher
    int a_mode, a_head, a_mod, b_mode, b_seg0;
    double *C;
    long long ldc, c_batch_stride;
    double alpha;
};

// tile width: 2 x 4 consumer warps of 64 x (BN / 4); the width that pads N least wins.
// nc = 300 (config 5): 320 columns (6 % padding) as 5 tiles of 64 or 2 of 160, instead of 3 of 128 (22 %)
inline int tma_pick_bn(int N) {
    const int cand[4] = {160, 128, 96, 64};
    int best = 128;
    long long best_pad = -1;
    // ties go to the narrower tile: with stream-K a narrow tile is split over fewer CTAs (smaller, fewer partial tiles at the end) and the
    // 64-wide configuration runs the main loop as well as the wide ones (config-5 gather product 0.141 -> 0.125 ms; SCLMD_TMA_TIE_WIDE=1: A/B)
    static const bool narrow = getenv("SCLMD_TMA_TIE_WIDE") == nullptr;
    for (int c : cand) {
        const long long padded = (long long)cdiv(N, c) * c;
        if (best_pad < 0 || padded < best_pad || (narrow && padded == best_pad)) { best_pad = padded; best = c; }
    }
    return best;
}

template <int BN, int STAGES>
inline int launch_dgemm_tma_bn(const TmaGemm &g, TmaWorkspace &w, int nsm, cudaStream_t st) {
    using Cfg = TmaCfg<TMA_BM, BN, 64, BN / 4, STAGES>;
    auto kern = dgemm_tma_kernel<TMA_BM, BN, 64, BN / 4, STAGES>;
    static bool configured = false;
    if (!configured) {
        SCLMD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        configured = true;
    }
    if (w.grid != nsm) {
        SCLMD_CUDA(w.ws.alloc((size_t)nsm * TMA_BM * TMA_BN_MAX));
        SCLMD_CUDA(w.flags.alloc(nsm));
        SCLMD_CUDA(w.ticket.alloc(2));
        w.grid = nsm;
        w.epoch = 0;
    }
    TmaOperand A = g.A, B = g.B;
    if (g.a_mode == 1) { A.box1 = 1; A.box2 = TMA_BM; } else { A.box1 = TMA_BM; A.box2 = 1; }
    B.box1 = BN; B.box2 = 1;
    CUtensorMap mA, mB;
    if (int e = tma_make_map(&mA, A)) return e;
    if (int e = tma_make_map(&mB, B)) return e;
    TmaGemmParams p{};
    p.M = g.M; p.N = g.N; p.mt = cdiv(g.M, TMA_BM); p.nt = cdiv(g.N, BN); p.ntiles = g.nbatch * p.mt * p.nt;
    p.kslabs = cdiv(g.K, Cfg::BK); p.KI = g.nseg * p.kslabs;
    p.a_mode = g.a_mode; p.a_head = g.a_head; p.a_mod = g.a_mod; p.b_mode = g.b_mode; p.b_seg0 = g.b_seg0;
    p.C = g.C; p.ldc = g.ldc; p.c_batch_stride = g.c_batch_stride; p.alpha = g.alpha;
    p.ws = w.ws.p; p.flags = w.flags.p; p.epoch = 1; ++w.epoch;
    const long long total = (long long)p.ntiles * p.KI;
    const int grid = (int)std::max<long long>(1, std::min<long long>(nsm, total));
    // CTAs of this grid wait on one another (stream-K fix-up), but only on CTAs that drew an earlier ticket: a plain launch is enough
    p.ticket = w.ticket.p;
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(mA, mB, p);
    const cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) {
        set_error("dgemm_tma_kernel launch failed: %s", cudaGetErrorString(ce));
        return SCLMD_ERR_CUDA;
    }
    return 0;
}
inline int launch_dgemm_tma(const TmaGemm &g, TmaWorkspace &w, int nsm, cudaStream_t st) {
    switch (tma_pick_bn(g.N)) {          // stages: as many 16-double slabs of (128 + BN) rows as fit ~190 KB
        case 160: return launch_dgemm_tma_bn<160, 5>(g, w, nsm, st);
        case 96: return launch_dgemm_tma_bn<96, 6>(g, w, nsm, st);
        case 64: return launch_dgemm_tma_bn<64, 8>(g, w, nsm, st);
        default: return launch_dgemm_tma_bn<128, 6>(g, w, nsm, st);
    }
}

// plain product C[M x N] = alpha A[M x K] . B[N x K]^T
inline int launch_dgemm_tma_plain(int M, int N, int K, const double *A, long long lda, const double *B, long long ldb, double *C, long long ldc,
                                  double alpha, TmaWorkspace &w, int nsm, cudaStream_t st) {
    TmaGemm g{};
    g.M = M; g.N = N; g.K = K; g.nbatch = 1; g.nseg = 1;
    g.A = TmaOperand{A, {(unsigned long long)K, (unsigned long long)M, 1}, {(unsigned long long)lda, (unsigned long long)lda * M}, 0, 0};
    g.B = TmaOperand{B, {(unsigned long long)K, (unsigned long long)N, 1}, {(unsigned long long)ldb, (unsigned long long)ldb * N}, 0, 0};
    g.C = C; g.ldc = ldc; g.c_batch_stride = 0; g.alpha = alpha;
    return launch_dgemm_tma(g, w, nsm, st);
}
#endif

}  // namespace sclmd
