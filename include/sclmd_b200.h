/* sclmd_b200.h -- C ABI of libsclmd_b200.so (hand-written sm_100a CUDA).
 *
 * The reference (ydsbbt/sclmd) is pure Python/NumPy and has NO FFI layer for
 * this path: its boundary is the Python class API.  The
 * entry points below are what a ctypes binding inside the reference's own
 * classes would call; each one cites the reference code it replaces.
 * INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; the message is in
 *     sclmd_last_error() (thread-local).  No exception or abort crosses the ABI.
 *   - the caller owns every host buffer (copied at call time); the library
 *     owns all device memory behind opaque handles; *_destroy releases it.
 *   - all floating point data is IEEE binary64, complex is interleaved (re,im).
 *   - one handle <-> one CUDA device + one stream; handles are not thread-safe,
 *     distinct handles are independent.
 *   - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef SCLMD_B200_H
#define SCLMD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCLMD_OK 0
#define SCLMD_ERR_ARG (-1)
#define SCLMD_ERR_CUDA (-2)
#define SCLMD_ERR_STATE (-3)
#define SCLMD_ERR_NOCONV (-4) /* sig: >=100 decimation iterations (selfenergy.py:127-130) */

const char *sclmd_last_error(void);
int sclmd_version(void);
/* number of visible CUDA devices (<0 on error) and basic properties */
int sclmd_device_count(void);
int sclmd_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, uint64_t *mem_bytes);

/* FP64 peak probes for the roofline denominators: kind 0 = DFMA chain (FP64 pipe),
 * kind 1 = DMMA.8x8x4 chain (FP64 tensor path). Result in TFLOP/s. */
int sclmd_probe_fp64(int device, int kind, double *tflops);

/* ------------------------------------------------------------------ MD ---
 * Ensemble velocity-Verlet integrator: replaces md.vv / md.force /
 * md.potforce(harmonic) / bath.bforce (sclmd/md.py:367-474,
 * sclmd/baths.py:224-255,448-458, sclmd/functions.py:146-153) for `ntraj`
 * independent noise realisations advanced together. */
typedef struct sclmd_md sclmd_md;

/* md.__init__ (md.py:56-130): nph dofs, time step dt, noise period nmd */
int sclmd_md_create(int nph, int ntraj, double dt, int nmd, int device, sclmd_md **out);
int sclmd_md_destroy(sclmd_md *h);

/* md.setDyn (md.py:250-292): K[nph*nph] row-major, already PSD-projected by the host */
int sclmd_md_set_dyn(sclmd_md *h, const double *K);
/* The eigen-decomposition md.setDyn projects with (md.py:266-281: K = U diag(lam) U^T, hw = sqrt(lam)): lam[nph] (clipped at 0),
 * U[nph*nph] row-major, eigenvectors as COLUMNS (numpy.linalg.eigh).  Optional.  With it the handle propagates in the eigenbasis
 * (Q = U^T q, Pi = U^T p: diagonal harmonic force, one gather and one scatter product over the bath dofs per step, 4 nph sum(nc)
 * flops per trajectory-step instead of 2 nph^2) whenever the problem allows it: no constraints, no force driver, diagonal-kernel
 * baths on disjoint dofs with 2 sum(nc) <= nph.  Same trajectories to rounding; q, p are converted back on every read.  Must follow
 * sclmd_md_set_dyn (checked: max |K U - U diag(lam)| <= 1e-9 max(lam)); a new sclmd_md_set_dyn discards it. */
int sclmd_md_set_modes(sclmd_md *h, const double *lam, const double *U);
/* 1 (default): eigenbasis propagation when possible; 0: always real space (A/B measurements, tests) */
int sclmd_md_set_modal(sclmd_md *h, int on);
/* 1 if the next sclmd_md_run propagates in the eigenbasis, 0 if in real space */
int sclmd_md_modal_active(sclmd_md *h);
/* md.AddConstr (md.py:189, 782-794): dofs zeroed in p,q after every step; n=0 clears */
int sclmd_md_set_constraint(sclmd_md *h, const int32_t *idx, int n);

#define SCLMD_KERNEL_FULL 0 /* kernel[ml][nc][nc] */
#define SCLMD_KERNEL_DIAG 1 /* kernel[ml][nc]     */
/* md.AddBath (md.py:167-183) for a bath whose force is
 *   f = noise[t%nmd] - c0*sum_{j<ml} kernel[j].p_{t-j}[cids]      (c0 = dt if ml>1 else 1;
 *       baths.py:448-458, 234-241)
 *     + Mq.q_t[cids] + Mp.p_t[cids]                                (optional, ml==1 only:
 *       Mq = bias*(exim - zeta1), Mp = -bias*zeta2; baths.py:243-249; pass NULL when the
 *       reference's all-three-non-zero test at baths.py:233 fails)
 * returns the bath index in *bath_out. */
int sclmd_md_add_bath(sclmd_md *h, const int32_t *cids, int nc, int ml, const double *kernel,
                      int kernel_kind, const double *Mq, const double *Mp, int *bath_out);

/* bath.noise (baths.py:191,408): noise[ntraj_sel][nmd][nc] for trajectories
 * [traj0, traj0+ntraj_sel) */
int sclmd_md_set_noise(sclmd_md *h, int bath, int traj0, int ntraj_sel, const double *noise);
int sclmd_md_get_noise(sclmd_md *h, int bath, int traj0, int ntraj_sel, double *noise);

/* streaming form of the same table: rows[nslab][ntraj][nc] for time slabs [slab0, slab0+nslab) mod nmd.
 * ASYNCHRONOUS (copy stream): the upload overlaps the step in flight, the next sclmd_md_run orders itself
 * after it; `rows` must stay valid until the next synchronising call.  Use pinned host memory. */
int sclmd_md_set_noise_rows(sclmd_md *h, int bath, int slab0, int nslab, const double *rows);

/* md.p, md.q, md.t (md.py:372,411): q,p are [ntraj][nph]; NULL pointers are skipped */
int sclmd_md_set_state(sclmd_md *h, const double *q, const double *p, int64_t t);
int sclmd_md_get_state(sclmd_md *h, double *q, double *p, int64_t *t);
/* md.ResetHis (md.py:340-349): zero all history rings and friction tails */
int sclmd_md_reset_history(sclmd_md *h);
/* md.phis (md.py:346,387) restricted to the bath dofs, reference order (row 0 = newest):
 * phis[ntraj][ml][nc] */
int sclmd_md_get_history(sclmd_md *h, int bath, double *phis);
int sclmd_md_set_history(sclmd_md *h, int bath, const double *phis);

/* `nsteps` calls of md.vv (md.py:367-411) for every trajectory.  With elapsed_ms != NULL the call
 * synchronises and returns the device time measured with CUDA events on the handle's stream; with
 * elapsed_ms == NULL it only enqueues the work (every sclmd_md_get_* synchronises). */
int sclmd_md_run(sclmd_md *h, int64_t nsteps, float *elapsed_ms);

/* bath.cur (md.py:397) and md.etot (md.py:383): [ntraj][nmd], index t % nmd */
int sclmd_md_get_current(sclmd_md *h, int bath, double *cur);
int sclmd_md_get_etot(sclmd_md *h, double *etot);
/* restart of an unfinished run (md.py:515-534): the slots recorded by the process that wrote the checkpoint, same layout */
int sclmd_md_set_current(sclmd_md *h, int bath, const double *cur);
int sclmd_md_set_etot(sclmd_md *h, const double *etot);
/* per-bath sum over the nmd slots of cur for each trajectory (np.mean(cur)*nmd, md.py:663):
 * sums[ntraj] -- the payload of the multi-GPU all-reduce */
int sclmd_md_get_current_sums(sclmd_md *h, int bath, double *sums);

/* etot and every bath's cur recorded at time slab `slab` (md.py:383,397):
 * out[1+nbaths][ntraj] = etot, cur_0, cur_1, ...
 * For the slab of the most recent step the call returns as soon as that step's first kernel (evaluation A, which
 * produces these numbers) has finished -- the rest of the step keeps running while the caller prepares the next one;
 * any other slab waits for all queued work.  Rows given to sclmd_md_set_noise_rows may be reused after it returns. */
int sclmd_md_get_step_observables(sclmd_md *h, int slab, double *out);

/* per-kernel timing with CUDA events on the handle's stream around every history-tail and
 * potential-force launch of subsequent sclmd_md_run calls; totals since it was switched on */
int sclmd_md_set_profiling(sclmd_md *h, int on);
int sclmd_md_get_profile(sclmd_md *h, double *tail_ms, int64_t *tail_launches, double *potforce_ms,
                         int64_t *potforce_launches);

/* 1 (default): the FP64-bound K.q GEMM runs on a second stream concurrently with the HBM-bound
 * history-tail kernels; 0: everything on one stream */
int sclmd_md_set_overlap(sclmd_md *h, int on);
/* 1 (default): sclmd_md_run on a handle whose baths are all time-local and diagonal (the reference's own example,
 * examples/runmd.py) is ONE launch of a persistent kernel: for <= 2 trajectories of <= 1024 dofs a cooperative kernel with one grid
 * barrier per step; for larger ensembles of <= 760 dofs (runs of >= 8 steps; up to four time-local baths, one of them may be a dense
 * one of <= 64 dofs) the ensemble kernel, eight trajectories per CTA through all steps with no grid barrier.  0: per-step launches */
int sclmd_md_set_persistent(sclmd_md *h, int on);

/* Force drivers (md.AddPotential, md.py:457-459, 481-485; protocol lammpsdriver.py:83-84): the potential force is a host callback.
 * on = 1: the handle never multiplies by K; sclmd_md_run is refused and a step is the pair below.  Forces are the driver's own
 * (mass-weighted, reference sign), [ntraj][nph].  Evaluations A, B, C of the bath forces, the history tails, the constraint and the
 * observables stay on the device; the callback runs once per step (twice with constraints: q_{t+1} = constrain(q') != q', md.py:449).
 *     if (sclmd_md_force_needed(h) == 1) sclmd_md_set_force(h, f(q_t));      first step, after set_state, with constraints
 *     sclmd_md_step_begin(h, q_trial);   q_trial[ntraj][nph] = q' of md.py:392
 *     sclmd_md_step_end(h, f(q_trial));  */
int sclmd_md_set_external_force(sclmd_md *h, int on);
int sclmd_md_force_needed(sclmd_md *h);
int sclmd_md_set_force(sclmd_md *h, const double *f);
int sclmd_md_step_begin(sclmd_md *h, double *q_trial);
int sclmd_md_step_end(sclmd_md *h, const double *f_trial);

/* md.f, md.fbaths, md.fhis (md.py:390-398, 403, 411).  on = 1: every step also stores the total force of evaluation C (md.f) and
 * each bath's force of evaluation A (the one the heat current is built from, md.fhis) and of evaluation C (md.fbaths after vv);
 * after a step   sclmd_md_get_force(h, f[ntraj][nph])   and
 *                sclmd_md_get_bath_force(h, bath, evaluation (0 = A, 2 = C), fb[ntraj][nc])   return them. */
int sclmd_md_set_force_output(sclmd_md *h, int on);
int sclmd_md_get_force(sclmd_md *h, double *f);
int sclmd_md_get_bath_force(sclmd_md *h, int bath, int evaluation, double *fb);

/* How diagonal-kernel baths with ml >= 128 contract their history (same flops, same results to rounding in every mode):
 *   1 (default) time-blocked far/near tails over 32-step blocks: the ring is read from HBM once per 32 steps by the tensor-pipe
 *               far pass (Hankel x history products on DMMA.8x8x4, needs ml % 8 == 0, else mode 5), the pass of the next block
 *               worked off one slice per step so that every step costs the same;
 *   0           one full ring pass per step -- the direct single-tail algorithm on which the HBM roofline of SURVEY.md
 *               section 8d is defined;
 *   2, 3, 4, 5  earlier far-pass kernels kept for A/B measurements (plain loads; two-stage bulk copies; 32-step DFMA blocks;
 *               16-step DFMA blocks with a producer warp = the round-1 default). */
int sclmd_md_set_tail_block(sclmd_md *h, int on);
/* ms[4], n[4] since profiling was switched on: 0 direct tail, 1 potential force, 2 far pass, 3 near pass */
int sclmd_md_get_profile_all(sclmd_md *h, double *ms, int64_t *n);
/* ms[8], n[8]: the four kinds above, then the eigenbasis mode: 4 scatter product, 5 gather product, 6 bath-dof kernel, 7 modal update */
int sclmd_md_get_profile_ex(sclmd_md *h, double *ms, int64_t *n);

/* instrumentation: kernels launched by this handle so far; name/time of the dominant kernel */
int64_t sclmd_md_launch_count(sclmd_md *h);
/* C[M x N] = alpha * A[M x K] . B[N x K]^T on the device, host buffers in and out: the building block of the stand-alone force
 * evaluations md.potforce / md.force / bath.bforce (md.py:413-474, baths.py:224-255,448-458) of the Python layer. */
int sclmd_dgemm_nt(int device, int M, int N, int K, const double *A, const double *B, double alpha, double *C);

/* time `reps` launches of the history-tail kernel of `bath` alone (CUDA events): avg ms */
int sclmd_md_time_tail(sclmd_md *h, int bath, int reps, float *avg_ms);
int sclmd_md_time_potforce(sclmd_md *h, int reps, float *avg_ms);

/* --------------------------------------------------------------- noise ---
 * Coloured-noise generation for an ensemble: replaces noise.phnoise / noise.enoise
 * (sclmd/noise.py:50-100, 149-206), vargau (273-305) and myfft.iFourier1D
 * (functions.py:36-53).
 *
 * A plan holds, for every frequency w_i = 2 pi i/(dt nmd), i = 0..nmd/2, the factor
 *   L_i = V sqrt(clamp+(lambda))   of   A_i = hermitianize( sum_m (cre+i cim)[i][m] basis[idx[i][m]] )
 * computed on the device (one-sided Jacobi).  The host supplies the scalar weights:
 *   phnoise: basis = gamma grid, two terms from flinterp (functions.py:117-134) times
 *            (dt nmd) equ(w) (noise.py:77-78);   enoise: basis = {efric, exip, exim} with
 *            the weights of noise.py:174-185.  idx < 0 marks an unused term. */
typedef struct sclmd_noise_plan sclmd_noise_plan;
int sclmd_noise_plan_create(int device, int nmd, double dt, int nc, int nbasis, const double *basis,
                            int nterm, const int32_t *idx, const double *cre, const double *cim,
                            sclmd_noise_plan **out);
int sclmd_noise_plan_destroy(sclmd_noise_plan *pl);
int sclmd_noise_plan_dims(sclmd_noise_plan *pl, int *nmd, int *nc);
int sclmd_noise_plan_is_complex(sclmd_noise_plan *pl);
/* L[nmd/2+1][nc][nc] (interleaved complex if the plan is complex); set_factors injects e.g. the
 * reference's own eigen-factors for deterministic parity runs */
int sclmd_noise_plan_get_factors(sclmd_noise_plan *pl, double *L);
int sclmd_noise_plan_set_factors(sclmd_noise_plan *pl, const double *L, int is_complex);
/* x_i = L_i xi_i, mirror to negative frequencies (noise.py:87-94), series = Re FFT/(dt nmd)
 * (functions.py:51-53, baths.py:191,408).  xi: injected standard normals [ntraj][nmd/2+1][nc], or
 * NULL -> Philox4x32-10 + Box-Muller with counter (i, k, traj0+traj) and key seed.
 * out: [ntraj][nmd][nc] */
int sclmd_noise_plan_generate(sclmd_noise_plan *pl, int ntraj, const double *xi, uint64_t seed,
                              int64_t traj0, double *out);
/* same, straight into a trajectory-major device table [ntraj_total][nmd][ncp_table], trajectories
 * [traj_offset, traj_offset + ntraj) (used by sclmd_md_generate_noise; `table` is a DEVICE pointer) */
int sclmd_noise_plan_generate_into(sclmd_noise_plan *pl, int ntraj, uint64_t seed, int64_t traj0,
                                   double *table, int ntraj_total, int ncp_table, int traj_offset);
int64_t sclmd_noise_plan_launch_count(sclmd_noise_plan *pl);
/* ms[6]: device milliseconds (CUDA events on the plan's stream) of the stages of the last generate call (0 normal draws,
 * 1 x = L xi, 2 mirrored transform) and of the per-frequency factorisation at plan creation (3); how many frequencies were
 * factorised by the pivoted Cholesky kernel (4: positive semi-definite spectra) and by one-sided Jacobi (5: the others) */
int sclmd_noise_plan_get_profile(sclmd_noise_plan *pl, double *ms);
/* bath.gnoi() for every trajectory of an MD handle, no host round trip (baths.py:176-192,397-409) */
int sclmd_md_generate_noise(sclmd_md *h, int bath, sclmd_noise_plan *plan, uint64_t seed, int64_t traj0);

/* baths.gamt, eta_ad == 0 branch (baths.py:35-42):
 *   out[nt][m] = 2*mean_i( giT[m][i]*cos(wl_i*tl_t) )*wl[nw-1]/pi
 * giT[m][nw] = flinterp(wl_i, gwl, gam) flattened over the nc*nc (or nc) matrix entries and
 * transposed by the host (the interpolation indices are host-side, functions.py:117-143). */
int sclmd_gamt(int device, int nt, int nw, int m, const double *tl, const double *wl,
               const double *giT, double *out);
/* functions.myfft (functions.py:11-53) and the power spectra built on it (functions.py:203-236): batched complex transform with the
 * in-house radix-2/3/4/5 kernels, out[f][m] = scale * sum_k in[f][k] exp(sign * 2 pi i k m / n); planar [batch][n], im may be NULL.
 * myfft.iFourier1D = sign -1, scale dw/2pi; myfft.Fourier1D = sign +1, scale (2pi/dw)/n. */
int sclmd_fft(int device, int n, int batch, const double *re, const double *im, int sign, double scale, double *out_re,
              double *out_im);

/* general form: out[nt][m] = alpha * sum_i c(tl_t, wl_i) giT[m][i] with
 *   c = cos(w t)                                                   (eta == 0)
 *   c = e^{-eta t}(w^2 cos wt + w eta sin wt)/(w^2 + eta^2)        (eta != 0: half of the bracket of baths.py:48-49)
 * used for gamt with artificial damping (baths.py:43-50, alpha = 2 wl[-1]/(pi nw)) and for the re-derived
 * gamma(w) = dt sum_t kernel(t) cos(w t) of phbath.gmem (baths.py:437-445, alpha = dt) */
int sclmd_cos_transform(int device, int nt, int nw, int m, const double *tl, const double *wl,
                        const double *giT, double eta, double alpha, double *out);

/* Frees the device scratch the noise generator caches between calls (the frequency sweeps keep theirs in their handles). */
int sclmd_release_workspace(void);

/* ---------------------------------------------------------------- NEGF ---
 * The reference's class bpt (negf.py:8-277) as a handle: bpt.__init__ / getdynmat (negf.py:8-25, 39-102) leave the reduced
 * dynamical matrix (fixed dofs removed), the dofs of the two leads and the damping; sclmd_bpt_create uploads K once and the
 * handle owns everything its sweeps need on the device (batch workspace, streams, profiling state) until sclmd_bpt_destroy.
 * No state of the sweeps lives outside the handle; one sweep at a time per handle, handles are independent.
 *   K[n*n] row-major, idxL/idxR reduced indices (dofatomofbath - len(dofatomfixed[0]), negf.py:195-204), damp in ps, n <= 4096.
 * Diagonal lead self-energies Sigma = -i w/damp on the bath dofs (negf.py:153-157). */
typedef struct sclmd_bpt sclmd_bpt;
int sclmd_bpt_create(int device, int n, const double *K, const int32_t *idxL, int nL, const int32_t *idxR, int nR,
                     double damp, sclmd_bpt **out);
int sclmd_bpt_destroy(sclmd_bpt *h);
/* bpt.setbias (negf.py:27-37): a biased electron bath, Sigma_b^r = -i w bdamp - bias chiminus on the contiguous dof block
 * [b0, b0+nb) (negf.py:162-172; reduced numbering), bias in angular units (eV/hbar, as bpt stores it);
 * bdamp, chiplus, chiminus: [nb*nb], copied.  nb == 0 removes the block. */
int sclmd_bpt_set_bias(sclmd_bpt *h, int b0, int nb, const double *bdamp, const double *chiplus,
                       const double *chiminus, double bias);
/* bpt.tm over a frequency list (negf.py:104-119, 206-208, 240-242): T(w) = Re Tr[G Gamma_L G^dagger Gamma_R]; G carries the
 * retarded self-energy of the bias block when one is set.  Singular M(w): SCLMD_ERR_STATE (numpy.linalg.LinAlgError there). */
int sclmd_bpt_tm(sclmd_bpt *h, const double *omegas, int nw, double *tm_out);
/* bpt.ps without bias (negf.py:232): -2 w^2 nB Tr Im G[sel,sel]; nb[nw] = bosedist(w,T).  Error if a bias block is set. */
int sclmd_bpt_ps(sclmd_bpt *h, const double *omegas, const double *nb, int nw, const int32_t *sel, int nsel,
                 double *ps_out);
/* bpt.retargf / bpt.advangf (negf.py:206-212): the full Green function, green_out[nw][n][n] complex (interleaved re, im);
 * advanced != 0: advanced self-energies (the +i eps of z is kept, as negf.py:212 does). */
int sclmd_bpt_green(sclmd_bpt *h, const double *omegas, int nw, int advanced, double *green_out);
/* bpt.ps with bias (negf.py:234-236): w^2 Re Tr[(G^r Sigma^K G^a)[sel,sel]] with Sigma^K = totalkselfenergy
 * (negf.py:177-193) = kd[w] on the lead dofs + kr1[w] bdamp + kr2[w] chiplus + i ki[w] chiminus on the bias block;
 * the per-frequency weights carry the Bose factors (bosedist edge cases stay on the host side). */
int sclmd_bpt_ps_bias(sclmd_bpt *h, const double *omegas, const double *kd, const double *kr1, const double *kr2,
                      const double *ki, int nw, const int32_t *sel, int nsel, double *ps_out);
/* Per-kernel-class CUDA-event timing of this handle's sweeps (build, panel, rank-16 update, block trsm, rank-64 update, back
 * substitution, observable): ms[7], n[7], flops executed by the rank-64 update, device time of the last sweep.
 * Switching it on resets the totals; while on, sweeps run on one stream. */
int sclmd_bpt_set_profiling(sclmd_bpt *h, int on);
int sclmd_bpt_get_profile(sclmd_bpt *h, double *ms, int64_t *n, double *gemm_flops, double *last_device_ms);

/* sig.selfenergy / sig.getse (selfenergy.py:105-140, 153-166): Sancho-Rubio decimation.
 *   K00,K11,K01,K10: [m*m]; direction 'L' or 'R'; se_out: [nw][m][m] interleaved complex;
 *   iters_out (may be NULL): [nw] decimation iterations */
int sclmd_sig_selfenergy(int device, int m, const double *K00, const double *K11, const double *K01,
                         const double *K10, double eta, char direction, const double *omegas, int nw,
                         double *se_out, int32_t *iters_out);
/* sig.sgf (selfenergy.py:105-131): the surface Green function of one lead after the decimation,
 *   sgf_out[nw][m][m] complex (interleaved re, im); same arguments and errors as sclmd_sig_selfenergy */
int sclmd_sig_sgf(int device, int m, const double *K00, const double *K11, const double *K01,
                  const double *K10, double eta, char direction, const double *omegas, int nw,
                  double *sgf_out, int32_t *iters_out);
/* sig.retargf (selfenergy.py:145-147): green_out[nw][m][m] complex (interleaved re, im) */
int sclmd_sig_green(int device, int m, const double *K00, const double *K11, const double *K01,
                    const double *K10, double eta, const double *omegas, int nw, double *green_out);
/* sig.tm / sig.gettm (selfenergy.py:145-151, 168-178) */
int sclmd_sig_tm(int device, int m, const double *K00, const double *K11, const double *K01,
                 const double *K10, double eta, const double *omegas, int nw, double *tm_out);

#ifdef __cplusplus
}
#endif
#endif
