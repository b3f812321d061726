/*
 * frankb200.h -- C ABI of libfrankb200.so: B200 (sm_100a) implementation of discsim/frank's
 * visibility -> Gaussian-process normal-equations path and of the dense FP64 solves that consume it.
 *
 * Every entry point names the reference interface it replaces (paths are into discsim/frank 1.2.3).
 * Conventions: plain pointers and sizes, no C++ or torch types; every function returns an int status
 *   0            success
 *   < 0          CUDA / argument error (text from fb_last_error)
 *   FB_E_QRANGE  data reach beyond the last collocation point (reference raises ValueError,
 *                frank/statistical_models.py:526-535)
 *   FB_E_NOTPD   Cholesky hit a non-positive pivot (reference: numpy.linalg.LinAlgError -> SVD fallback,
 *                frank/statistical_models.py:747)
 *   FB_E_BADP    non-positive / NaN power spectrum (reference ValueError, statistical_models.py:688-698)
 *   FB_E_NOCONV  the Jacobi SVD of the fallback path did not converge (scipy.linalg.svd raises LinAlgError)
 *   FB_E_SLOPE   the line search met a non-descending slope (reference: ValueError("Round off in slope calculation"),
 *                frank/minimizer.py:130-133)
 *   FB_E_RETRY   (fb_map_sync only) the data reach beyond the device's J0 table; the table has been rebuilt and the
 *                asynchronous call must be submitted again (the synchronous entry points do this themselves)
 * Nothing throws or aborts across the ABI.  A context is bound to one device; calls on one context are
 * serialised on its stream and are complete (host-visible) when the function returns unless stated.
 * "dev" pointers are device memory on the context's device, "host" pointers are host memory.
 */
#ifndef FRANKB200_H
#define FRANKB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FB_E_QRANGE 1
#define FB_E_NOTPD 2
#define FB_E_BADP 3
#define FB_E_NOCONV 4
#define FB_E_RETRY 5
#define FB_E_SLOPE 6

#define FB_MODEL_OPT_THICK 0
#define FB_MODEL_OPT_THIN 1
#define FB_MODEL_DEBRIS 2

typedef struct fb_ctx fb_ctx;

/* Geometry scalars exactly as the reference forms them on the host before touching the arrays
 * (frank/geometry.py:69-70, 111-115): a_ra = dRA * 2pi/rad_to_arcsec, a_dec likewise,
 * cos/sin of PA*deg_to_rad and inc*deg_to_rad. */
typedef struct fb_geometry {
    double a_ra, a_dec;
    double cos_pa, sin_pa;
    double cos_inc, sin_inc;
} fb_geometry;

/* ---- context ------------------------------------------------------------------------------- */
int fb_ctx_create(fb_ctx **ctx, int device);
int fb_ctx_destroy(fb_ctx *ctx);
const char *fb_last_error(fb_ctx *ctx);
int fb_version(void);
/* Tuning knobs of the host entry point's copy / compute pipeline (also read from the environment at fb_ctx_create):
 * "map_chunk" = smallest first chunk in visibilities (FB_MAP_CHUNK, default 2.5e5), "map_growth" = ratio of consecutive
 * chunk sizes (FB_MAP_GROWTH, default 1.5; 0 adapts it to the copy and kernel rates measured on the previous call, at the
 * price of call-to-call bit-reproducibility),
 * "map_kmax" = most chunks per call (FB_MAP_KMAX, default 8), "stage_threads" = host threads that gather pageable
 * inputs into the pinned staging ring (FB_STAGE_THREADS, default min(8, cores / LOCAL_WORLD_SIZE)), "force_staging" = treat
 * pinned inputs as pageable (tests). */
int fb_set_option(fb_ctx *ctx, const char *name, double value);

/* ---- DiscreteHankelTransform tables (frank/hankel.py:55-93) ---------------------------------
 * Host-side tables are O(N^2) setup and are passed in; the library builds the device-side J0
 * interpolation table covering arguments up to x_max (= a_max * j_nk[N-1]).
 *   j_nk[N]     first N zeros of J0                (hankel.py:72-73)
 *   coef[N]     norm * scale_factor = 1/(pi Qmax^2) / J1(j_nk)^2   (hankel.py:188, 201)
 *   Qmax        j_{N+1} / (2 pi Rmax)              (hankel.py:75)
 *   Ycoef[N*N]  coefficients(q=None) = 0.5 j_{N+1} norm Ykm, row-major (hankel.py:198-199), may be NULL
 *               when only the mapping is used. */
int fb_dht_setup(fb_ctx *ctx, int N, double Qmax, const double *host_j_nk, const double *host_coef,
                 const double *host_Ycoef, double x_max);

/* ---- VisibilityMapping.map_visibilities (frank/statistical_models.py:109-237) -----------------
 * Replaces: SourceGeometry.apply_correction (geometry.py:202-236), q = hypot (statistical_models.py:166),
 * _check_uv_range (:512-535), the chunk loop X = H(q); M += (X^T w) X; j += (X^T w) V (:192-214) and
 * H0 (:218).
 *   n            visibilities on this rank
 *   u, v         [n]      baselines / lambda
 *   V_reim       [2n]     complex visibilities, interleaved (re, im)  (numpy complex128 layout)
 *   w, w_stride  weights; w_stride = 1 for an array, 0 for a broadcast scalar (radial_fitters.py:544)
 *   chan, nchan  optional int32 channel index per visibility in [0, nchan) (np.unique order,
 *                statistical_models.py:180-189); NULL / 1 for a single channel
 *   kz2_H2       debris model only: H2[N] = 0.5 (2 pi h(r)/rad_to_arcsec)^2 (statistical_models.py:101-102)
 *   model_scale  cos(inc*deg_to_rad) for opt_thick, 1 otherwise (statistical_models.py:486-493)
 *   check_qbounds nonzero: return FB_E_QRANGE (outputs untouched except qminmax) when q_last < max(q)
 *   q_last       last collocation frequency q[N-1]
 * Outputs (dev or host according to the entry point):
 *   M [nchan*N*N] row-major, j [nchan*N], H0 [1] null likelihood (over all channels), qminmax [2] = min(q), max(q).
 * Multi-frequency data (statistical_models.py:175-214): chan[i] is the index of visibility i's frequency in
 * np.unique(frequencies); the channel becomes the high part of the sort key, every channel's run is padded to whole
 * tiles and accumulated by its own Gram launch -- one call, no host-side splitting.
 * In a multi-GPU job each rank passes its slice; with a communicator attached (fb_comm_init) the partial (M, j, H0)
 * are summed and qminmax / the status combined over the ranks inside the call, on the library's stream, before
 * anything is read back; without one the caller combines them (the sums are over independent visibilities).
 *
 * fb_map_visibilities_dev        inputs and outputs resident on the device; returns when the results are complete.
 * fb_map_visibilities_dev_async  the same, enqueue only: no host synchronisation and no host read anywhere in the call
 *                                (range check, J0-table check, sort scale, channel segments and the work table all
 *                                live on the device), so it can be overlapped with other work or followed by further
 *                                stream-ordered work of the library; fb_map_sync waits and returns the status
 *                                (0, FB_E_QRANGE, or FB_E_RETRY) and qminmax.
 * fb_map_visibilities_host       host arrays (pinned or pageable): a K-deep chunk pipeline (chunks of geometrically
 *                                growing size, see fb_set_option) copies chunk k+1 while chunk k is in the kernels;
 *                                pageable inputs are gathered into a pinned staging ring by a few host threads
 *                                first.  Chunks are summed in order and their sizes follow n alone: the same bits on every
 *                                call. */
int fb_map_visibilities_dev(fb_ctx *ctx, int64_t n, const double *dev_u, const double *dev_v,
                            const double *dev_V_reim, const double *dev_w, int w_stride,
                            const int32_t *dev_chan, int nchan, const fb_geometry *geom, int vis_model,
                            double model_scale, const double *host_H2, int check_qbounds, double q_last,
                            double *dev_M, double *dev_j, double *dev_H0, double *host_qminmax);

int fb_map_visibilities_dev_async(fb_ctx *ctx, int64_t n, const double *dev_u, const double *dev_v,
                                  const double *dev_V_reim, const double *dev_w, int w_stride,
                                  const int32_t *dev_chan, int nchan, const fb_geometry *geom, int vis_model,
                                  double model_scale, const double *host_H2, int check_qbounds, double q_last,
                                  double *dev_M, double *dev_j, double *dev_H0);
int fb_map_sync(fb_ctx *ctx, double *host_qminmax);

int fb_map_visibilities_host(fb_ctx *ctx, int64_t n, const double *host_u, const double *host_v,
                             const double *host_V_reim, const double *host_w, int w_stride,
                             const int32_t *host_chan, int nchan, const fb_geometry *geom, int vis_model,
                             double model_scale, const double *host_H2, int check_qbounds, double q_last,
                             double *host_M, double *host_j, double *host_H0, double *host_qminmax);

/* ---- multi-GPU: one process per GPU, one communicator per context ----------------------------------
 * The reference is a single process (no collective anywhere in frank/*.py); the visibility axis it already blocks over
 * (statistical_models.py:192-214) is the data-parallel axis here.
 *   fb_comm_unique_id   rank 0 creates the 128-byte NCCL id; the host framework broadcasts it (torch.distributed, MPI...)
 *   fb_comm_init        every rank joins with it (ncclCommInitRank on the context's device); from then on the mapping
 *                       entry points all-reduce their outputs in-call
 *   fb_comm_allgather   `count` doubles per rank, host buffers, rank order (results of a sweep sharded by grid point)
 *   fb_comm_allreduce_sum_dev   in-place sum of a device buffer on the library stream (asynchronous)
 *   fb_comm_info        returns 1 when a communicator is attached; rank / nranks optional
 * NCCL is loaded with dlopen at fb_comm_init: a single-GPU user never needs it. */
int fb_comm_unique_id(void *out128);
int fb_comm_init(fb_ctx *ctx, int nranks, int rank, const void *id128);
int fb_comm_destroy(fb_ctx *ctx);
int fb_comm_info(fb_ctx *ctx, int *rank, int *nranks);
int fb_comm_allgather(fb_ctx *ctx, const double *host_send, int64_t count, double *host_recv);
int fb_comm_allreduce_sum_dev(fb_ctx *ctx, double *dev_buf, int64_t count);

/* Timing of the most recent map call, milliseconds by CUDA events on the context's stream:
 * out[0] prepass (deproject / phase shift / hypot / H0 / min-max), out[1] J0+Gram kernel,
 * out[2] split-K reduction + scaling, out[3] host<->device copies (host entry point only). */
int fb_last_map_timing(fb_ctx *ctx, double *out4);

/* CUDA-event stopwatch on the context's stream (all library kernels are launched on it): start records an event,
 * stop records a second one, waits for it and returns the device time between them in milliseconds. */
int fb_timer_start(fb_ctx *ctx);
int fb_timer_stop(fb_ctx *ctx, double *elapsed_ms);

/* SourceGeometry.apply_correction (frank/geometry.py:202-236 = apply_phase_shift(inverse=True) :41-79 + deproject
 * :82-131) as a stand-alone pass over device-resident arrays, for callers that need the corrected arrays themselves
 * (uv binning of deprojected baselines).  Outputs (each optional, device): up, vp, wp [n]; Vp interleaved complex [2n]
 * (needs V); q = hypot(up, vp) [n], bit-equal to np.hypot.  Same correctly rounded operation order as the mapping's
 * pre-pass.  Returns after the stream has drained. */
int fb_apply_correction_dev(fb_ctx *ctx, int64_t n, const double *dev_u, const double *dev_v, const double *dev_V_reim,
                            const fb_geometry *geom, double *dev_up, double *dev_vp, double *dev_wp, double *dev_Vp_reim,
                            double *dev_q);

/* Pre-passed visibilities of the most recent map call in the (baseline-sorted) order the Gram kernel reads
 * them, for parity tests of the geometry pre-pass (geometry.py:202-236): a = q * (1/Qmax) [n], kz [n],
 * Re V' [n] and perm [n] (sorted position -> index into the caller's arrays), device -> host. */
int fb_debug_prepped(fb_ctx *ctx, int64_t n, double *host_a, double *host_kz, double *host_Vre, uint32_t *host_perm);

/* ---- GaussianModel (frank/statistical_models.py:650-781), batched over B power spectra ---------
 * D^-1 = M + Y^T diag(1/p_b) Y (has_prior != 0) or D^-1 = M; upper Cholesky D^-1 = U^T U (scipy.linalg.cho_factor
 * default); mu_b = D j.  Host pointers: M [N*N], j [N], p [B*N]; outputs mu [B*N], chol [B*N*N] (optional: the
 * upper triangle holds U), info [B] (0, or 1 + index of the first non-positive pivot).
 * Returns FB_E_NOTPD when a factorisation failed (reference: LinAlgError -> SVD fallback, :747-755),
 * FB_E_BADP for a non-positive / NaN spectrum (:688-698). */
int fb_gaussian_fit(fb_ctx *ctx, int B, const double *host_M, const double *host_j, const double *host_p, int has_prior,
                    double *host_mu, double *host_chol, int *host_info);

/* GaussianModel.Dsolve (frank/statistical_models.py:762-781, the Cholesky branch): X = (U^T U)^-1 B for nrhs right-hand sides
 * with the upper factor U [N*N] returned by fb_gaussian_fit / fb_frank_normal_loop.  B and X are [nrhs * N], one right-hand
 * side per row (host); one CTA per right-hand side (register-resident sweeps for N <= 512, shared-memory panels above).
 * Used for the posterior covariance (Dsolve of the identity), covariance_MAP and the Laplace evidence. */
int fb_chol_solve(fb_ctx *ctx, const double *host_U, int nrhs, const double *host_B, double *host_X);

/* SVD fallback of GaussianModel._fit (frank/statistical_models.py:747-755: scipy.linalg.svd(Dinv) when cho_factor
 * raises LinAlgError).  D^-1 as above (one power spectrum); outputs U [N*N], s [N] (descending), Vt [N*N] with
 * D^-1 = U diag(s) Vt, computed by one-sided Jacobi rotations on the device (D^-1 is symmetric: s_i = |lambda_i|,
 * the sign of lambda_i is folded into U).  The caller forms s1 = where(s > 0, 1/s, 0) and mu = Vt^T (s1 * (U^T j))
 * exactly as the reference does.  sweeps (optional) receives the number of Jacobi sweeps. */
int fb_gaussian_svd(fb_ctx *ctx, const double *host_M, const double *host_p, int has_prior, double *host_U,
                    double *host_s, double *host_Vt, int *host_sweeps);

/* ---- FrankFitter._fit power-spectrum loop, Normal method (frank/radial_fitters.py:765-785) --------
 * Device-resident loop, batched over B hyper-parameter points sharing M and j:
 *     while not converged(p, p_old) and count <= max_iter:
 *         p_old = p ; p = CriticalFilter.update_power_spectrum(fit) ; fit = GaussianModel(p) ; count += 1
 * entered with the fit of p_init.  alpha [B], p0 [B] are the inverse-gamma prior parameters (filter.py:170-173);
 * Tinv [B*N*N] is the dense inverse of the SPD pentadiagonal matrix T + I of each point (filter.py:23-62, 155;
 * condition number <= ~1e6), formed once per filter on the host: the reference's sparse solve (filter.py:175) becomes
 * a matrix-vector product.
 * Outputs: p [B*N] final spectrum, mu [B*N] its posterior mean, chol [B*N*N] (optional), niter [B] = count,
 * converged [B], info [B]; hist_p / hist_mu [B*hist_cap*N] (optional) receive every iteration's p and mu
 * (FrankFitter's iteration_diagnostics). */
int fb_frank_normal_loop(fb_ctx *ctx, int B, const double *host_M, const double *host_j, const double *host_p_init,
                         const double *host_alpha, const double *host_p0, const double *host_Tinv, double tol, int max_iter,
                         double *host_p, double *host_mu, double *host_chol, int *host_niter, int *host_converged,
                         int *host_info, double *host_hist_p, double *host_hist_mu, int hist_cap);

/* ---- LogNormalMAPModel (frank/statistical_models.py:998-1160), one channel / one field / unit scale --------
 * The Newton iteration's control flow (MinimizeNewton / LineSearch, frank/minimizer.py) stays with the caller;
 * these entry points do its O(N^2) / O(N^3) arithmetic on the device:
 *   fb_ln_setup            upload M [N*N], j [N]; s0 = log(I_scale); full_hessian as in the reference
 *   fb_ln_set_spectrum     S^-1 = Y^T diag(1/p) Y                           (:1064-1065)
 *   fb_ln_eval             f(s) = 1/2 s^T S^-1 s + 1/2 I^T M I - I.j, I = exp(s + s0); g(s) [N] optional (:1088-1111)
 *   fb_ln_newton_direction g(s) and dx = -Hess^-1 g(s); refactor != 0 rebuilds Hess(s) = diag(I) M diag(I) +
 *                          full_hessian diag(I o (M I - j)) + S^-1 and factorises it (:1113-1132; the reference
 *                          uses LU, minimizer.py:238 -- the Hessians met on the fit path are positive definite,
 *                          FB_E_NOTPD otherwise)
 *   fb_ln_posterior        factorise Hess(s_MAP) (:1148-1150), return the upper factor (optional) and one
 *                          CriticalFilter.update_power_spectrum with it (filter.py:154-177) */
int fb_ln_setup(fb_ctx *ctx, const double *host_M, const double *host_j, double s0, double full_hessian);
int fb_ln_set_spectrum(fb_ctx *ctx, const double *host_p);
int fb_ln_eval(fb_ctx *ctx, const double *host_s, double *host_f, double *host_g);
int fb_ln_newton_direction(fb_ctx *ctx, const double *host_s, int refactor, double *host_g, double *host_dx, int *host_info);
int fb_ln_posterior(fb_ctx *ctx, const double *host_s, const double *host_p, double alpha, double p0, const double *host_Tinv,
                    double *host_chol, double *host_p_new, int *host_info);

/* The whole log-normal fit in one call (K7): LogNormalMAPModel._fit = MinimizeNewton + LineSearch (statistical_models.py:
 * 1073-1160, minimizer.py:74-283) and, for max_iter >= 0, the power-spectrum iteration of FrankFitter._fit around it
 * (radial_fitters.py:765-785):
 *     fit = LogNormalMAPModel(p_init, guess)
 *     while not converged(p, p_old) and count <= max_iter:
 *         p_old = p ; p = CriticalFilter.update_power_spectrum(fit) ; fit = LogNormalMAPModel(p, guess=fit.MAP) ; count += 1
 * Every vector stays on the device; the host thread of the call takes the scalar decisions of the line search and of the
 * Newton iteration from a few doubles in mapped pinned memory.  The Hessian is factorised by the blocked Cholesky of
 * the Normal path where the reference uses LU (minimizer.py:238): the Hessians of a descending iteration are positive
 * definite; an indefinite one makes that step fall back to gradient descent, like a non-descending Newton direction
 * does in the reference (:244-253).  max_iter < 0: the single fit only (what LogNormalMAPModel's constructor does).
 * Host pointers: M [N*N], j [N], p_init [N], guess [N] (s = log I - s0), Tinv [N*N] as in fb_frank_normal_loop.
 * Outputs: s [N] the MAP point of the last fit, p [N] its power spectrum, chol [N*N] (optional) the upper factor of the
 * Hessian at s, niter = count, converged, info (potrf), stats [7] (optional) = Newton steps, function evaluations,
 * Hessians, and how many fits ended with MinimizeNewton status 0, 1, 2, 3; hist_p / hist_s [hist_cap*N] (optional).
 * Returns FB_E_BADP when an update produced a non-positive / NaN spectrum (reference: ValueError), FB_E_NOTPD when the
 * Hessian at a MAP point is not positive definite (reference: SVD pseudo-inverse, :1152-1158), FB_E_SLOPE. */
int fb_frank_lognormal_loop(fb_ctx *ctx, const double *host_M, const double *host_j, const double *host_p_init,
                            const double *host_guess, double s0, double full_hessian, double alpha, double p0,
                            const double *host_Tinv, double tol, int max_iter, double newton_tol, double *host_s,
                            double *host_p, double *host_chol, int *host_niter, int *host_converged, int *host_info,
                            long long *host_stats, double *host_hist_p, double *host_hist_s, int hist_cap);

/* ---- UVDataBinner (frank/utilities.py:180-400) -------------------------------------------------------------
 * fb_uv_max: max(uv) (the caller forms nbins = ceil(max / width) with the reference's guard, utilities.py:205-208).
 * fb_uv_bin: bin index per visibility (bit-exact with the reference's floor(uv * (1/width)) and its three fix-ups,
 *   :338-347), counts per bin (int64), sums [nbins*4] = (sum w uv, sum w, sum w Re V, sum w Im V) (:349-361) and
 *   err [nbins*2] = (sum w^2 (Re V - mu)^2, sum w^2 (Im V - mu)^2) with mu the weighted bin mean (:236-247).
 *   V is interleaved complex (v_is_complex != 0) or real.  Visibilities are stably sorted by bin and each bin
 *   is reduced in a fixed order (deterministic).
 * fb_uv_bin_dev: the same with every array already resident on the device (idx [n], counts [nbins], sums [nbins*4],
 *   err [nbins*2] are caller-allocated device buffers); returns after the stream has drained. */
int fb_uv_max(fb_ctx *ctx, int64_t n, const double *host_uv, double *host_max);
int fb_uv_bin(fb_ctx *ctx, int64_t n, const double *host_uv, const double *host_V, int v_is_complex, const double *host_w,
              int w_stride, double bin_width, int nbins, int32_t *host_idx, long long *host_counts, double *host_sums,
              double *host_err);
int fb_uv_bin_dev(fb_ctx *ctx, int64_t n, const double *dev_uv, const double *dev_V, int v_is_complex, const double *dev_w,
                  int w_stride, double bin_width, int nbins, int32_t *dev_idx, long long *dev_counts, double *dev_sums,
                  double *dev_err);

/* ---- VisibilityMapping.predict_visibilities (frank/statistical_models.py:279-329) ---------------------------
 * V_i = sum_k H_ik I_k with the same design rows as the mapping; q [n] deprojected baselines, kz [n] (debris model
 * only), I [N] brightness at the collocation points (host), V [n] out.  The J0 table grows by itself when q reaches beyond it.
 *   fb_predict_visibilities      host arrays
 *   fb_predict_visibilities_dev  q, kz, V resident on the device
 *   fb_predict_sky_dev           FrankRadialFit.predict (frank/radial_fitters.py:56-98) in one pass over device-resident
 *                                SKY-plane baselines u, v [n]: deproject (geometry.py:111-131), q = hypot, H(q) I, then
 *                                undo_correction (geometry.py:239-265: re-project, rotate the phase); V_reim [2n] interleaved
 *                                complex out.  The inner loop of FitGeometryFourierBessel (geometry.py:678-703). */
int fb_predict_visibilities(fb_ctx *ctx, int64_t n, const double *host_q, const double *host_kz, const double *host_I,
                            int vis_model, double model_scale, const double *host_H2, double *host_V);
int fb_predict_visibilities_dev(fb_ctx *ctx, int64_t n, const double *dev_q, const double *dev_kz, const double *host_I,
                                int vis_model, double model_scale, const double *host_H2, double *dev_V);
int fb_predict_sky_dev(fb_ctx *ctx, int64_t n, const double *dev_u, const double *dev_v, const fb_geometry *geom,
                       const double *host_I, int vis_model, double model_scale, const double *host_H2, double *dev_V_reim);

/* G = X^T X for K <= 8 device-resident columns x_a [n] (host array of K device pointers; G [K*K] host, row-major), summed in a
 * fixed order.  The normal equations of a Levenberg-Marquardt step whose residual and Jacobian columns stay on the device: what
 * FitGeometryFourierBessel (frank/geometry.py:745-746, scipy.optimize.least_squares on a 2n x 4 host Jacobian) needs per
 * iteration is J^T J (4 x 4), J^T r (4) and r.r. */
int fb_columns_gram_dev(fb_ctx *ctx, int64_t n, int K, const double *const *dev_cols, double *host_G);

/* J0 as the Gram kernel evaluates it (device table), for accuracy tests: out[i] = J0(x[i]).
 * fb_debug_j0 uses the table row nearest to x (the per-visibility gather path); fb_debug_j0_far uses the
 * neighbouring row on the far side of x, the worst case of the one-row-per-stage path. */
int fb_debug_j0(fb_ctx *ctx, int64_t n, const double *host_x, double *host_out);
int fb_debug_j0_far(fb_ctx *ctx, int64_t n, const double *host_x, double *host_out);

#ifdef __cplusplus
}
#endif
#endif
