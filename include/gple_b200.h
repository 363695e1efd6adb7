/* gple_b200.h -- C-ABI of the B200-native GPR-MQCLE hot path.
 *
 * The reference (kaigu1997/gaussian_process_liouville_equation, directory
 * gaussian_process_liouville_equation/ = "gple/") has no FFI layer: its hot path sits behind ordinary
 * C++ headers.  Each entry point below replaces the work done behind one of those interfaces and is
 * what the reference-side shim binds.  Conventions:
 *   - every function returns an int status (GPLE_OK or an error code); no C++ exceptions cross;
 *   - array arguments may be HOST or DEVICE pointers (detected with cudaPointerGetAttributes);
 *     host inputs are copied to the device on entry, host outputs copied back before return;
 *   - phase-space coordinates are interleaved (x, p) pairs = the reference's `PhasePoints`
 *     (Eigen::Matrix<double, 2, Dynamic>, column-major; gple/stdafx.h:153);
 *   - complex numbers are interleaved (re, im) pairs = std::complex<double>;
 *   - matrices returned to the caller are column-major like Eigen::MatrixXd;
 *   - evolved points are the 32-byte AoS `PhaseSpacePoint` {x, p, Re rho, Im rho} (gple/storage.h:232-297);
 *   - non-finite scalars are returned as they are: the host shim applies `make_normal` (gple/opt.cpp:420-431).
 * A context is bound to one CUDA device and one stream; calls on one context are serialised by the caller.
 */
#ifndef GPLE_B200_H
#define GPLE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C"
{
#endif

	typedef struct gple_ctx gple_ctx;
	typedef struct gple_model gple_model; /* one trained density-matrix element (real or complex kernel) */

	enum gple_status
	{
		GPLE_OK = 0,
		GPLE_ERR_ARG = 1,	  /* bad argument (null pointer, zero size, wrong model kind) */
		GPLE_ERR_NOT_SPD = 2, /* factorisation met a non-positive pivot; scalars are NaN */
		GPLE_ERR_CUDA = 3,	  /* CUDA runtime failure, see gple_last_error */
		GPLE_ERR_STATE = 4,	  /* quantity was not computed for this model (flag missing) */
		GPLE_ERR_COMM = 5	  /* NCCL failure or NCCL not loadable (multi-GPU entry points only), see gple_last_error */
	};

	/* flags of gple_train_*: the three booleans of TrainingKernel's constructor (gple/kernel.h:128-134) */
	enum gple_train_flags
	{
		GPLE_CALC_ERROR = 1,
		GPLE_CALC_AVERAGE = 2,
		GPLE_CALC_DERIVATIVE = 4
	};

	/* Tully models (gple/pes.h:28-36); the reference selects one at compile time (pes.h:38-41) */
	enum gple_pes_model
	{
		GPLE_SAC = 0,
		GPLE_DAC = 1,
		GPLE_ECR = 2
	};

	/* ---- context -------------------------------------------------------------------------------- */
	int gple_ctx_create(int device, gple_ctx** ctx);
	int gple_ctx_destroy(gple_ctx* ctx);
	/* Launch on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = the context's own stream. */
	int gple_ctx_set_stream(gple_ctx* ctx, void* cuda_stream);
	int gple_ctx_sync(gple_ctx* ctx);
	const char* gple_last_error(const gple_ctx* ctx);
	/* Number of CUDA kernels this context has launched since creation (bench.py's `gpu_launches`). */
	unsigned long long gple_launch_count(const gple_ctx* ctx);
	const char* gple_version(void);
	/* Options.  GPLE_OPT_GATED_VARIANCE (default 1): when a caller only asks for the CUTOFF prediction (gple_evolve,
	 * gple_new_point_predict, gple_predict_* with var_out == NULL), the variance GEMM is run only for the queries whose
	 * gate (gple/kernel.h:301-332) is not already decided by a bound on the variance:
	 *   |f|^2 >= 4 k**                  => gate == 1  (the computed variance k** - sum Z^2 never exceeds the prior k**);
	 *   |f|^2 <= sigma_f^2 sigma_n^2 / 2 => gate == 0  (var >= sigma_f^2 sigma_n^2 for a query that coincides with no
	 *                                                   training point; only used while the noise is >= 1e-9 k**).
	 * The undecided queries then meet a STAGED bound: sum Z^2 over the columns of Z = K* L^-T that belong to the first 128-blocks
	 * of training points is the variance explained by those points alone, so
	 *   |f|^2 >= 4 (k** - partial sum)  => gate == 1,
	 * and only the queries this does not decide see further columns (the triangular products of the first blocks are the short
	 * ones: block t costs t + 1 tile products).  The schedule is a list of cumulative block boundaries; after each stage but the
	 * last the decided queries leave the list.  Automatic schedule (GPLE_OPT_GATE_STAGE_TILES = -1, the default): boundaries
	 * 1, then a geometric progression (ratio about 2.4) up to 21/32 of the blocks -- 1, 2, 5, 10 of 16 blocks; 1, 2, 6, 14, 35, 84 of
	 * 128 (complex element: no Im block in the first stage, one block of Im rows per five Re blocks afterwards, then all Re blocks
	 * with 3/8 of the Im blocks); gple_gate_schedule_automatic returns it, gple_ctx_set_gate_schedule replaces it.
	 * The decided queries get exactly the value the full computation gives; the others go through the variance GEMM in a
	 * different batch composition (summation order may differ).  GPLE_OPT_GATED_VARIANCE = 0 forces every variance;
	 * GPLE_OPT_GATE_STAGE_TILES = 0 disables the staged bound; a value t > 0 gives the two boundaries (t, GPLE_OPT_GATE_STAGE2_TILES)
	 * of the earlier schedule. */
	enum gple_option
	{
		GPLE_OPT_GATED_VARIANCE = 1,
		GPLE_OPT_GATE_STAGE_TILES = 2,
		GPLE_OPT_GATE_STAGE_TILES_IM = 3, /* with GPLE_OPT_GATE_STAGE_TILES = t > 0, complex element: blocks of Im rows in the stage (default -1 = a quarter of the Re blocks, at least 1) */
		/* (default 1) one step of iterative refinement of v = K^-1 y' after the factorisation (residual against the regenerated
		 * covariance): v at the accuracy of the reference's LDLT solve (kernel.cpp:281-284) instead of that of the explicit
		 * triangular inverse; 0 only for measuring the difference (tests/test_gpu_baseline_sizes.py) */
		GPLE_OPT_REFINE_SOLUTION = 4,
		/* (default 1) the factorisation (Cholesky + triangular inverse: about a hundred short, dependent launches on two streams at
		 * n = 2048) is captured once per (size, buffers) into a CUDA graph and replayed: the launch-bound inner loop of every model
		 * rebuild and of every loss evaluation of the optimiser.  Same kernels, same results; 0 for measuring the difference. */
		GPLE_OPT_FACTORISE_GRAPHS = 5,
		/* with GPLE_OPT_GATE_STAGE_TILES = t > 0: second boundary (default -1 = five eighths of the training blocks): the queries the stage-A bound leaves open
		 * first see the blocks [stage, stage2) only; the tighter bound decides another fifth of them before the long products of
		 * the last blocks.  A value <= the stage runs the rest in one part (round-1 schedule). */
		GPLE_OPT_GATE_STAGE2_TILES = 6
	};
	int gple_ctx_set_option(gple_ctx* ctx, int option, int value);
	/* Explicit schedule of the staged bound for the real (complex_element == 0) or the complex element: `stages` cumulative
	 * boundaries in 128-blocks of training points, re_end[k] blocks of (Re) rows and im_end[k] blocks of Im rows (ignored for the
	 * real element; may be NULL) seen after stage k; the last stage (everything left) is implied.  Boundaries beyond the model's
	 * block count are clipped, empty stages dropped.  stages == 0 restores the automatic schedule.  At most 15 stages. */
	int gple_ctx_set_gate_schedule(gple_ctx* ctx, int complex_element, int stages, const int* re_end, const int* im_end);
	/* The automatic schedule for an element model of `blocks` 128-blocks of training points (ceil(N / 128)): returns the number of
	 * stages before the last one (>= 0; negative on a bad argument) and writes the first `capacity` boundaries.  Host only (no
	 * context, no device): what gple_evolve / gple_predict_* will use unless a schedule or GPLE_OPT_GATE_STAGE_TILES is set. */
	int gple_gate_schedule_automatic(int complex_element, int blocks, int* re_end, int* im_end, int capacity);
	/* Gated predictions since the last call (then reset): out = {composite rows seen, rows sent through the first stage of the
	 * variance GEMM, rows decided gate == 0 by the noise floor, rows that reached the last stage (the full variance)}. */
	int gple_gate_statistics(gple_ctx* ctx, unsigned long long out[4]);

	/* ---- kernel matrices ---------------------------------------------------------------------------
	 * Replaces KernelBase::KernelBase + delta_kernel (gple/kernel.cpp:8-31,217-242) and calculate_derivative
	 * (kernel.cpp:168-215).  theta = (sigma_f, l_x, l_p, sigma_n) (gple/kernel.h:33,41).  same_set != 0 is the
	 * reference's `LeftFeature.data() == RightFeature.data()` (training-set) case.
	 * K_out: nL x nR; dK_out (may be NULL): 4 matrices nL x nR, one per parameter. */
	int gple_kernel_real(gple_ctx* ctx, const double* XL, size_t nL, const double* XR, size_t nR, const double theta[4], int same_set, double* K_out, double* dK_out);
	/* Replaces ComplexKernelBase::ComplexKernelBase (gple/complex_kernel.cpp:134-200).
	 * theta = (sigma, sigma_R, l_Rx, l_Rp, sigma_I, l_Ix, l_Ip, sigma_n) (complex_kernel.cpp:230-256).
	 * K_out: nL x nR real; Kt_out: nL x nR complex (pseudo-covariance). */
	int gple_kernel_complex(gple_ctx* ctx, const double* XL, size_t nL, const double* XR, size_t nR, const double theta[8], int same_set, double* K_out, double* Kt_out);
	/* calculate_derivative / calculate_pseudo_derivative of the complex kernel (gple/complex_kernel.cpp:20-59, 74-132), materialised:
	 * dK_out = 8 real nL x nR matrices, dKt_out = 8 complex ones (either may be NULL), parameter order sigma, sigma_R, l_Rx, l_Rp,
	 * sigma_I, l_Ix, l_Ip, sigma_n, column-major.  The training path never stores them (it works in composite form); this entry
	 * exists for the getters / parity with the reference's derivative arrays, quirk q2 included. */
	int gple_kernel_complex_derivatives(gple_ctx* ctx, const double* XL, size_t nL, const double* XR, size_t nR, const double theta[8], int same_set, double* dK_out, double* dKt_out);

	/* ---- training ----------------------------------------------------------------------------------
	 * Replaces TrainingKernel::TrainingKernel (gple/kernel.cpp:244-479) and its getters (kernel.h:136-243).
	 * X: N points; y: N complex labels (imaginary part ignored, kernel.cpp:279-280). */
	typedef struct gple_real_scalars
	{
		double rescale;		   /* get_rescale_factor()  kernel.cpp:279 */
		double error;		   /* get_error()           kernel.cpp:285 (LOOCV squared error) */
		double population;	   /* get_population()      kernel.cpp:286-297 */
		double first_order[2]; /* get_1st_order_average() kernel.cpp:298-312 */
		double purity;		   /* get_purity()          kernel.cpp:325-335 */
		double magnitude;	   /* get_magnitude()       kernel.h:167-179 */
		double d_error[4];	   /* get_error_derivative()      kernel.cpp:381-400 */
		double d_population[4]; /* get_population_derivative() kernel.cpp:401-435 */
		double d_purity[4];	   /* get_purity_derivative()     kernel.cpp:436-477 */
	} gple_real_scalars;
	int gple_train_real(gple_ctx* ctx, const double* X, const double* y, size_t N, const double theta[4], unsigned flags, gple_model** model, gple_real_scalars* out);

	/* Replaces TrainingComplexKernel::TrainingComplexKernel (gple/complex_kernel.cpp:221-592). */
	typedef struct gple_complex_scalars
	{
		double rescale;		/* complex_kernel.cpp:262 */
		double error;		/* complex_kernel.cpp:270-286 */
		double purity;		/* complex_kernel.cpp:357-377 */
		double magnitude;	/* complex_kernel.h:192-204 */
		double d_error[8];	/* complex_kernel.cpp:444-474 */
		double d_purity[8]; /* complex_kernel.cpp:475-590 */
	} gple_complex_scalars;
	int gple_train_complex(gple_ctx* ctx, const double* X, const double* y, size_t N, const double theta[8], unsigned flags, gple_model** model, gple_complex_scalars* out);

	/* Matrix / vector getters of a trained element.  `which`: */
	enum gple_model_field
	{
		GPLE_FIELD_INVERSE = 1,		  /* real: get_inverse() N x N (kernel.h:153) */
		GPLE_FIELD_INV_LABEL = 2,	  /* real: get_inverse_times_label() N; complex: upper part of augmented inverse label, N complex */
		GPLE_FIELD_LABEL = 3,		  /* rescaled label, N real / N complex */
		GPLE_FIELD_UPPER_LEFT = 4,	  /* complex: get_upper_left_block_of_augmented_inverse()  N x N complex (complex_kernel.h:208) */
		GPLE_FIELD_LOWER_LEFT = 5,	  /* complex: get_lower_left_block_of_augmented_inverse()  N x N complex (complex_kernel.h:215) */
		GPLE_FIELD_INV_LABEL_DERIV = 8 /* + parameter index: derivative of INV_LABEL over that parameter */
	};
	int gple_model_get(gple_ctx* ctx, const gple_model* model, int which, double* out);
	int gple_model_is_complex(const gple_model* model);
	/* Negative log marginal likelihood with LLT -- the objective of the reference's test programme
	 * (test/gpr.cpp:470-532, formula :475-496) -- of a trained element, as an alternative loss on the same kernels:
	 *   value = y'^T K^-1 y' / 2 + sum ln L_ii   (y' = the model's rescaled labels; the n/2 ln 2 pi constant is dropped),
	 *   grad[p] = tr[(K^-1 - b b^T) dK/dtheta_p] / 2, b = K^-1 y'   (4 or 8 doubles; NULL: value only).
	 * Complex element: the likelihood of the composite [Re f; Im f] process (complex_kernel.h:12-13), with the true
	 * derivatives (the reference's complex derivative arrays lack sigma^2, complex_kernel.cpp:37-51). */
	int gple_model_nlml(gple_ctx* ctx, gple_model* model, double* value, double* grad);
	size_t gple_model_size(const gple_model* model);
	int gple_model_destroy(gple_ctx* ctx, gple_model* model);

	/* ---- batched prediction --------------------------------------------------------------------------
	 * Replaces PredictiveKernel::PredictiveKernel (gple/kernel.cpp:481-544): raw prediction K* v (:495),
	 * variance k** - k K^-1 k^T (:496-518), cutoff prediction (:519, kernel.h:301-332) and, when yq != NULL,
	 * the squared validation error (:522) plus (if the model was trained with GPLE_CALC_DERIVATIVE and
	 * derr_out != NULL) its parameter gradient (:524-541).  Any output pointer may be NULL. */
	int gple_predict_real(gple_ctx* ctx, const gple_model* model, const double* Xq, size_t Q, const double* yq, double* pred_out, double* var_out, double* cutoff_out, double* err_out, double* derr_out);
	/* Replaces PredictiveComplexKernel::PredictiveComplexKernel (gple/complex_kernel.cpp:594-670).
	 * yq, pred_out, cutoff_out are complex (interleaved); var_out is real. */
	int gple_predict_complex(gple_ctx* ctx, const gple_model* model, const double* Xq, size_t Q, const double* yq, double* pred_out, double* var_out, double* cutoff_out, double* err_out, double* derr_out);

	/* Replaces loose_function (gple/opt.cpp:441-482): LOOCV error of the training set + squared error on the
	 * extra set (+ gradient when grad != NULL).  nparam = 4 (real kernel) or 8 (complex kernel).  The value is
	 * returned un-clamped; the caller applies make_normal. */
	/* The validation half of loose_function on an ALREADY trained model (opt.cpp:455-470): *error = |K* v - s y_e|^2 over the M extra
	 * points; grad (4 / 8 doubles, or NULL) its parameter gradient (model trained with GPLE_CALC_DERIVATIVE).  Lets the host share
	 * one trained model between the objective and the constraint callbacks that NLopt evaluates at the same parameters. */
	int gple_validation_error(gple_ctx* ctx, const gple_model* model, const double* Xe, const double* ye, size_t M, double* error, double* grad);
	int gple_loose_function(gple_ctx* ctx, const double* x, int nparam, double* grad, const double* X, const double* y, size_t N, const double* Xe, const double* ye, size_t M, double* value);

	/* ---- dynamics ------------------------------------------------------------------------------------
	 * Replaces adiabatic_potential / adiabatic_force / adiabatic_coupling (gple/pes.cpp:127-189) for n positions:
	 * E: 2 per point (E_0, E_1); F: 3 per point (F_00, F_10, F_11); D: 1 per point (d_10 = -d_01). */
	int gple_pes(gple_ctx* ctx, int pes_model, const double* x, size_t n, double* E, double* F, double* D);

	/* Replaces evolve() (gple/evolve.cpp:377-423) with the GPR-backed `predict_distribution` of
	 * gple/main.cpp:75-101 as the DistributionFunction: all points of the three elements
	 * (lower-triangular order rho00, rho10, rho11) are moved forward by dt and their densities are rebuilt
	 * from the 9 backward-propagated predictions per point (evolve.cpp:184-372).  A NULL model means the
	 * element has no predictor (predicts 0, main.cpp:85-99).  Points are updated in place. */
	int gple_evolve(gple_ctx* ctx, int pes_model, const gple_model* m00, const gple_model* m10, const gple_model* m11, double* pts00, size_t n00, double* pts10, size_t n10, double* pts11, size_t n11, double mass, double dt);

	/* ---- multi-GPU: one process per GPU, one NCCL communicator per context ----------------
	 * The path shards over the evolved points; a factorisation never spans GPUs.  The reference's tick
	 * (gple/main.cpp:140-141, 176) evolves every point and then rebuilds the element models from ALL evolved points
	 * (predict.cpp:246-280), so the one exchange per step is the all-gather of the evolved (r, rho) sets, after which
	 * every rank holds the full sets and rebuilds its models redundantly (BASELINE.json north_star).
	 * Bootstrap: rank 0 calls gple_comm_unique_id and hands the 128 bytes to the other ranks by whatever channel the host
	 * has (MPI, a torch.distributed store, a file); every rank then calls gple_ctx_comm_init on its context.
	 * NCCL is bound at run time (libnccl.so.2); without it these entry points return GPLE_ERR_COMM and nothing else changes. */
#define GPLE_COMM_ID_BYTES 128
	int gple_comm_unique_id(unsigned char id[GPLE_COMM_ID_BYTES]);
	int gple_ctx_comm_init(gple_ctx* ctx, int rank, int nranks, const unsigned char id[GPLE_COMM_ID_BYTES]);
	/* rank / size of the context's communicator; (0, 1) for a context without one */
	int gple_ctx_comm_info(const gple_ctx* ctx, int* rank, int* nranks);
	/* the block [lo, hi) of `total` records that `rank` owns (blocks differ by at most one record) */
	int gple_partition(size_t total, int rank, int nranks, size_t* lo, size_t* hi);
	/* In-place all-gather of a block-partitioned point set: pts = the FULL array of `total` PhaseSpacePoints (host or
	 * device) of which this rank's block is valid on entry; all blocks are valid on return.  ncclAllGather on the context
	 * stream (grouped ncclBroadcast when the blocks are uneven).  A no-op on a context without a communicator. */
	int gple_allgather_points(gple_ctx* ctx, double* pts, size_t total);
	/* sum over the ranks of `count` doubles, in place (partial sums of gple_observables / gple_validation_error over sharded points) */
	int gple_allreduce_sum(gple_ctx* ctx, double* values, size_t count);
	/* Element models are independent (gple/predict.cpp:390-393 builds them in a parallel loop): on G GPUs each
	 * model of the TrainingKernels rebuild is trained on ONE rank (a factorisation never spans GPUs) and replicated with this
	 * call -- collective over the communicator.  On `root`, *model is the trained model (NULL: element not populated); on the
	 * other ranks *model receives a copy that lives in this context's pool (or NULL), to be freed with gple_model_destroy.
	 * Replicates what prediction and the scalar getters need (X, L^-1, K^-1 y', labels, diag K^-1); the on-demand full
	 * inverse and the derivative arrays stay with the root. */
	int gple_model_bcast(gple_ctx* ctx, gple_model** model, int root);
	/* gple_evolve over point sets that are block-partitioned over the ranks: pts?? are the FULL sets (n?? = their full sizes);
	 * every rank evolves its own block of each element and the blocks are all-gathered in place, so that every rank returns
	 * with the full evolved sets -- evolve(density) of main.cpp:140 on G GPUs.  Identical to gple_evolve without a communicator. */
	int gple_evolve_sharded(gple_ctx* ctx, int pes_model, const gple_model* m00, const gple_model* m10, const gple_model* m11, double* pts00, size_t n00, double* pts10, size_t n10, double* pts11, size_t n11, double mass, double dt);

	/* Replaces new_point_predict() (gple/evolve.cpp:425-443) for n phase points r of element (row, col):
	 * out = n complex densities. */
	int gple_new_point_predict(gple_ctx* ctx, int pes_model, const gple_model* m00, const gple_model* m10, const gple_model* m11, const double* r, size_t n, int row, int col, double mass, double dt, double* out);

	/* Replaces the transform_reduce observables of gple/predict.cpp:43-244 for one element.
	 * out[9] = sum Re rho, sum x Re rho, sum p Re rho, sum x, sum p, sum x^2, sum p^2,
	 *          sum (p^2/2m + E_pes(x)) Re rho, sum |rho|^2. */
	int gple_observables(gple_ctx* ctx, int pes_model, const double* pts, size_t n, double mass, int pes_index, double out[9]);

	/* ---- Metropolis sampling (gple/mc.cpp:125-403) ------------------------------------------------------
	 * Replaces generate_markov_chain (mc.cpp:143-188) called once per point under par_unseq by element_monte_carlo
	 * (:339-378), acceptance_optimize_displacement (:288-337) and autocorrelation_optimize_steps (:197-279): all n chains
	 * of one element advance in lock-step, one batched density evaluation per step.  Target density
	 * distribution(r, row, col) of the reference's DistributionFunction:
	 *   GPLE_MC_ANALYTIC   initial_distribution (mc.cpp:30-50), analytic = {x0, p0, sigma_x, sigma_p, pop0, pop1, phase0, phase1}
	 *   GPLE_MC_PREDICT    predict_distribution (main.cpp:75-101): cutoff prediction of the element's own model (0 if absent)
	 *   GPLE_MC_NEW_POINT  new_point_predict (evolve.cpp:425-443) over the three models (pes_model, mass, dt)
	 * The reference's shared clock-seeded mt19937 (mc.cpp:17) is replaced by one Philox4x32-10 stream per chain:
	 * key = seed, counter = (chain0 + k, step, stream, block), so results do not depend on scheduling or sharding. */
	enum gple_mc_kind
	{
		GPLE_MC_ANALYTIC = 0,
		GPLE_MC_PREDICT = 1,
		GPLE_MC_NEW_POINT = 2
	};
	typedef struct gple_mc_source
	{
		int kind;
		int row, col; /* lower-triangular element index: (0,0), (1,0), (1,1) */
		double analytic[8];
		const gple_model *m00, *m10, *m11; /* NULL = element absent */
		int pes_model;
		double mass, dt;
	} gple_mc_source;
	/* pts: n x 4 (x, p, Re rho, Im rho), start points in, last state of every chain and its density out (mc.cpp:366-369).
	 * accept_ratio: n doubles or NULL.  chains: n x (num_steps + 1) x 2 doubles (every state of every chain) or NULL. */
	int gple_markov_chains(gple_ctx* ctx, const gple_mc_source* source, double* pts, size_t n, size_t num_steps, double max_displacement, unsigned long long seed, unsigned long long stream, unsigned long long chain0, double* accept_ratio, double* chains);
	/* Mean autocorrelation of the chains (mc.cpp:230-243): out[j] = mean_k sum_i (r_i - <r>).(r_{i+j} - <r>) / (len - j), j < len / 2 */
	int gple_chain_autocorrelation(gple_ctx* ctx, const double* chains, size_t n, size_t len, double* out);

	/* ---- measurement helpers (bench.py) --------------------------------------------------------------- */
	/* Per-kernel CUDA-event timing on the launching stream.  While enabled, every launch of the kernels below is
	 * bracketed by an event pair; gple_profile_read synchronises, sums the elapsed times since the last read and
	 * resets.  `work` is the executed work of those launches: FP64 flops (DMMA kernels) or HBM bytes (kernel build). */
	enum gple_profile_slot
	{
		GPLE_PROF_VARIANCE_GEMM = 0, /* var_gemm_kernel: Z = K* W^T with fused row sum of squares; flops */
		GPLE_PROF_KERNEL_BUILD = 1,	 /* kstar_kernel: K* rows + fused mean; bytes written */
		GPLE_PROF_FACTORISE = 2,	 /* potrf + trtri (all their launches together); flops = 2 n^3 / 3 */
		GPLE_PROF_MEAN = 3			 /* kmean_kernel: mean K* v without storing K*; work = exp evaluations */
	};
	int gple_profile_enable(gple_ctx* ctx, int on);
	int gple_profile_read(gple_ctx* ctx, int slot, double* total_ms, unsigned long long* launches, double* work);
	/* Tile-configuration tuning of the variance GEMM: time `iters` launches of variant 0..6 on zero operands of
	 * `rows` x n (both multiples of 128); gple_set_variance_gemm_variant selects the variant used by predictions. */
	int gple_tune_variance_gemm(gple_ctx* ctx, int variant, int rows, int n, int iters, double* ms_per_launch);
	int gple_set_variance_gemm_variant(int variant);
	/* Factorisation schedule: blocks of at most `n` rows are factorised by a right-looking sweep over 128-blocks (short
	 * K = 128 GEMMs, latency-optimal), larger ones recursively (long-K GEMMs, throughput-optimal).  Returns the old value. */
	int gple_set_potrf_flat(int n);
	/* Register-resident DMMA / DFMA loops: measured FP64 tensor and vector peaks of this GPU, in TFLOP/s. */
	int gple_measure_fp64_peak(gple_ctx* ctx, double* dmma_tflops, double* dfma_tflops);
	/* DMMA rate of the GEMM's own register tile (8 warps / SM, 32 accumulators, 8 + 4 changing operands, no memory):
	 * the ceiling of a register-tiled mma.sync FP64 kernel at the occupancy of var_gemm_kernel. */
	int gple_measure_dmma_tile_peak(gple_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* GPLE_B200_H */
