// gple_host.hpp -- C++ host mirror of the reference's hot-path headers over the C-ABI (include/gple_b200.h).
//
// Same class names, constructor arguments and getter semantics as the reference
// (kaigu1997/gaussian_process_liouville_equation, gaussian_process_liouville_equation/ = gple/):
//   gple/kernel.h          TrainingKernel (:111-280), PredictiveKernel (:336-403)
//   gple/complex_kernel.h  TrainingComplexKernel (:150-318), PredictiveComplexKernel (:323-391)
//   gple/predict.h         TrainingKernels (:89-143), calculate_* observables (:19-72)
//   gple/evolve.h          evolve (:16-21)      gple/pes.h  adiabatic_potential / force / coupling (:47-59)
//   gple/storage.h         PhaseSpacePoint (:232-297), ElementPoints / AllPoints (:327-329)
// Eigen is not available, so containers are plain std types with the reference's memory layout
// (PhasePoints = 2 x n column-major doubles; matrices column-major).  Objects are thin handles: the data lives
// on the GPU, scalars are copied back eagerly, matrices lazily.  Header-only; link with libgple_b200.so.
#pragma once
#include "../../include/gple_b200.h"

#include <cstdint>
#include <algorithm>
#include <array>
#include <cassert>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <complex>
#include <condition_variable>
#include <functional>
#include <thread>
#include <limits>
#include <atomic>
#include <memory>
#include <mutex>
#include <optional>
#include <ostream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace gple_host
{
constexpr std::size_t NumPES = 2, PhaseDim = 2, NumElements = 3; // gple/stdafx.h:111-121 (lower-triangular elements)

/// gple/stdafx.h:153 -- 2 x n, column-major: (x_0, p_0, x_1, p_1, ...)
struct PhasePoints
{
	std::vector<double> d;
	PhasePoints() = default;
	explicit PhasePoints(std::size_t n): d(2 * n, 0.0) {}
	std::size_t cols() const { return d.size() / 2; }
	double& operator()(std::size_t row, std::size_t col) { return d[2 * col + row]; }
	double operator()(std::size_t row, std::size_t col) const { return d[2 * col + row]; }
	const double* data() const { return d.data(); }
};
using ParameterVector = std::vector<double>;												 // gple/kernel.h:10
using ElementTrainingSet = std::tuple<PhasePoints, std::vector<std::complex<double>>>;		 // gple/kernel.h:14
using ClassicalPhaseVector = std::array<double, PhaseDim>;									 // gple/stdafx.h:147

/// gple/storage.h:232-297 (32 bytes, same layout)
struct PhaseSpacePoint
{
	ClassicalPhaseVector r;
	std::complex<double> rho;
};
static_assert(sizeof(PhaseSpacePoint) == 32, "PhaseSpacePoint must stay the reference's 32-byte AoS");
using ElementPoints = std::vector<PhaseSpacePoint>;			 // gple/storage.h:327
using AllPoints = std::array<ElementPoints, NumElements>;	 // lower-triangular order rho00, rho10, rho11

/// One library context (CUDA stream + workspace) per worker slot.  The C-ABI is thread-safe across contexts, so the
/// independent element models of a step / of a loss evaluation are built concurrently, each on the context of its slot
/// (ContextSlot sets the slot of the calling thread).  Slot 0 is the main context.  The slots come in GROUPS of three (one per
/// element): group 0 is the caller's, groups 1 and 2 serve the restart stages that Optimization::optimize runs ahead of time
/// on threads of their own (ContextGroup).
class Context
{
public:
	static constexpr int NumGroups = 3, SlotsPerGroup = 3, NumSlots = NumGroups * SlotsPerGroup;
	static int& slot()
	{
		static thread_local int s = 0;
		return s;
	}
	static int& group()
	{
		static thread_local int g = 0;
		return g;
	}
	/// RAII: slot 0 for the calling thread within a scope
	struct ContextSlot0
	{
		int previous;
		ContextSlot0(): previous(slot()) { slot() = 0; }
		~ContextSlot0() { slot() = previous; }
	};
	/// device of this process (one process per GPU): set before the first use of the library, default 0
	static int& device()
	{
		static int d = 0;
		return d;
	}
	static gple_ctx* get()
	{
		static Context c(device());
		return c.at(slot());
	}
	/// Multi-GPU (one process per GPU): gives the main context (slot 0) its NCCL communicator.  Rank 0 draws
	/// the unique id and publishes it in `id_file` (any path all ranks see: a shared directory of the job); the others
	/// wait for it.  A launcher that has its own channel (MPI_Bcast) can call gple_comm_unique_id / gple_ctx_comm_init itself.
	/// After this, evolve() shards the points over the ranks and all-gathers the evolved sets (gple_evolve_sharded).
	static void init_distributed(const int rank, const int nranks, const std::string& id_file, const int local_device)
	{
		device() = local_device;
		unsigned char id[GPLE_COMM_ID_BYTES];
		if (rank == 0)
		{
			check(gple_comm_unique_id(id), "gple_comm_unique_id");
			const std::string tmp = id_file + ".tmp";
			std::FILE* f = std::fopen(tmp.c_str(), "wb");
			if (f == nullptr || std::fwrite(id, 1, sizeof(id), f) != sizeof(id))
			{
				throw std::runtime_error("init_distributed: cannot write " + tmp);
			}
			std::fclose(f);
			std::rename(tmp.c_str(), id_file.c_str()); // atomic: a reader never sees a partial id
		}
		else
		{
			for (int tries = 0;; tries++)
			{
				std::FILE* f = std::fopen(id_file.c_str(), "rb");
				if (f != nullptr)
				{
					const std::size_t got = std::fread(id, 1, sizeof(id), f);
					std::fclose(f);
					if (got == sizeof(id))
					{
						break;
					}
				}
				if (tries > 6000)
				{
					throw std::runtime_error("init_distributed: no communicator id in " + id_file + " after 60 s");
				}
				std::this_thread::sleep_for(std::chrono::milliseconds(10));
			}
		}
		const ContextSlot0 main_slot;
		check(gple_ctx_comm_init(get(), rank, nranks, id), "gple_ctx_comm_init");
	}
	static int rank()
	{
		int r = 0;
		gple_ctx_comm_info(main(), &r, nullptr);
		return r;
	}
	static int num_ranks()
	{
		int n = 1;
		gple_ctx_comm_info(main(), nullptr, &n);
		return n;
	}
	/// the context of slot 0, whatever the calling thread's slot (it owns the communicator)
	static gple_ctx* main()
	{
		const ContextSlot0 main_slot;
		return get();
	}
	static void check(int rc, const char* where, bool allow_not_spd = false)
	{
		if (rc == GPLE_OK || (allow_not_spd && rc == GPLE_ERR_NOT_SPD))
		{
			return;
		}
		throw std::runtime_error(std::string(where) + ": gple status " + std::to_string(rc) + ": " + gple_last_error(get()));
	}

private:
	gple_ctx* ctx[NumSlots] = {};
	int dev;
	std::mutex create_mutex;
	std::atomic<bool> ready[NumGroups] = {};
	explicit Context(int device): dev(device) { create_group(0); }
	/// the three contexts of a slot group come into being when the group is first used
	void create_group(const int g)
	{
		const std::lock_guard<std::mutex> lock(create_mutex);
		if (ready[g].load(std::memory_order_acquire))
		{
			return;
		}
		for (int k = 0; k < SlotsPerGroup; k++)
		{
			if (gple_ctx_create(dev, &ctx[g * SlotsPerGroup + k]) != GPLE_OK)
			{
				throw std::runtime_error("gple_ctx_create failed: no CUDA device (there is no CPU fallback)");
			}
		}
		ready[g].store(true, std::memory_order_release);
	}
	gple_ctx* at(const int s)
	{
		const int g = s / SlotsPerGroup;
		if (!ready[g].load(std::memory_order_acquire))
		{
			create_group(g);
		}
		return ctx[s];
	}
	~Context()
	{
		for (auto& c : ctx)
		{
			if (c != nullptr)
			{
				gple_ctx_destroy(c);
			}
		}
	}
};

/// RAII: the calling thread uses the context of slot k OF ITS GROUP until the guard goes out of scope
struct ContextSlot
{
	int previous;
	explicit ContextSlot(const int k): previous(Context::slot()) { Context::slot() = Context::group() * Context::SlotsPerGroup + k % Context::SlotsPerGroup; }
	~ContextSlot() { Context::slot() = previous; }
	ContextSlot(const ContextSlot&) = delete;
	ContextSlot& operator=(const ContextSlot&) = delete;
};
/// RAII: the calling thread works in slot group g (its own three contexts, element workers and model-cache entries)
struct ContextGroup
{
	int previous_group, previous_slot;
	explicit ContextGroup(const int g): previous_group(Context::group()), previous_slot(Context::slot())
	{
		Context::group() = g % Context::NumGroups;
		Context::slot() = Context::group() * Context::SlotsPerGroup;
	}
	~ContextGroup()
	{
		Context::group() = previous_group;
		Context::slot() = previous_slot;
	}
	ContextGroup(const ContextGroup&) = delete;
	ContextGroup& operator=(const ContextGroup&) = delete;
};

/// Multi-GPU: the element models are independent, so with G ranks element e is trained / optimised on ONE
/// rank: the off-diagonal (complex, order 2N: the most expensive) element on rank 0, the diagonal ones on ranks 1 and 2
/// (both on rank 1 when G = 2, everything on rank 0 when G = 1).
inline int element_owner(const std::size_t element, const int num_ranks)
{
	const int owner[3] = {1 % num_ranks, 0, 2 % num_ranks};
	return owner[element];
}
/// In-place sum over the ranks of a small host vector (every rank contributes zeros for what it does not own)
inline void all_reduce_sum(std::vector<double>& v)
{
	if (Context::num_ranks() > 1 && !v.empty())
	{
		Context::check(gple_allreduce_sum(Context::main(), v.data(), v.size()), "all_reduce_sum");
	}
}

/// Two persistent worker threads per slot group, bound to the group's context slots 1 and 2 (slot 0 is the calling thread).  An
/// optimisation makes thousands of three-element evaluations of a millisecond or less each: spawning threads per evaluation
/// would cost as much as the evaluation itself.
class ElementWorkers
{
public:
	/// the workers of the calling thread's group (created on first use)
	static ElementWorkers& instance()
	{
		static std::mutex m;
		static std::array<std::unique_ptr<ElementWorkers>, Context::NumGroups> w;
		const int g = Context::group();
		const std::lock_guard<std::mutex> lock(m);
		if (!w[g])
		{
			w[g].reset(new ElementWorkers(g));
		}
		return *w[g];
	}
	static bool& inside_worker()
	{
		static thread_local bool flag = false;
		return flag;
	}
	void post(const int k, std::function<void()> job)
	{
		Worker& w = Workers[k - 1];
		{
			const std::lock_guard<std::mutex> lock(w.m);
			w.job = std::move(job);
			w.busy = true;
			w.error = nullptr;
		}
		w.cv.notify_all();
	}
	/// wait for worker k; re-throws what its job threw
	void wait(const int k)
	{
		Worker& w = Workers[k - 1];
		std::unique_lock<std::mutex> lock(w.m);
		w.cv.wait(lock, [&w]() { return !w.busy; });
		if (w.error)
		{
			std::rethrow_exception(w.error);
		}
	}

private:
	struct Worker
	{
		std::mutex m;
		std::condition_variable cv;
		std::function<void()> job;
		bool busy = false, stop = false;
		std::exception_ptr error;
		std::thread th;
	};
	std::array<Worker, 2> Workers;
	explicit ElementWorkers(const int group)
	{
		for (int k = 1; k <= 2; k++)
		{
			Worker& w = Workers[k - 1];
			w.th = std::thread(
				[&w, k, group]()
				{
					const ContextGroup in_group(group);
					const ContextSlot guard(k);
					inside_worker() = true;
					std::unique_lock<std::mutex> lock(w.m);
					while (true)
					{
						w.cv.wait(lock, [&w]() { return w.busy || w.stop; });
						if (w.stop)
						{
							return;
						}
						std::function<void()> job = std::move(w.job);
						lock.unlock();
						std::exception_ptr err;
						try
						{
							job();
						}
						catch (...)
						{
							err = std::current_exception();
						}
						lock.lock();
						w.error = err;
						w.busy = false;
						w.cv.notify_all();
					}
				}
			);
		}
	}

public:
	~ElementWorkers()
	{
		for (Worker& w : Workers)
		{
			{
				const std::lock_guard<std::mutex> lock(w.m);
				w.stop = true;
			}
			w.cv.notify_all();
			if (w.th.joinable())
			{
				w.th.join();
			}
		}
	}
};

/// Run f(0), f(1), f(2) concurrently, f(k) on the context of slot k; exceptions are re-thrown in the caller.
template <typename F>
inline void for_each_element_concurrently(F&& f)
{
	if (ElementWorkers::inside_worker())
	{
		// already on a worker (no caller in this library nests, but a nested call must not wait for its own thread)
		for (std::size_t k = 0; k < 3; k++)
		{
			f(k);
		}
		return;
	}
	Context::get(); // the contexts exist before the workers use them
	ElementWorkers& workers = ElementWorkers::instance();
	workers.post(1, [&f]() { f(std::size_t(1)); });
	workers.post(2, [&f]() { f(std::size_t(2)); });
	std::exception_ptr mine;
	try
	{
		f(std::size_t(0));
	}
	catch (...)
	{
		mine = std::current_exception();
	}
	std::exception_ptr theirs;
	for (int k = 1; k <= 2; k++)
	{
		try
		{
			workers.wait(k);
		}
		catch (...)
		{
			theirs = std::current_exception();
		}
	}
	if (mine)
	{
		std::rethrow_exception(mine);
	}
	if (theirs)
	{
		std::rethrow_exception(theirs);
	}
}

struct ModelDeleter
{
	void operator()(gple_model* m) const { gple_model_destroy(Context::get(), m); }
};
using ModelHandle = std::unique_ptr<gple_model, ModelDeleter>;

/// gple/opt.cpp:420-431
inline void make_normal(double& d)
{
	if (!(d == d) || d == std::numeric_limits<double>::infinity() || d == -std::numeric_limits<double>::infinity())
	{
		d = std::numeric_limits<double>::max();
	}
}

/// Column-major dense matrix with the layout of the reference's Eigen::MatrixXd / MatrixXcd (Eigen is not available here)
template <typename T>
struct Matrix
{
	std::size_t Rows = 0, Cols = 0;
	std::vector<T> d;
	Matrix() = default;
	Matrix(std::size_t r, std::size_t c): Rows(r), Cols(c), d(r * c) {}
	std::size_t rows() const { return Rows; }
	std::size_t cols() const { return Cols; }
	T& operator()(std::size_t r, std::size_t c) { return d[c * Rows + r]; }
	const T& operator()(std::size_t r, std::size_t c) const { return d[c * Rows + r]; }
	T* data() { return d.data(); }
	const T* data() const { return d.data(); }
};
using MatrixXd = Matrix<double>;
using MatrixXcd = Matrix<std::complex<double>>;

/// gple/kernel.h:29-106 -- kernel matrix of two point sets and, optionally, its derivatives over (sigma_f, l_x, l_p, sigma_n).
/// As in the reference (kernel.cpp:217-242, :8-31) the two sets count as "the same set" when they are the same buffer.
class KernelBase
{
public:
	static constexpr std::size_t NumTotalParameters = 1 + PhaseDim + 1;
	static constexpr double RescaleMaximum = 10.0;
	using KernelParameter = std::tuple<double, ClassicalPhaseVector, double>; // magnitude, characteristic lengths, noise
	template <typename T>
	using ParameterArray = std::array<T, NumTotalParameters>;

	KernelBase(const KernelParameter& Parameter, const PhasePoints& left_feature, const PhasePoints& right_feature, const bool IsToCalculateDerivative):
		KernelParams(Parameter), LeftFeature(left_feature), RightFeature(right_feature), KernelMatrix(left_feature.cols(), right_feature.cols())
	{
		const auto& [mag, l, noise] = KernelParams;
		const double theta[4] = {mag, l[0], l[1], noise};
		const std::size_t nL = left_feature.cols(), nR = right_feature.cols();
		std::vector<double> dk(IsToCalculateDerivative ? 4 * nL * nR : 0);
		Context::check(gple_kernel_real(Context::get(), left_feature.data(), nL, right_feature.data(), nR, theta, left_feature.data() == right_feature.data() ? 1 : 0, KernelMatrix.data(), IsToCalculateDerivative ? dk.data() : nullptr), "KernelBase");
		if (IsToCalculateDerivative)
		{
			Derivatives.emplace();
			for (std::size_t p = 0; p < NumTotalParameters; p++)
			{
				(*Derivatives)[p] = MatrixXd(nL, nR);
				std::copy(dk.cbegin() + p * nL * nR, dk.cbegin() + (p + 1) * nL * nR, (*Derivatives)[p].d.begin());
			}
		}
	}
	const KernelParameter& get_formatted_parameters() const { return KernelParams; }
	const PhasePoints& get_left_feature() const { return LeftFeature; }
	const PhasePoints& get_right_feature() const { return RightFeature; }
	const MatrixXd& get_kernel() const { return KernelMatrix; }
	const ParameterArray<MatrixXd>& get_derivative() const
	{
		assert(Derivatives.has_value());
		return Derivatives.value();
	}

private:
	const KernelParameter KernelParams;
	const PhasePoints LeftFeature;
	const PhasePoints RightFeature;
	MatrixXd KernelMatrix;
	std::optional<ParameterArray<MatrixXd>> Derivatives;
};

/// gple/complex_kernel.h:14-145 -- covariance K and pseudo-covariance K~ of the widely-linear complex process and their
/// derivative arrays over (sigma, sigma_R, l_Rx, l_Rp, sigma_I, l_Ix, l_Ip, sigma_n) (complex_kernel.cpp:20-59, 74-132)
class ComplexKernelBase
{
public:
	static constexpr std::size_t NumTotalParameters = 1 + 2 * (1 + PhaseDim) + 1;
	template <typename T>
	using ParameterArray = std::array<T, NumTotalParameters>;

	ComplexKernelBase(const ParameterVector& Parameter, const PhasePoints& left_feature, const PhasePoints& right_feature, const bool IsToCalculateDerivative):
		Params(Parameter), KernelMatrix(left_feature.cols(), right_feature.cols()), PseudoKernelMatrix(left_feature.cols(), right_feature.cols())
	{
		assert(Parameter.size() == NumTotalParameters);
		const std::size_t nL = left_feature.cols(), nR = right_feature.cols();
		const int same = left_feature.data() == right_feature.data() ? 1 : 0;
		Context::check(gple_kernel_complex(Context::get(), left_feature.data(), nL, right_feature.data(), nR, Params.data(), same, KernelMatrix.data(), reinterpret_cast<double*>(PseudoKernelMatrix.data())), "ComplexKernelBase");
		if (IsToCalculateDerivative)
		{
			std::vector<double> dk(8 * nL * nR);
			std::vector<std::complex<double>> dkt(8 * nL * nR);
			Context::check(gple_kernel_complex_derivatives(Context::get(), left_feature.data(), nL, right_feature.data(), nR, Params.data(), same, dk.data(), reinterpret_cast<double*>(dkt.data())), "ComplexKernelBase derivatives");
			Derivatives.emplace();
			PseudoDerivatives.emplace();
			for (std::size_t p = 0; p < NumTotalParameters; p++)
			{
				(*Derivatives)[p] = MatrixXd(nL, nR);
				(*PseudoDerivatives)[p] = MatrixXcd(nL, nR);
				std::copy(dk.cbegin() + p * nL * nR, dk.cbegin() + (p + 1) * nL * nR, (*Derivatives)[p].d.begin());
				std::copy(dkt.cbegin() + p * nL * nR, dkt.cbegin() + (p + 1) * nL * nR, (*PseudoDerivatives)[p].d.begin());
			}
		}
	}
	const ParameterVector& get_parameters() const { return Params; }
	const MatrixXd& get_kernel() const { return KernelMatrix; }
	const MatrixXcd& get_pseudo_kernel() const { return PseudoKernelMatrix; }
	const ParameterArray<MatrixXd>& get_derivatives() const
	{
		assert(Derivatives.has_value());
		return Derivatives.value();
	}
	const ParameterArray<MatrixXcd>& get_pseudo_derivatives() const
	{
		assert(PseudoDerivatives.has_value());
		return PseudoDerivatives.value();
	}

private:
	const ParameterVector Params;
	MatrixXd KernelMatrix;
	MatrixXcd PseudoKernelMatrix;
	std::optional<ParameterArray<MatrixXd>> Derivatives;
	std::optional<ParameterArray<MatrixXcd>> PseudoDerivatives;
};

/// gple/kernel.h:111-280
class TrainingKernel
{
public:
	static constexpr std::size_t NumTotalParameters = 4;
	template <typename T>
	using ParameterArray = std::array<T, NumTotalParameters>;

	TrainingKernel(const ParameterVector& Parameter, const ElementTrainingSet& TrainingSet, bool IsToCalculateError, bool IsToCalculateAverage, bool IsToCalculateDerivative):
		Params(Parameter), Feature(std::get<0>(TrainingSet)), N(std::get<0>(TrainingSet).cols()),
		Flags((IsToCalculateError ? unsigned(GPLE_CALC_ERROR) : 0u) | (IsToCalculateAverage ? unsigned(GPLE_CALC_AVERAGE) : 0u) | (IsToCalculateDerivative ? unsigned(GPLE_CALC_DERIVATIVE) : 0u))
	{
		assert(Parameter.size() == NumTotalParameters);
		gple_model* m = nullptr;
		Status = gple_train_real(Context::get(), Feature.data(), reinterpret_cast<const double*>(std::get<1>(TrainingSet).data()), N, Params.data(), Flags, &m, &S);
		Context::check(Status, "TrainingKernel", true);
		Handle.reset(m);
	}
	const ParameterVector& get_parameters() const { return Params; }
	const PhasePoints& get_left_feature() const { return Feature; }
	double get_rescale_factor() const { return S.rescale; }
	double get_magnitude() const { return S.magnitude; }
	/// NLML / LLT objective of test/gpr.cpp:470-532 on this model; Gradient (4 entries) is filled when non-null
	double get_negative_log_marginal_likelihood(ParameterArray<double>* Gradient = nullptr) const
	{
		double v = 0.0;
		Context::check(gple_model_nlml(Context::get(), Handle.get(), &v, Gradient != nullptr ? Gradient->data() : nullptr), "nlml");
		return v;
	}
	double get_error() const { assert(Flags & GPLE_CALC_ERROR); return S.error; }
	double get_population() const { assert(Flags & GPLE_CALC_AVERAGE); return S.population; }
	ClassicalPhaseVector get_1st_order_average() const { assert(Flags & GPLE_CALC_AVERAGE); return {S.first_order[0], S.first_order[1]}; }
	double get_purity() const { assert(Flags & GPLE_CALC_AVERAGE); return S.purity; }
	ParameterArray<double> get_error_derivative() const { return arr(S.d_error); }
	ParameterArray<double> get_population_derivative() const { return arr(S.d_population); }
	ParameterArray<double> get_purity_derivative() const { return arr(S.d_purity); }
	/// column-major N x N
	std::vector<double> get_inverse() const { return field(GPLE_FIELD_INVERSE, N * N); }
	std::vector<double> get_inverse_times_label() const { return field(GPLE_FIELD_INV_LABEL, N); }
	const gple_model* handle() const { return Handle.get(); }
	int status() const { return Status; }
	std::size_t size() const { return N; }

private:
	ParameterVector Params;
	PhasePoints Feature;
	std::size_t N;
	unsigned Flags;
	int Status = GPLE_OK;
	gple_real_scalars S{};
	ModelHandle Handle;
	static ParameterArray<double> arr(const double* p) { return {p[0], p[1], p[2], p[3]}; }
	std::vector<double> field(int which, std::size_t count) const
	{
		std::vector<double> out(count);
		Context::check(gple_model_get(Context::get(), Handle.get(), which, out.data()), "gple_model_get");
		return out;
	}
};

/// gple/kernel.h:336-403
class PredictiveKernel
{
public:
	PredictiveKernel(const PhasePoints& TestFeature, const TrainingKernel& kernel, bool IsToCalculateDerivative, const std::optional<std::vector<double>>& TestLabel = std::nullopt):
		Variance(TestFeature.cols()), Cutoff(TestFeature.cols()), Prediction(TestFeature.cols())
	{
		const bool grad = IsToCalculateDerivative && TestLabel.has_value();
		Context::check(
			gple_predict_real(Context::get(), kernel.handle(), TestFeature.data(), TestFeature.cols(), TestLabel ? TestLabel->data() : nullptr, Prediction.data(), Variance.data(), Cutoff.data(), TestLabel ? &Error : nullptr, grad ? ErrorDerivatives.data() : nullptr),
			"PredictiveKernel"
		);
	}
	const std::vector<double>& get_variance() const { return Variance; }
	const std::vector<double>& get_cutoff_prediction() const { return Cutoff; }
	const std::vector<double>& get_prediction() const { return Prediction; }
	double get_error() const { return Error; }
	const std::array<double, 4>& get_error_derivative() const { return ErrorDerivatives; }

private:
	std::vector<double> Variance, Cutoff, Prediction;
	double Error = std::numeric_limits<double>::quiet_NaN();
	std::array<double, 4> ErrorDerivatives{};
};

/// gple/complex_kernel.h:150-318
class TrainingComplexKernel
{
public:
	static constexpr std::size_t NumTotalParameters = 8;
	template <typename T>
	using ParameterArray = std::array<T, NumTotalParameters>;
	TrainingComplexKernel(const ParameterVector& Parameter, const ElementTrainingSet& TrainingSet, bool IsToCalculateError, bool IsToCalculateAverage, bool IsToCalculateDerivative):
		Params(Parameter), N(std::get<0>(TrainingSet).cols()),
		Flags((IsToCalculateError ? unsigned(GPLE_CALC_ERROR) : 0u) | (IsToCalculateAverage ? unsigned(GPLE_CALC_AVERAGE) : 0u) | (IsToCalculateDerivative ? unsigned(GPLE_CALC_DERIVATIVE) : 0u))
	{
		assert(Parameter.size() == NumTotalParameters);
		gple_model* m = nullptr;
		Status = gple_train_complex(Context::get(), std::get<0>(TrainingSet).data(), reinterpret_cast<const double*>(std::get<1>(TrainingSet).data()), N, Params.data(), Flags, &m, &S);
		Context::check(Status, "TrainingComplexKernel", true);
		Handle.reset(m);
	}
	const ParameterVector& get_parameters() const { return Params; }
	double get_rescale_factor() const { return S.rescale; }
	double get_magnitude() const { return S.magnitude; }
	/// NLML / LLT objective of test/gpr.cpp:470-532 on this model; Gradient (8 entries) is filled when non-null
	double get_negative_log_marginal_likelihood(ParameterArray<double>* Gradient = nullptr) const
	{
		double v = 0.0;
		Context::check(gple_model_nlml(Context::get(), Handle.get(), &v, Gradient != nullptr ? Gradient->data() : nullptr), "nlml");
		return v;
	}
	double get_error() const { assert(Flags & GPLE_CALC_ERROR); return S.error; }
	double get_purity() const { assert(Flags & GPLE_CALC_AVERAGE); return S.purity; }
	ParameterArray<double> get_error_derivative() const { return arr(S.d_error); }
	ParameterArray<double> get_purity_derivative() const { return arr(S.d_purity); }
	std::vector<std::complex<double>> get_upper_left_block_of_augmented_inverse() const { return cfield(GPLE_FIELD_UPPER_LEFT, N * N); }
	std::vector<std::complex<double>> get_lower_left_block_of_augmented_inverse() const { return cfield(GPLE_FIELD_LOWER_LEFT, N * N); }
	std::vector<std::complex<double>> get_upper_part_of_augmented_inverse_times_label() const { return cfield(GPLE_FIELD_INV_LABEL, N); }
	const gple_model* handle() const { return Handle.get(); }
	int status() const { return Status; }

private:
	ParameterVector Params;
	std::size_t N;
	unsigned Flags;
	int Status = GPLE_OK;
	gple_complex_scalars S{};
	ModelHandle Handle;
	static ParameterArray<double> arr(const double* p) { return {p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]}; }
	std::vector<std::complex<double>> cfield(int which, std::size_t count) const
	{
		std::vector<std::complex<double>> out(count);
		Context::check(gple_model_get(Context::get(), Handle.get(), which, reinterpret_cast<double*>(out.data())), "gple_model_get");
		return out;
	}
};

/// gple/complex_kernel.h:323-391
class PredictiveComplexKernel
{
public:
	PredictiveComplexKernel(const PhasePoints& TestFeature, const TrainingComplexKernel& kernel, bool /*IsToCalculateDerivative*/, const std::optional<std::vector<std::complex<double>>>& TestLabel = std::nullopt):
		Variance(TestFeature.cols()), Cutoff(TestFeature.cols()), Prediction(TestFeature.cols())
	{
		Context::check(
			gple_predict_complex(Context::get(), kernel.handle(), TestFeature.data(), TestFeature.cols(), TestLabel ? reinterpret_cast<const double*>(TestLabel->data()) : nullptr, reinterpret_cast<double*>(Prediction.data()), Variance.data(), reinterpret_cast<double*>(Cutoff.data()), TestLabel ? &Error : nullptr, nullptr),
			"PredictiveComplexKernel"
		);
	}
	const std::vector<double>& get_variance() const { return Variance; }
	const std::vector<std::complex<double>>& get_cutoff_prediction() const { return Cutoff; }
	double get_error() const { return Error; }

private:
	std::vector<double> Variance;
	std::vector<std::complex<double>> Cutoff, Prediction;
	double Error = std::numeric_limits<double>::quiet_NaN();
};

using AllTrainingSets = std::array<ElementTrainingSet, NumElements>; // gple/predict.h:14 (lower-triangular order rho00, rho10, rho11)

/// construct_training_sets (gple/predict.cpp:246-280): AoS points -> (2 x n features, n complex labels) per element
inline AllTrainingSets construct_training_sets(const AllPoints& density)
{
	AllTrainingSets result;
	for (std::size_t e = 0; e < NumElements; e++)
	{
		const std::size_t n = density[e].size();
		result[e] = ElementTrainingSet{PhasePoints(n), std::vector<std::complex<double>>(n)};
		for (std::size_t i = 0; i < n; i++)
		{
			std::get<0>(result[e])(0, i) = density[e][i].r[0];
			std::get<0>(result[e])(1, i) = density[e][i].r[1];
			std::get<1>(result[e])[i] = density[e][i].rho;
		}
	}
	return result;
}

/// One trained element model per (training set, parameters).  NLopt evaluates the objective and the constraint callbacks of the
/// constrained stages at the same parameters in turn (opt.cpp:594-617 / 644-719, 844-870 / 879-929), and each of them builds
/// the element models from scratch; sharing the model halves the factorisations and derivative products of those stages.  An
/// entry is keyed by the address of the element's training set (caller-owned and immutable during an optimisation,
/// opt.cpp:538-545) and holds the last parameters seen; a request is a hit when the parameters are identical and the entry was
/// trained with at least the requested quantities.  `clear()` at the start and end of an optimisation bounds its lifetime.
class ModelCache
{
public:
	static ModelCache& instance()
	{
		static ModelCache c;
		return c;
	}
	std::shared_ptr<const TrainingKernel> real(const ElementTrainingSet& ts, const ParameterVector& theta, const bool err, const bool avg, const bool deriv)
	{
		return get<TrainingKernel>(Real, ts, theta, err, avg, deriv);
	}
	std::shared_ptr<const TrainingComplexKernel> complex(const ElementTrainingSet& ts, const ParameterVector& theta, const bool err, const bool avg, const bool deriv)
	{
		return get<TrainingComplexKernel>(Complex, ts, theta, err, avg, deriv);
	}
	void clear()
	{
		const std::lock_guard<std::mutex> lock(Mutex);
		Real.clear();
		Complex.clear();
	}
	std::size_t hits() const { return Hits; }
	std::size_t misses() const { return Misses; }

private:
	/// An entry is keyed by the ADDRESS of the training set (the callbacks of one optimisation pass the same object again and
	/// again) plus a fingerprint of its CONTENT -- size and every feature / label value folded into 64 bits -- so that a
	/// different set that happens to live at a recycled address can never be served a stale model
	static std::uint64_t fingerprint(const ElementTrainingSet& ts)
	{
		std::uint64_t h = 1469598103934665603ull;
		// word-wise (every buffer here is a whole number of 8-byte values): the cache is consulted several times per loss evaluation
		auto fold = [&h](const void* p, const std::size_t bytes)
		{
			const unsigned char* c = static_cast<const unsigned char*>(p);
			for (std::size_t i = 0; i + 8 <= bytes; i += 8)
			{
				std::uint64_t w;
				std::memcpy(&w, c + i, 8);
				h = (h ^ w) * 1099511628211ull;
				h ^= h >> 29;
			}
		};
		const auto& [X, y] = ts;
		const std::size_t n = X.cols();
		fold(&n, sizeof(n));
		fold(X.data(), 2 * n * sizeof(double));
		fold(y.data(), y.size() * sizeof(std::complex<double>));
		return h;
	}
	template <typename K>
	struct Entry
	{
		const ElementTrainingSet* key;
		int group; // slot group of the thread that built it: concurrent restart stages keep one entry per training set EACH
		std::uint64_t print;
		ParameterVector theta;
		unsigned flags;
		std::shared_ptr<const K> model;
	};
	std::mutex Mutex;
	std::vector<Entry<TrainingKernel>> Real;
	std::vector<Entry<TrainingComplexKernel>> Complex;
	std::atomic<std::size_t> Hits{0}, Misses{0};

	template <typename K>
	std::shared_ptr<const K> get(std::vector<Entry<K>>& entries, const ElementTrainingSet& ts, const ParameterVector& theta, const bool err, const bool avg, const bool deriv)
	{
		const unsigned want = (err ? 1u : 0u) | (avg ? 2u : 0u) | (deriv ? 4u : 0u);
		const std::uint64_t print = fingerprint(ts);
		const int group = Context::group();
		{
			const std::lock_guard<std::mutex> lock(Mutex);
			for (const auto& e : entries)
			{
				if (e.key == &ts && e.group == group && e.print == print && e.theta == theta && (e.flags & want) == want)
				{
					Hits++;
					return e.model;
				}
			}
		}
		Misses++;
		// within a slot group the same element is never evaluated by two threads at once, so training happens outside the lock
		auto model = std::make_shared<const K>(theta, ts, err, avg, deriv);
		const std::lock_guard<std::mutex> lock(Mutex);
		for (auto& e : entries)
		{
			if (e.key == &ts && e.group == group)
			{
				e = Entry<K>{&ts, group, print, theta, want, model};
				return model;
			}
		}
		entries.push_back(Entry<K>{&ts, group, print, theta, want, model});
		return model;
	}
};

/// gple/predict.h:89-143 -- the (up to) three element models of one time step; the three factorisations run concurrently
class TrainingKernels
{
public:
	static constexpr std::size_t NumTotalParameters = 4 * NumPES + 8; // gple/predict.h:17
	using QuantumVectorD = std::array<double, NumPES>;

	/// gple/predict.cpp:362-388; an element stays nullopt when its training set is empty or all its parameters are 0 (:339-357)
	/// Distributed (multi-GPU, optimiser callbacks only): every rank trains the elements it owns (element_owner) and the
	/// averages and their derivatives are summed over the ranks, so the calculate_* / *_derivative getters return the same
	/// numbers everywhere while each factorisation happens once; handle() / Diagonal / OffDiagonal are local-only then.
	TrainingKernels(const std::array<ParameterVector, NumElements>& ParameterVectors, const AllTrainingSets& TrainingSets, bool IsToCalculateError, bool IsToCalculateAverage, bool IsToCalculateDerivative, ModelCache* Cache = nullptr, bool Distributed = false)
	{
		const int ranks = Distributed ? Context::num_ranks() : 1, me = ranks > 1 ? Context::rank() : 0;
		for_each_element_concurrently(
			[&](const std::size_t e)
			{
				const bool all_zero = std::all_of(ParameterVectors[e].cbegin(), ParameterVectors[e].cend(), [](double d) { return d == 0.0; });
				if (std::get<0>(TrainingSets[e]).cols() == 0 || all_zero || element_owner(e, ranks) != me)
				{
					return;
				}
				if (e == 1)
				{
					OffDiagonal = Cache != nullptr ? Cache->complex(TrainingSets[e], ParameterVectors[e], IsToCalculateError, IsToCalculateAverage, IsToCalculateDerivative)
												   : std::make_shared<const TrainingComplexKernel>(ParameterVectors[e], TrainingSets[e], IsToCalculateError, IsToCalculateAverage, IsToCalculateDerivative);
				}
				else
				{
					Diagonal[e / 2] = Cache != nullptr ? Cache->real(TrainingSets[e], ParameterVectors[e], IsToCalculateError, IsToCalculateAverage, IsToCalculateDerivative)
													   : std::make_shared<const TrainingKernel>(ParameterVectors[e], TrainingSets[e], IsToCalculateError, IsToCalculateAverage, IsToCalculateDerivative);
				}
			}
		);
		// averages of the elements: [population, <x>, <p>, purity, d population (4), d purity (8)] per element
		Averages.assign(NumElements * AvgStride, 0.0);
		for (std::size_t i = 0; i < NumPES; i++)
		{
			if (Diagonal[i] && IsToCalculateAverage)
			{
				double* a = &Averages[2 * i * AvgStride];
				a[0] = Diagonal[i]->get_population();
				const ClassicalPhaseVector r = Diagonal[i]->get_1st_order_average();
				a[1] = r[0];
				a[2] = r[1];
				a[3] = Diagonal[i]->get_purity();
				if (IsToCalculateDerivative)
				{
					const auto dp = Diagonal[i]->get_population_derivative(), du = Diagonal[i]->get_purity_derivative();
					std::copy(dp.cbegin(), dp.cend(), a + 4);
					std::copy(du.cbegin(), du.cend(), a + 8);
				}
			}
		}
		if (OffDiagonal && IsToCalculateAverage)
		{
			double* a = &Averages[AvgStride];
			a[3] = OffDiagonal->get_purity();
			if (IsToCalculateDerivative)
			{
				const auto du = OffDiagonal->get_purity_derivative();
				std::copy(du.cbegin(), du.cend(), a + 8);
			}
		}
		if (ranks > 1)
		{
			all_reduce_sum(Averages);
		}
	}
	/// gple/predict.cpp:390-393 (error = true, average = true, derivative = false)
	TrainingKernels(const std::array<ParameterVector, NumElements>& ParameterVectors, const AllPoints& density):
		TrainingKernels(ParameterVectors, construct_training_sets(density), true, true, false)
	{
	}
	double calculate_population() const // gple/predict.cpp:395-406
	{
		return avg(0)[0] + avg(2)[0];
	}
	ClassicalPhaseVector calculate_1st_order_average() const // gple/predict.cpp:408-419
	{
		return {avg(0)[1] + avg(2)[1], avg(0)[2] + avg(2)[2]};
	}
	double calculate_total_energy_average(const QuantumVectorD& Energies) const // gple/predict.cpp:423-436
	{
		double r = 0.0;
		for (std::size_t i = 0; i < NumPES; i++)
		{
			r += avg(2 * i)[0] * Energies[i];
		}
		return r;
	}
	double calculate_purity() const // gple/predict.cpp:439-463
	{
		double r = 2.0 * avg(1)[3];
		for (std::size_t i = 0; i < NumPES; i++)
		{
			r += avg(2 * i)[3];
		}
		return r;
	}
	/// gple/predict.cpp:465-484: diagonal parameters only (2 x 4)
	ParameterVector population_derivative() const
	{
		ParameterVector r(4 * NumPES, 0.0);
		for (std::size_t i = 0; i < NumPES; i++)
		{
			std::copy(avg(2 * i) + 4, avg(2 * i) + 8, r.begin() + 4 * i);
		}
		return r;
	}
	/// gple/predict.cpp:486-510
	ParameterVector total_energy_derivative(const QuantumVectorD& Energies) const
	{
		ParameterVector r = population_derivative();
		for (std::size_t i = 0; i < NumPES; i++)
		{
			for (std::size_t p = 0; p < 4; p++)
			{
				r[4 * i + p] *= Energies[i];
			}
		}
		return r;
	}
	/// gple/predict.cpp:512-559: all 16 parameters in element order, the off-diagonal element weighted by 2
	ParameterVector purity_derivative() const
	{
		ParameterVector r(NumTotalParameters, 0.0);
		std::copy(avg(0) + 8, avg(0) + 12, r.begin());
		std::transform(avg(1) + 8, avg(1) + 16, r.begin() + 4, [](double x) { return 2.0 * x; });
		std::copy(avg(2) + 8, avg(2) + 12, r.begin() + 12);
		return r;
	}
	const gple_model* handle(std::size_t element) const
	{
		if (element == 1)
		{
			return OffDiagonal ? OffDiagonal->handle() : nullptr;
		}
		return Diagonal[element / 2] ? Diagonal[element / 2]->handle() : nullptr;
	}
	/// empty = the reference's std::nullopt (element not populated); shared so that a ModelCache can serve the same model twice
	std::array<std::shared_ptr<const TrainingKernel>, NumPES> Diagonal;
	std::shared_ptr<const TrainingComplexKernel> OffDiagonal;

private:
	static constexpr std::size_t AvgStride = 16;
	std::vector<double> Averages; // NumElements x AvgStride, zeros for an absent element
	const double* avg(const std::size_t element) const { return &Averages[element * AvgStride]; }
};

/// gple/evolve.h:16-21 with the GPR-backed predict_distribution of gple/main.cpp:75-101
inline void evolve(AllPoints& density, double mass, double dt, const TrainingKernels& kernels, int pes_model)
{
	Context::check(
		gple_evolve_sharded(Context::main(), pes_model, kernels.handle(0), kernels.handle(1), kernels.handle(2), reinterpret_cast<double*>(density[0].data()), density[0].size(), reinterpret_cast<double*>(density[1].data()), density[1].size(), reinterpret_cast<double*>(density[2].data()), density[2].size(), mass, dt),
		"evolve"
	);
}

/// output_phase (gple/output.cpp:181-233): cutoff prediction (two lines: real and imaginary part) and variance (one line) of
/// every element on the output grid, elements in lower-triangular order, absent elements as zeros; values separated by single
/// blanks at the stream's precision like Eigen's VectorFormatter (gple/stdafx.h:129); a blank line closes each record.
inline void output_phase(std::ostream& phase, std::ostream& variance, const TrainingKernels& AllKernels, const PhasePoints& PhaseGrids)
{
	const std::size_t NumPoints = PhaseGrids.cols();
	auto line = [NumPoints](std::ostream& os, auto&& value)
	{
		for (std::size_t i = 0; i < NumPoints; i++)
		{
			os << (i == 0 ? "" : " ") << value(i);
		}
		os << '\n';
	};
	auto zero = [](std::size_t) { return 0.0; };
	for (std::size_t e = 0; e < NumElements; e++)
	{
		if (e != 1 && AllKernels.Diagonal[e / 2])
		{
			const PredictiveKernel k(PhaseGrids, *AllKernels.Diagonal[e / 2], false);
			line(phase, [&k](std::size_t i) { return k.get_cutoff_prediction()[i]; });
			line(phase, zero);
			line(variance, [&k](std::size_t i) { return k.get_variance()[i]; });
		}
		else if (e == 1 && AllKernels.OffDiagonal)
		{
			const PredictiveComplexKernel ck(PhaseGrids, *AllKernels.OffDiagonal, false);
			line(phase, [&ck](std::size_t i) { return ck.get_cutoff_prediction()[i].real(); });
			line(phase, [&ck](std::size_t i) { return ck.get_cutoff_prediction()[i].imag(); });
			line(variance, [&ck](std::size_t i) { return ck.get_variance()[i]; });
		}
		else
		{
			line(phase, zero);
			line(phase, zero);
			line(variance, zero);
		}
	}
	phase << '\n';
	variance << '\n';
}

/// gple/pes.h:47 (one position)
inline std::array<double, NumPES> adiabatic_potential(double x, int pes_model)
{
	double E[2], F[3], D[1];
	Context::check(gple_pes(Context::get(), pes_model, &x, 1, E, F, D), "adiabatic_potential");
	return {E[0], E[1]};
}

/// gple/predict.cpp:65-87
inline std::array<double, NumPES> calculate_population_each_surface(const AllPoints& density, double mass, int pes_model)
{
	std::array<double, NumPES> r{0.0, 0.0};
	for (std::size_t s = 0; s < NumPES; s++)
	{
		const ElementPoints& pts = density[2 * s];
		if (!pts.empty())
		{
			double o[9];
			Context::check(gple_observables(Context::get(), pes_model, reinterpret_cast<const double*>(pts.data()), pts.size(), mass, int(s), o), "observables");
			r[s] = o[0];
		}
	}
	const double sum = r[0] + r[1];
	return {r[0] / sum, r[1] / sum};
}

} // namespace gple_host
