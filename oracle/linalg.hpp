// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// Nothing under oracle/ may be imported, linked or executed by the product path.
//
// Minimal dependency-free dense linear algebra used by the oracle restatement:
// column-major matrices (same layout as Eigen::MatrixXd / MatrixXcd in the reference),
// a pivoted LDL^T that follows the published algorithm of Eigen 3.4's
// `Eigen::LDLT` (Eigen/src/Cholesky/LDLT.h: `ldlt_inplace<Lower>::unblocked` and
// `LDLT::_solve_impl`).  Eigen is NOT vendored in /root/reference and is absent from this
// image (third-party dependency, version unpinned by the reference: `#include <Eigen/Eigen>`
// at gaussian_process_liouville_equation/stdafx.h:50), so this is a restatement of the
// published algorithm, anchored on the reference call sites kernel.cpp:281-283 and
// complex_kernel.cpp:264-266.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdlib>
#include <functional>
#include <limits>
#include <thread>
#include <vector>

namespace orc
{
using cplx = std::complex<double>;

inline unsigned num_threads()
{
	static const unsigned n = []() -> unsigned
	{
		if (const char* e = std::getenv("ORACLE_THREADS"))
		{
			const int v = std::atoi(e);
			if (v > 0)
			{
				return static_cast<unsigned>(v);
			}
		}
		const unsigned hc = std::thread::hardware_concurrency();
		return hc == 0 ? 1u : hc;
	}();
	return n;
}

/// 0 on a thread that may open a parallel loop, 1 inside a worker (nested loops then run serially)
inline int& parallel_depth()
{
	static thread_local int depth = 0;
	return depth;
}

/// Stand-in for the reference's `std::for_each(std::execution::par_unseq, indices...)` loops.
template <typename F>
inline void parallel_for(const std::size_t n, F&& f, const std::size_t min_grain = 1)
{
	const unsigned nt = static_cast<unsigned>(std::min<std::size_t>(num_threads(), std::max<std::size_t>(1, n / std::max<std::size_t>(1, min_grain))));
	if (nt <= 1 || n == 0 || parallel_depth() > 0)
	{
		// serial, also for a loop nested inside a worker of an enclosing parallel loop
		for (std::size_t i = 0; i < n; i++)
		{
			f(i);
		}
		return;
	}
	std::vector<std::thread> pool;
	pool.reserve(nt);
	for (unsigned t = 0; t < nt; t++)
	{
		pool.emplace_back(
			[t, nt, n, &f]()
			{
				parallel_depth() = 1;
				const std::size_t lo = n * t / nt, hi = n * (t + 1) / nt;
				for (std::size_t i = lo; i < hi; i++)
				{
					f(i);
				}
			}
		);
	}
	for (auto& th : pool)
	{
		th.join();
	}
}

inline double conj_of(const double x)
{
	return x;
}
inline cplx conj_of(const cplx& x)
{
	return std::conj(x);
}
inline double real_of(const double x)
{
	return x;
}
inline double real_of(const cplx& x)
{
	return x.real();
}

template <typename T>
struct Matrix
{
	std::size_t rows = 0, cols = 0;
	std::vector<T> d;
	Matrix() = default;
	Matrix(std::size_t r, std::size_t c, T fill = T(0)): rows(r), cols(c), d(r * c, fill) {}
	T& operator()(std::size_t i, std::size_t j) { return d[j * rows + i]; }
	const T& operator()(std::size_t i, std::size_t j) const { return d[j * rows + i]; }
	T* col(std::size_t j) { return d.data() + j * rows; }
	const T* col(std::size_t j) const { return d.data() + j * rows; }
	static Matrix identity(std::size_t n)
	{
		Matrix m(n, n);
		for (std::size_t i = 0; i < n; i++)
		{
			m(i, i) = T(1);
		}
		return m;
	}
};
using Mat = Matrix<double>;
using CMat = Matrix<cplx>;
using Vec = std::vector<double>;
using CVec = std::vector<cplx>;

/// C = op(A) * B with optional scalar; plain column-saxpy GEMM threaded over columns of C.
/// opA: 'N' none, 'C' conjugate (no transpose), 'H' adjoint
template <typename TA, typename TB>
auto matmul(const Matrix<TA>& A, const Matrix<TB>& B, const char opA = 'N') -> Matrix<decltype(TA() * TB())>
{
	using TC = decltype(TA() * TB());
	const std::size_t M = (opA == 'H') ? A.cols : A.rows, K = (opA == 'H') ? A.rows : A.cols, N = B.cols;
	assert(K == B.rows);
	Matrix<TC> C(M, N);
	parallel_for(
		N,
		[&](const std::size_t j)
		{
			TC* cj = C.col(j);
			if (opA == 'H')
			{
				for (std::size_t i = 0; i < M; i++)
				{
					TC s(0);
					const TA* ai = A.col(i);
					const TB* bj = B.col(j);
					for (std::size_t k = 0; k < K; k++)
					{
						s += conj_of(ai[k]) * bj[k];
					}
					cj[i] = s;
				}
			}
			else
			{
				for (std::size_t k = 0; k < K; k++)
				{
					const TB b = B(k, j);
					const TA* ak = A.col(k);
					if (opA == 'C')
					{
						for (std::size_t i = 0; i < M; i++)
						{
							cj[i] += conj_of(ak[i]) * b;
						}
					}
					else
					{
						for (std::size_t i = 0; i < M; i++)
						{
							cj[i] += ak[i] * b;
						}
					}
				}
			}
		}
	);
	return C;
}

template <typename TA, typename TB>
auto matvec(const Matrix<TA>& A, const std::vector<TB>& x) -> std::vector<decltype(TA() * TB())>
{
	using TC = decltype(TA() * TB());
	assert(A.cols == x.size());
	std::vector<TC> y(A.rows, TC(0));
	const unsigned nt = num_threads();
	if (A.rows * A.cols < (1u << 16) || nt <= 1)
	{
		for (std::size_t k = 0; k < A.cols; k++)
		{
			const TA* ak = A.col(k);
			for (std::size_t i = 0; i < A.rows; i++)
			{
				y[i] += ak[i] * x[k];
			}
		}
		return y;
	}
	// row-chunked so that every y[i] is still accumulated in k order (same sum order as serial)
	const std::size_t chunks = nt;
	parallel_for(
		chunks,
		[&](const std::size_t c)
		{
			const std::size_t lo = A.rows * c / chunks, hi = A.rows * (c + 1) / chunks;
			for (std::size_t k = 0; k < A.cols; k++)
			{
				const TA* ak = A.col(k);
				for (std::size_t i = lo; i < hi; i++)
				{
					y[i] += ak[i] * x[k];
				}
			}
		}
	);
	return y;
}

template <typename T>
Matrix<T> conjugate(const Matrix<T>& A)
{
	Matrix<T> R = A;
	for (auto& x : R.d)
	{
		x = conj_of(x);
	}
	return R;
}

template <typename T>
Matrix<T> adjoint(const Matrix<T>& A)
{
	Matrix<T> R(A.cols, A.rows);
	for (std::size_t j = 0; j < A.cols; j++)
	{
		for (std::size_t i = 0; i < A.rows; i++)
		{
			R(j, i) = conj_of(A(i, j));
		}
	}
	return R;
}

/// Dense copy of `A.selfadjointView<Eigen::Lower>()`
template <typename T>
Matrix<T> selfadjoint_lower(const Matrix<T>& A)
{
	Matrix<T> R = A;
	for (std::size_t j = 0; j < A.cols; j++)
	{
		for (std::size_t i = 0; i < j; i++)
		{
			R(i, j) = conj_of(A(j, i));
		}
	}
	return R;
}

/// Pivoted LDL^T of a self-adjoint matrix (lower part referenced), after Eigen::LDLT.
template <typename T>
class LDLT
{
public:
	explicit LDLT(const Matrix<T>& A): m(A), n(A.rows), tr(A.rows)
	{
		assert(A.rows == A.cols);
		factorize();
	}

	/// X = A^{-1} B, after Eigen::LDLT::_solve_impl; threaded over right-hand sides
	Matrix<T> solve(const Matrix<T>& B) const
	{
		assert(B.rows == n);
		Matrix<T> X = B;
		parallel_for(X.cols, [&](const std::size_t j) { solve_column(X.col(j)); });
		return X;
	}
	std::vector<T> solve(const std::vector<T>& b) const
	{
		std::vector<T> x = b;
		solve_column(x.data());
		return x;
	}
	const Matrix<T>& packed() const { return m; }

private:
	Matrix<T> m;
	std::size_t n;
	std::vector<std::size_t> tr;

	void factorize()
	{
		std::vector<T> temp(n);
		for (std::size_t k = 0; k < n; k++)
		{
			// largest |diagonal| in the remaining corner
			std::size_t piv = k;
			double big = std::abs(m(k, k));
			for (std::size_t i = k + 1; i < n; i++)
			{
				const double a = std::abs(m(i, i));
				if (a > big)
				{
					big = a;
					piv = i;
				}
			}
			tr[k] = piv;
			if (piv != k)
			{
				for (std::size_t c = 0; c < k; c++)
				{
					std::swap(m(k, c), m(piv, c));
				}
				for (std::size_t r = piv + 1; r < n; r++)
				{
					std::swap(m(r, k), m(r, piv));
				}
				std::swap(m(k, k), m(piv, piv));
				for (std::size_t i = k + 1; i < piv; i++)
				{
					const T tmp = m(i, k);
					m(i, k) = conj_of(m(piv, i));
					m(piv, i) = conj_of(tmp);
				}
				m(piv, k) = conj_of(m(piv, k));
			}
			const std::size_t rs = n - k - 1;
			if (k > 0)
			{
				for (std::size_t c = 0; c < k; c++)
				{
					temp[c] = real_of(m(c, c)) * conj_of(m(k, c));
				}
				T s(0);
				for (std::size_t c = 0; c < k; c++)
				{
					s += m(k, c) * temp[c];
				}
				m(k, k) -= s;
				if (rs > 0)
				{
					// A21 -= A20 * temp ; threaded over row chunks, per-row sums keep column order
					const std::size_t work = rs * k;
					const std::size_t chunks = work > (1u << 15) ? num_threads() : 1;
					parallel_for(
						chunks,
						[&](const std::size_t ch)
						{
							const std::size_t lo = k + 1 + rs * ch / chunks, hi = k + 1 + rs * (ch + 1) / chunks;
							T* a21 = m.col(k);
							for (std::size_t c = 0; c < k; c++)
							{
								const T t = temp[c];
								const T* a20 = m.col(c);
								for (std::size_t r = lo; r < hi; r++)
								{
									a21[r] -= a20[r] * t;
								}
							}
						}
					);
				}
			}
			const double akk = real_of(m(k, k));
			if (rs > 0 && std::abs(akk) > 0.0)
			{
				T* a21 = m.col(k);
				for (std::size_t r = k + 1; r < n; r++)
				{
					a21[r] /= akk;
				}
			}
		}
	}

	void solve_column(T* x) const
	{
		for (std::size_t k = 0; k < n; k++)
		{
			if (tr[k] != k)
			{
				std::swap(x[k], x[tr[k]]);
			}
		}
		// unit lower forward substitution
		for (std::size_t k = 0; k < n; k++)
		{
			const T xk = x[k];
			const T* lk = m.col(k);
			for (std::size_t r = k + 1; r < n; r++)
			{
				x[r] -= lk[r] * xk;
			}
		}
		const double tol = std::numeric_limits<double>::min();
		for (std::size_t k = 0; k < n; k++)
		{
			const double dk = real_of(m(k, k));
			if (std::abs(dk) > tol)
			{
				x[k] /= dk;
			}
			else
			{
				x[k] = T(0);
			}
		}
		// L^H backward substitution
		for (std::size_t kk = n; kk-- > 0;)
		{
			T s = x[kk];
			const T* lk = m.col(kk);
			for (std::size_t r = kk + 1; r < n; r++)
			{
				s -= conj_of(lk[r]) * x[r];
			}
			x[kk] = s;
		}
		for (std::size_t kk = n; kk-- > 0;)
		{
			if (tr[kk] != kk)
			{
				std::swap(x[kk], x[tr[kk]]);
			}
		}
	}
};

} // namespace orc
