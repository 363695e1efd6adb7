// TEST INFRASTRUCTURE ONLY -- oracle/refstub/refstub_blas.hpp
//
// Large dense products of the Eigen stand-in.  The reference builds with EIGEN_USE_MKL_ALL
// (gple/stdafx.h:9-11, makefile:4), i.e. its big GEMMs run on a tuned host BLAS.  MKL is not in this
// image; the same role is played by the OpenBLAS that scipy bundles (path passed in REFSTUB_BLAS by
// oracle/ref.py, symbols scipy_cblas_dgemm / scipy_cblas_zgemm), or, when that cannot be loaded, by a
// threaded column-sweep product.  Column-major operands, C = A(m x k) * B(k x n).
#pragma once
#include <complex>
#include <cstddef>
#include <cstdlib>
#include <dlfcn.h>

#include "../linalg.hpp"

namespace refstub
{
using dgemm_fn = void (*)(int, int, int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
using zgemm_fn = void (*)(int, int, int, int, int, int, const void*, const void*, int, const void*, int, const void*, void*, int);

struct BlasHandle
{
	dgemm_fn dgemm = nullptr;
	zgemm_fn zgemm = nullptr;
	BlasHandle()
	{
		const char* path = std::getenv("REFSTUB_BLAS");
		if (path == nullptr || *path == 0)
		{
			return;
		}
		void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
		if (h == nullptr)
		{
			return;
		}
		dgemm = reinterpret_cast<dgemm_fn>(dlsym(h, "scipy_cblas_dgemm"));
		zgemm = reinterpret_cast<zgemm_fn>(dlsym(h, "scipy_cblas_zgemm"));
		if (dgemm == nullptr)
		{
			dgemm = reinterpret_cast<dgemm_fn>(dlsym(h, "cblas_dgemm"));
			zgemm = reinterpret_cast<zgemm_fn>(dlsym(h, "cblas_zgemm"));
		}
	}
};
inline const BlasHandle& blas()
{
	static const BlasHandle h;
	return h;
}
inline bool have_blas()
{
	return blas().dgemm != nullptr && blas().zgemm != nullptr;
}

template <typename T>
inline void gemm_fallback(std::ptrdiff_t m, std::ptrdiff_t n, std::ptrdiff_t k, const T* a, const T* b, T* c)
{
	orc::parallel_for(
		static_cast<std::size_t>(n),
		[&](const std::size_t j)
		{
			T* cj = c + j * m;
			for (std::ptrdiff_t i = 0; i < m; i++)
			{
				cj[i] = T(0);
			}
			for (std::ptrdiff_t l = 0; l < k; l++)
			{
				const T x = b[l + static_cast<std::ptrdiff_t>(j) * k];
				const T* al = a + l * m;
				for (std::ptrdiff_t i = 0; i < m; i++)
				{
					cj[i] += al[i] * x;
				}
			}
		}
	);
}

inline void gemm(std::ptrdiff_t m, std::ptrdiff_t n, std::ptrdiff_t k, const double* a, const double* b, double* c)
{
	if (blas().dgemm != nullptr)
	{
		// CblasColMajor = 102, CblasNoTrans = 111
		blas().dgemm(102, 111, 111, static_cast<int>(m), static_cast<int>(n), static_cast<int>(k), 1.0, a, static_cast<int>(m), b, static_cast<int>(k), 0.0, c, static_cast<int>(m));
	}
	else
	{
		gemm_fallback(m, n, k, a, b, c);
	}
}
inline void gemm(std::ptrdiff_t m, std::ptrdiff_t n, std::ptrdiff_t k, const std::complex<double>* a, const std::complex<double>* b, std::complex<double>* c)
{
	if (blas().zgemm != nullptr)
	{
		const std::complex<double> one(1.0, 0.0), zero(0.0, 0.0);
		blas().zgemm(102, 111, 111, static_cast<int>(m), static_cast<int>(n), static_cast<int>(k), &one, a, static_cast<int>(m), b, static_cast<int>(k), &zero, c, static_cast<int>(m));
	}
	else
	{
		gemm_fallback(m, n, k, a, b, c);
	}
}
} // namespace refstub
