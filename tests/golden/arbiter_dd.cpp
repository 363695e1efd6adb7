// TEST INFRASTRUCTURE ONLY -- high-precision arbiter at BASELINE.json's own sizes.
//
// Evaluates the reference's FORMULAS for one density-matrix element in double-double arithmetic (106-bit significand,
// error-free transformations on FMA): covariance (gple/kernel.cpp:38-85, complex_kernel.cpp:134-164), K^-1 y', the
// LOOCV error (kernel.cpp:279-287, complex_kernel.cpp:262-286), prediction, variance and cutoff prediction
// (kernel.cpp:481-522, kernel.h:301-332, complex_kernel.cpp:594-646), from the SAME double-precision inputs every
// double-precision evaluation sees.  With cond(K) ~ 1e7..1e8 a double evaluation of these quantities carries
// eps * cond(K) ~ 1e-9..1e-8; the 1e-32 * cond(K) of this program is exact for the purpose of deciding which of the
// double evaluations (compiled reference, oracle restatement, CUDA path) is closer to the true value
// (tests/test_arbiter.py).  The mpmath arbiter (make_arbiter.py, N = 96) pins this program in turn.
//
// The complex element is evaluated as the covariance of [Re f; Im f] (order 2N), which is the same Gaussian process as
// the widely-linear form of complex_kernel.h:12-13 in exact arithmetic:
//     C = s^2 [[K_R + sn^2/2 d, K_C], [K_C, K_I + sn^2/2 d]],  M = C^-1,  w = M [Re y'; Im y'],
//     P_ii = (Mrr + Mii)_ii / 4,  Q_ii = (Mrr - Mii)_ii / 4 - i (Mri)_ii / 2,  v = (w_r + i w_i) / 2.
//
// usage: arbiter_dd in.bin out.bin      (driven by make_arbiter_dd.py)
//   in : int64 kind (0 real, 1 complex; +2: labels rescaled exactly instead of in double, as the mpmath arbiter does), N, Q; double theta[8]; X[N][2]; y[N][2] (re, im); Xq[Q][2]
//   out: double error; v[N][2]; pred[Q][2]; var[Q]; cutoff[Q][2]     (rounded to double)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct dd
{
	double hi, lo;
	dd(): hi(0.0), lo(0.0) {}
	dd(double h): hi(h), lo(0.0) {}
	dd(double h, double l): hi(h), lo(l) {}
};
static inline dd quick_two_sum(double a, double b)
{
	const double s = a + b;
	return dd(s, b - (s - a));
}
static inline dd two_sum(double a, double b)
{
	const double s = a + b, bb = s - a;
	return dd(s, (a - (s - bb)) + (b - bb));
}
static inline dd two_prod(double a, double b)
{
	const double p = a * b;
	return dd(p, std::fma(a, b, -p));
}
static inline dd operator+(const dd a, const dd b)
{
	dd s = two_sum(a.hi, b.hi);
	const dd t = two_sum(a.lo, b.lo);
	s.lo += t.hi;
	s = quick_two_sum(s.hi, s.lo);
	s.lo += t.lo;
	return quick_two_sum(s.hi, s.lo);
}
static inline dd operator-(const dd a) { return dd(-a.hi, -a.lo); }
static inline dd operator-(const dd a, const dd b) { return a + (-b); }
static inline dd operator*(const dd a, const dd b)
{
	dd p = two_prod(a.hi, b.hi);
	p.lo += a.hi * b.lo + a.lo * b.hi;
	return quick_two_sum(p.hi, p.lo);
}
static inline dd operator/(const dd a, const dd b)
{
	const double q1 = a.hi / b.hi;
	dd r = a - b * dd(q1);
	const double q2 = r.hi / b.hi;
	r = r - b * dd(q2);
	const double q3 = r.hi / b.hi;
	return quick_two_sum(q1, q2) + dd(q3);
}
static inline dd sqrt_dd(const dd a)
{
	if (a.hi <= 0.0)
	{
		return dd(a.hi == 0.0 ? 0.0 : std::nan(""));
	}
	const double x = 1.0 / std::sqrt(a.hi), ax = a.hi * x;
	const dd r = a - two_prod(ax, ax);
	return two_sum(ax, r.hi * (x * 0.5));
}
static inline dd ldexp_dd(const dd a, int e) { return dd(std::ldexp(a.hi, e), std::ldexp(a.lo, e)); }
static inline bool operator>=(const dd a, const dd b) { return a.hi > b.hi || (a.hi == b.hi && a.lo >= b.lo); }
static inline bool operator<=(const dd a, const dd b) { return b >= a; }
/// exp(x), x <= 0: x = k ln2 + r, Taylor series of expm1(r / 512), nine squarings of 1 + s in the form 2 s + s^2
static dd exp_dd(const dd x)
{
	if (x.hi < -740.0)
	{
		return dd(0.0);
	}
	static const dd ln2(0.6931471805599453094, 2.3190468138462995584e-17);
	const double k = std::nearbyint(x.hi / ln2.hi);
	const dd r = ldexp_dd(x - ln2 * dd(k), -9);
	dd term = r, s = r;
	for (int i = 2; i <= 14; i++)
	{
		term = term * r / dd(double(i));
		s = s + term;
	}
	for (int i = 0; i < 9; i++)
	{
		s = ldexp_dd(s, 1) + s * s;
	}
	return ldexp_dd(dd(1.0) + s, int(k));
}
static const dd PI_DD(3.141592653589793116, 1.2246467991473532072e-16);

struct Block
{
	dd mag2, lx, lp, diag;
};
static inline dd gauss(const Block& b, const double* a, const double* c)
{
	// coordinates are doubles: their difference is exact in dd
	const dd dx = dd(two_sum(a[0], -c[0])) / b.lx, dp = dd(two_sum(a[1], -c[1])) / b.lp;
	return b.mag2 * exp_dd(ldexp_dd(-(dx * dx + dp * dp), -1));
}
/// dot product of two contiguous dd rows with four independent accumulators
static inline dd dot(const dd* a, const dd* b, const long n)
{
	dd s0, s1, s2, s3;
	long k = 0;
	for (; k + 4 <= n; k += 4)
	{
		s0 = s0 + a[k] * b[k];
		s1 = s1 + a[k + 1] * b[k + 1];
		s2 = s2 + a[k + 2] * b[k + 2];
		s3 = s3 + a[k + 3] * b[k + 3];
	}
	for (; k < n; k++)
	{
		s0 = s0 + a[k] * b[k];
	}
	return (s0 + s1) + (s2 + s3);
}
static dd gate(const dd f2, const dd var)
{
	if (f2 >= dd(4.0) * var)
	{
		return dd(1.0);
	}
	if (f2 <= var)
	{
		return dd(0.0);
	}
	const dd a = sqrt_dd(f2) / sqrt_dd(var), am1 = a - dd(1.0);
	return (dd(5.0) - dd(2.0) * a) * am1 * am1;
}

int main(int argc, char** argv)
{
	if (argc < 3)
	{
		return 2;
	}
	FILE* in = std::fopen(argv[1], "rb");
	int64_t head[3];
	double th[8];
	if (in == nullptr || std::fread(head, 8, 3, in) != 3 || std::fread(th, 8, 8, in) != 8)
	{
		return 3;
	}
	const int kind = int(head[0] & 1);
	const bool exact_labels = (head[0] & 2) != 0; // pin mode: y' = y * (10 / max|y|) without the rounding to double
	const long N = head[1], Q = head[2];
	std::vector<double> X(2 * N), y(2 * N), Xq(2 * Q);
	if (std::fread(X.data(), 8, 2 * N, in) != size_t(2 * N) || std::fread(y.data(), 8, 2 * N, in) != size_t(2 * N) || std::fread(Xq.data(), 8, 2 * Q, in) != size_t(2 * Q))
	{
		return 3;
	}
	std::fclose(in);
	const int nb = kind ? 2 : 1;
	const long n = nb * N;
	Block b[2][2];
	dd prior;
	if (!kind)
	{
		const dd sf(th[0]), sn(th[3]);
		b[0][0] = Block{sf * sf, dd(th[1]), dd(th[2]), sf * sf * sn * sn};
		prior = sf * sf * (dd(1.0) + sn * sn);
	}
	else
	{
		// complex_kernel.cpp:142-157
		const dd sg(th[0]), sr(th[1]), lrx(th[2]), lrp(th[3]), si(th[4]), lix(th[5]), lip(th[6]), sn(th[7]);
		const dd ssx = lrx * lrx + lix * lix, ssp = lrp * lrp + lip * lip;
		const dd lcx = sqrt_dd(ldexp_dd(ssx, -1)), lcp = sqrt_dd(ldexp_dd(ssp, -1));
		const dd sc2 = sr * si * (dd(2.0) * lrx * lix / ssx) * (dd(2.0) * lrp * lip / ssp); // sigma_C^2
		const dd s2 = sg * sg, hn = ldexp_dd(s2 * sn * sn, -1);
		b[0][0] = Block{s2 * sr * sr, lrx, lrp, hn};
		b[1][1] = Block{s2 * si * si, lix, lip, hn};
		b[0][1] = b[1][0] = Block{s2 * sc2, lcx, lcp, dd(0.0)};
		prior = s2 * (sr * sr + si * si + sn * sn);
	}
	// rescale factor exactly as the reference computes it (kernel.cpp:279, in double), labels = y * rescale in double: the
	// rounded labels are what every double-precision evaluation solves for (their rounding, amplified by 1 / sigma_n^2, would
	// otherwise count as error of all three)
	double mx = 0.0;
	for (long i = 0; i < N; i++)
	{
		mx = std::fmax(mx, kind ? std::hypot(y[2 * i], y[2 * i + 1]) : std::fabs(y[2 * i]));
	}
	const double rescale = 10.0 / mx;
	std::vector<dd> lab(n);
	const dd exact_rescale = dd(10.0) / dd(mx);
	for (long i = 0; i < N; i++)
	{
		lab[i] = exact_labels ? dd(y[2 * i]) * exact_rescale : dd(y[2 * i] * rescale);
		if (kind)
		{
			lab[N + i] = exact_labels ? dd(y[2 * i + 1]) * exact_rescale : dd(y[2 * i + 1] * rescale);
		}
	}
	// lower triangle of the covariance, row-major
	std::vector<dd> L(size_t(n) * n);
#pragma omp parallel for schedule(dynamic, 8)
	for (long I = 0; I < n; I++)
	{
		const int rb = int(I / N);
		const long i = I - rb * N;
		for (long J = 0; J <= I; J++)
		{
			const int cb = int(J / N);
			const long j = J - cb * N;
			dd v = gauss(b[rb][cb], &X[2 * i], &X[2 * j]);
			if (i == j)
			{
				v = v + b[rb][cb].diag;
			}
			L[size_t(I) * n + J] = v;
		}
	}
	// Cholesky, left-looking by columns: row dot products are contiguous
	for (long j = 0; j < n; j++)
	{
		const dd d = sqrt_dd(L[size_t(j) * n + j] - dot(&L[size_t(j) * n], &L[size_t(j) * n], j));
		if (!(d.hi > 0.0))
		{
			std::fprintf(stderr, "not positive definite at %ld\n", j);
			return 4;
		}
		L[size_t(j) * n + j] = d;
#pragma omp parallel for schedule(static)
		for (long i = j + 1; i < n; i++)
		{
			L[size_t(i) * n + j] = (L[size_t(i) * n + j] - dot(&L[size_t(i) * n], &L[size_t(j) * n], j)) / d;
		}
	}
	// Wt[j][i] = (L^-1)[i][j], i >= j: forward substitution per column j, row-major in i
	std::vector<dd> Wt(size_t(n) * n);
#pragma omp parallel for schedule(dynamic, 4)
	for (long j = 0; j < n; j++)
	{
		dd* w = &Wt[size_t(j) * n];
		w[j] = dd(1.0) / L[size_t(j) * n + j];
		for (long i = j + 1; i < n; i++)
		{
			w[i] = -dot(&L[size_t(i) * n + j], w + j, i - j) / L[size_t(i) * n + i];
		}
	}
	// z = W y', w = W^T z; diag(M) and diag(M_ri)
	std::vector<dd> z(n), sol(n), mdiag(n), mcross(N);
#pragma omp parallel for schedule(dynamic, 16)
	for (long i = 0; i < n; i++)
	{
		dd s;
		for (long k = 0; k <= i; k++)
		{
			s = s + Wt[size_t(k) * n + i] * lab[k];
		}
		z[i] = s;
	}
#pragma omp parallel for schedule(dynamic, 16)
	for (long k = 0; k < n; k++)
	{
		sol[k] = dot(&Wt[size_t(k) * n + k], &z[k], n - k);
		mdiag[k] = dot(&Wt[size_t(k) * n + k], &Wt[size_t(k) * n + k], n - k);
		if (kind && k < N)
		{
			mcross[k] = dot(&Wt[size_t(k) * n + N + k], &Wt[size_t(N + k) * n + N + k], n - N - k);
		}
	}
	dd err;
	std::vector<double> out_v(2 * N, 0.0);
	for (long i = 0; i < N; i++)
	{
		if (!kind)
		{
			const dd r = sol[i] / mdiag[i];
			err = err + r * r;
			out_v[2 * i] = sol[i].hi;
		}
		else
		{
			const dd p = ldexp_dd(mdiag[i] + mdiag[N + i], -2), qr = ldexp_dd(mdiag[i] - mdiag[N + i], -2), qi = -ldexp_dd(mcross[i], -1);
			const dd vr = ldexp_dd(sol[i], -1), vi = ldexp_dd(sol[N + i], -1);
			const dd qvr = qr * vr - qi * vi, qvi = qr * vi + qi * vr;
			const dd nr = p * vr - qvr, ni = p * vi + qvi, den = p * p - (qr * qr + qi * qi);
			const dd dr = nr / den, di = ni / den;
			err = err + dr * dr + di * di;
			out_v[2 * i] = vr.hi;
			out_v[2 * i + 1] = vi.hi;
		}
	}
	// queries
	std::vector<double> out_pred(2 * Q, 0.0), out_var(Q), out_cut(2 * Q, 0.0);
#pragma omp parallel for schedule(dynamic, 1)
	for (long m = 0; m < Q; m++)
	{
		dd f[2], qsum;
		std::vector<dd> ks(n), u(n);
		for (int rb = 0; rb < nb; rb++)
		{
			for (long J = 0; J < n; J++)
			{
				const int cb = int(J / N);
				const long j = J - cb * N;
				dd v = gauss(b[rb][cb], &Xq[2 * m], &X[2 * j]);
				if (Xq[2 * m] == X[2 * j] && Xq[2 * m + 1] == X[2 * j + 1])
				{
					v = v + b[rb][cb].diag; // delta_kernel, kernel.cpp:8-31
				}
				ks[J] = v;
			}
			f[rb] = dot(ks.data(), sol.data(), n);
			// u = W k^T, accumulated column by column of W (rows of Wt)
			for (long i = 0; i < n; i++)
			{
				u[i] = dd(0.0);
			}
			for (long j = 0; j < n; j++)
			{
				const dd kj = ks[j];
				const dd* w = &Wt[size_t(j) * n];
				for (long i = j; i < n; i++)
				{
					u[i] = u[i] + w[i] * kj;
				}
			}
			qsum = qsum + dot(u.data(), u.data(), n);
		}
		const dd var = prior - qsum;
		const dd f2 = f[0] * f[0] + f[1] * f[1];
		const dd g = gate(f2, var);
		out_pred[2 * m] = f[0].hi;
		out_pred[2 * m + 1] = f[1].hi;
		out_var[m] = var.hi;
		const dd rs = exact_labels ? exact_rescale : dd(rescale);
		out_cut[2 * m] = (f[0] * g / rs).hi;
		out_cut[2 * m + 1] = (f[1] * g / rs).hi;
	}
	FILE* out = std::fopen(argv[2], "wb");
	const double e = err.hi;
	std::fwrite(&e, 8, 1, out);
	std::fwrite(out_v.data(), 8, 2 * N, out);
	std::fwrite(out_pred.data(), 8, 2 * Q, out);
	std::fwrite(out_var.data(), 8, Q, out);
	std::fwrite(out_cut.data(), 8, 2 * Q, out);
	std::fclose(out);
	return 0;
}
