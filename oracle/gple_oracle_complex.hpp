// TEST INFRASTRUCTURE ONLY -- CPU oracle, complex (widely-linear) GPR part.
// Restates gple/complex_kernel.cpp and gple/complex_kernel.h of the reference.  PARITY UNPINNED
// (see gple_oracle.hpp header).
#pragma once
#include "gple_oracle.hpp"

namespace orc
{
using namespace std::literals::complex_literals;

/// gple/complex_kernel.h:28: (magnitude, [(magnitude, char_length)] x {R, I}, noise)
struct CKParam
{
	double mag = 1.0;
	std::array<double, 2> sub_mag{1.0, 1.0};
	std::array<std::array<double, PhaseDim>, 2> sub_l{};
	double noise = 0.0;
};

/// parameter order: gple/complex_kernel.cpp:230-256
inline CKParam unpack_complex(const double* th)
{
	CKParam c;
	std::size_t i = 0;
	c.mag = th[i++];
	for (std::size_t k = 0; k < 2; k++)
	{
		c.sub_mag[k] = th[i++];
		for (std::size_t d = 0; d < PhaseDim; d++)
		{
			c.sub_l[k][d] = th[i++];
		}
	}
	c.noise = th[i++];
	return c;
}

inline Mat lincomb(const double a, const Mat& A, const double b, const Mat& B)
{
	Mat R(A.rows, A.cols);
	for (std::size_t i = 0; i < R.d.size(); i++)
	{
		R.d[i] = a * A.d[i] + b * B.d[i];
	}
	return R;
}

inline CMat to_complex(const Mat& A)
{
	CMat R(A.rows, A.cols);
	for (std::size_t i = 0; i < R.d.size(); i++)
	{
		R.d[i] = A.d[i];
	}
	return R;
}

/// gple/complex_kernel.cpp:134-200
struct ComplexKernelBase
{
	CKParam prm;
	KParam real_p, imag_p, corr_p;
	Points L, R;
	bool same;
	std::unique_ptr<KernelBase> KR, KI, KC;
	Mat K;
	CMat Kt;
	std::optional<std::array<Mat, NumComplexParams>> dK;
	std::optional<std::array<CMat, NumComplexParams>> dKt;

	ComplexKernelBase(const CKParam& P, const Points& left, const Points& right, const bool same_buffer, const bool deriv):
		prm(P), L(left), R(right), same(same_buffer)
	{
		// complex_kernel.cpp:142-157
		real_p = KParam{prm.sub_mag[0], prm.sub_l[0], 0.0};
		imag_p = KParam{prm.sub_mag[1], prm.sub_l[1], 0.0};
		{
			double prod = 1.0;
			for (std::size_t d = 0; d < PhaseDim; d++)
			{
				const double ss = sq(real_p.l[d]) + sq(imag_p.l[d]);
				prod *= 2.0 * real_p.l[d] * imag_p.l[d] / ss;
				corr_p.l[d] = std::sqrt(ss / 2.0);
			}
			corr_p.mag = std::sqrt(real_p.mag * imag_p.mag * prod);
			corr_p.noise = 0.0;
		}
		// complex_kernel.cpp:160-162: sub-kernels take the member copies -> never "same buffer"
		KR = std::make_unique<KernelBase>(real_p, L, R, false, deriv);
		KI = std::make_unique<KernelBase>(imag_p, L, R, false, deriv);
		KC = std::make_unique<KernelBase>(corr_p, L, R, false, deriv);
		// complex_kernel.cpp:163-164
		const Mat delta = delta_kernel(L, R, same);
		const double m2 = sq(prm.mag), n2 = sq(prm.noise);
		K = Mat(L.n, R.n);
		Kt = CMat(L.n, R.n);
		for (std::size_t i = 0; i < K.d.size(); i++)
		{
			K.d[i] = m2 * (KR->K.d[i] + KI->K.d[i] + n2 * delta.d[i]);
			Kt.d[i] = m2 * (KR->K.d[i] - KI->K.d[i] + 2.0i * KC->K.d[i]);
		}
		if (deriv)
		{
			const auto& dR = *KR->dK;
			const auto& dI = *KI->dK;
			const auto& dC = *KC->dK;
			// complex_kernel.cpp:20-59 (quirk q2: no sigma^2 on sub-kernel / noise derivatives)
			std::array<Mat, NumComplexParams> D;
			D[0] = K;
			for (auto& x : D[0].d)
			{
				x *= 2.0 / prm.mag;
			}
			for (std::size_t p = 0; p < 3; p++)
			{
				D[1 + p] = dR[p];
				D[4 + p] = dI[p];
			}
			D[7] = Mat(L.n, R.n);
			if (same)
			{
				for (std::size_t i = 0; i < L.n; i++)
				{
					D[7](i, i) = 2.0 * prm.noise;
				}
			}
			// complex_kernel.cpp:74-132
			std::array<CMat, NumComplexParams> Dt;
			Dt[0] = Kt;
			for (auto& x : Dt[0].d)
			{
				x *= 2.0 / prm.mag;
			}
			const Mat& C = KC->K;
			auto fill = [&](const std::size_t off, const KParam& sub, const std::array<Mat, NumRealParams>& dsub, const double sign)
			{
				Dt[off] = CMat(L.n, R.n);
				for (std::size_t i = 0; i < C.d.size(); i++)
				{
					Dt[off].d[i] = sign * dsub[0].d[i] + 2.0i / sub.mag * C.d[i];
				}
				for (std::size_t d = 0; d < PhaseDim; d++)
				{
					Dt[off + 1 + d] = CMat(L.n, R.n);
					const cplx c1 = 2.0i * (1.0 / sub.l[d] - sub.l[d] / sq(corr_p.l[d]));
					const cplx c2 = 1.0i * sub.l[d] / corr_p.l[d];
					for (std::size_t i = 0; i < C.d.size(); i++)
					{
						Dt[off + 1 + d].d[i] = sign * dsub[1 + d].d[i] + c1 * C.d[i] + c2 * dC[1 + d].d[i];
					}
				}
			};
			fill(1, real_p, dR, 1.0);
			fill(4, imag_p, dI, -1.0);
			Dt[7] = CMat(L.n, R.n);
			dK = std::move(D);
			dKt = std::move(Dt);
		}
	}
};

/// gple/complex_kernel.cpp:206-219
inline KParam construct_purity_auxiliary_mixed_kernel_params(const KParam& a, const KParam& b)
{
	KParam r;
	double prod = 1.0;
	for (std::size_t d = 0; d < PhaseDim; d++)
	{
		prod *= 0.5 * (sq(1.0 / a.l[d]) + sq(1.0 / b.l[d]));
		r.l[d] = std::sqrt(sq(a.l[d]) + sq(b.l[d]));
	}
	r.mag = a.mag * b.mag / std::sqrt(std::sqrt(prod));
	r.noise = 0.0;
	return r;
}

template <typename TA>
inline cplx quad_adjoint(const CVec& v, const Matrix<TA>& A, const CVec& w)
{
	// v^H A w
	const auto Aw = matvec(A, w);
	cplx s = 0;
	for (std::size_t i = 0; i < v.size(); i++)
	{
		s += std::conj(v[i]) * Aw[i];
	}
	return s;
}
template <typename TA>
inline cplx quad_transpose(const CVec& v, const Matrix<TA>& A, const CVec& w)
{
	// v^T A w
	const auto Aw = matvec(A, w);
	cplx s = 0;
	for (std::size_t i = 0; i < v.size(); i++)
	{
		s += v[i] * Aw[i];
	}
	return s;
}

/// gple/complex_kernel.cpp:221-592
struct TrainingComplexKernel
{
	std::vector<double> coords;
	Points X;
	CKParam prm;
	std::array<double, NumComplexParams> theta;
	std::unique_ptr<ComplexKernelBase> base;
	double rescale = 1.0;
	CVec label;
	CMat A, P, Q; // K^-1 conj(Kt); upper-left and lower-left blocks of the augmented inverse
	CVec v;
	std::optional<double> error, purity;
	std::optional<std::array<CVec, NumComplexParams>> dv;
	std::optional<std::array<double, NumComplexParams>> derror, dpurity;

	TrainingComplexKernel(const double* th, const double* feat, const cplx* y, const std::size_t N, const bool is_err, const bool is_avg, const bool is_deriv):
		coords(feat, feat + 2 * N)
	{
		X = Points{coords.data(), N};
		for (std::size_t i = 0; i < NumComplexParams; i++)
		{
			theta[i] = th[i];
		}
		prm = unpack_complex(th);
		base = std::make_unique<ComplexKernelBase>(prm, X, X, true, is_deriv);
		// complex_kernel.cpp:262-263
		double mx = 0.0;
		for (std::size_t i = 0; i < N; i++)
		{
			mx = std::max(mx, std::abs(y[i]));
		}
		rescale = RescaleMaximum / mx;
		label.resize(N);
		for (std::size_t i = 0; i < N; i++)
		{
			label[i] = y[i] * rescale;
		}
		// complex_kernel.cpp:264-268 (LDLT<MatrixXcd> of the real K promoted to complex)
		const CMat Kc = to_complex(base->K);
		const LDLT<cplx> dec(Kc);
		A = dec.solve(conjugate(base->Kt));
		{
			CMat S = matmul(base->Kt, A);
			for (std::size_t i = 0; i < S.d.size(); i++)
			{
				S.d[i] = Kc.d[i] - S.d[i];
			}
			const LDLT<cplx> dec2(selfadjoint_lower(S));
			P = selfadjoint_lower(dec2.solve(CMat::identity(N)));
		}
		Q = matmul(A, P);
		for (auto& x : Q.d)
		{
			x = -x;
		}
		{
			const CVec Py = matvec(P, label), Qy = matvec(Q, label);
			v.resize(N);
			for (std::size_t i = 0; i < N; i++)
			{
				v[i] = Py[i] + std::conj(Qy[i]);
			}
		}
		// complex_kernel.cpp:270-286
		if (is_err)
		{
			double e = 0;
			for (std::size_t i = 0; i < N; i++)
			{
				const cplx pd = P(i, i), qd = Q(i, i);
				const cplx diff = (pd * v[i] - std::conj(qd * v[i])) / (sq(pd.real()) - std::norm(qd));
				e += std::norm(diff);
			}
			error = e;
		}
		// complex_kernel.cpp:287-377
		std::unique_ptr<KernelBase> KRp, KIp, KCp, KRC, KIC;
		KParam rc_p, ic_p;
		if (is_avg || is_deriv)
		{
			rc_p = construct_purity_auxiliary_mixed_kernel_params(base->real_p, base->corr_p);
			ic_p = construct_purity_auxiliary_mixed_kernel_params(base->imag_p, base->corr_p);
		}
		Mat K1;
		CMat K2;
		if (is_avg)
		{
			KRp = std::make_unique<KernelBase>(construct_purity_auxiliary_kernel_params(base->real_p), X, X, true, is_deriv);
			KIp = std::make_unique<KernelBase>(construct_purity_auxiliary_kernel_params(base->imag_p), X, X, true, is_deriv);
			KCp = std::make_unique<KernelBase>(construct_purity_auxiliary_kernel_params(base->corr_p), X, X, true, is_deriv);
			KRC = std::make_unique<KernelBase>(rc_p, X, X, true, is_deriv);
			KIC = std::make_unique<KernelBase>(ic_p, X, X, true, is_deriv);
			K1 = Mat(N, N);
			K2 = CMat(N, N);
			for (std::size_t i = 0; i < K1.d.size(); i++)
			{
				K1.d[i] = KRp->K.d[i] + KIp->K.d[i] + 2.0 * KCp->K.d[i];
				K2.d[i] = KRp->K.d[i] - KIp->K.d[i] - 2.0i * (KRC->K.d[i] + KIC->K.d[i]);
			}
			const double gf = PurityFactor * 2.0 * std::numbers::pi;
			const double ttf = gf * sq(sq(prm.mag));
			purity = ttf * (quad_adjoint(v, K1, v).real() + quad_transpose(v, K2, v).real()) / sq(rescale);
		}
		if (is_deriv)
		{
			const auto& D = *base->dK;
			const auto& Dt = *base->dKt;
			const CMat Qh = adjoint(Q);
			std::array<CMat, NumComplexParams> dP, dQ;
			std::array<CVec, NumComplexParams> DV;
			for (std::size_t p = 0; p < NumComplexParams; p++)
			{
				// complex_kernel.cpp:388-396
				const CMat Dc = to_complex(D[p]);
				const CMat DP = matmul(Dc, P), DQ = matmul(Dc, Q);
				const CMat DtQ = matmul(Dt[p], Q), DtcP = matmul(Dt[p], P, 'C');
				CMat r = matmul(P, DP);
				const CMat t2 = matmul(Qh, DQ), t3 = matmul(P, DtQ), t4 = matmul(Qh, DtcP);
				for (std::size_t i = 0; i < r.d.size(); i++)
				{
					r.d[i] += t2.d[i] + t3.d[i] + t4.d[i];
				}
				const CMat rh = adjoint(r);
				for (std::size_t i = 0; i < r.d.size(); i++)
				{
					r.d[i] = -(r.d[i] + rh.d[i]) / 2.0;
				}
				dP[p] = std::move(r);
				// complex_kernel.cpp:414-420
				CMat rhs = DQ;
				for (std::size_t i = 0; i < rhs.d.size(); i++)
				{
					rhs.d[i] += DtcP.d[i];
				}
				CMat q = dec.solve(rhs);
				const CMat AdP = matmul(A, dP[p]);
				for (std::size_t i = 0; i < q.d.size(); i++)
				{
					q.d[i] = -q.d[i] - AdP.d[i];
				}
				dQ[p] = std::move(q);
				// complex_kernel.cpp:426-442
				const CVec a = matvec(dP[p], label), b = matvec(dQ[p], label);
				DV[p].resize(N);
				for (std::size_t i = 0; i < N; i++)
				{
					DV[p][i] = a[i] + std::conj(b[i]);
				}
			}
			// complex_kernel.cpp:444-474
			if (is_err)
			{
				std::array<double, NumComplexParams> de{};
				for (std::size_t p = 0; p < NumComplexParams; p++)
				{
					double s = 0;
					for (std::size_t i = 0; i < N; i++)
					{
						const cplx pd = P(i, i), qd = Q(i, i);
						const double sqd = sq(pd.real()) - std::norm(qd);
						const cplx diff = (pd * v[i] - std::conj(qd * v[i])) / sqd;
						const cplx pdd = dP[p](i, i), qdd = dQ[p](i, i), vd = DV[p][i];
						const cplx num = std::conj(diff) * (pdd * v[i] + pd * vd - std::conj(qdd * v[i] + qd * vd));
						const cplx den = -2.0 * std::norm(diff) * (pd * pdd - (std::conj(qd) * qdd).real());
						s += ((num + den) / sqd).real();
					}
					de[p] = 2.0 * s;
				}
				derror = de;
			}
			// complex_kernel.cpp:475-590
			if (is_avg)
			{
				const double gf = PurityFactor * 2.0 * std::numbers::pi;
				const KParam &rp = base->real_p, &ip = base->imag_p, &cp = base->corr_p;
				std::array<double, PhaseDim> ROverC, IOverC, ROverC2, IOverC2, ROverRC, ROverIC, IOverRC, IOverIC;
				for (std::size_t d = 0; d < PhaseDim; d++)
				{
					ROverC[d] = rp.l[d] / cp.l[d];
					IOverC[d] = ip.l[d] / cp.l[d];
					ROverC2[d] = rp.l[d] / sq(cp.l[d]);
					IOverC2[d] = ip.l[d] / sq(cp.l[d]);
					ROverRC[d] = rp.l[d] / rc_p.l[d];
					ROverIC[d] = rp.l[d] / ic_p.l[d];
					IOverRC[d] = ip.l[d] / rc_p.l[d];
					IOverIC[d] = ip.l[d] / ic_p.l[d];
				}
				const Mat Z(N, N);
				std::array<Mat, NumComplexParams> dRp, dIp, dCp, dRC, dIC;
				auto scaled = [](const double a, const Mat& M)
				{
					Mat R = M;
					for (auto& x : R.d)
					{
						x *= a;
					}
					return R;
				};
				dRp[0] = dIp[0] = dCp[0] = dRC[0] = dIC[0] = Z;
				// real kernel parameters (complex_kernel.cpp:523-544)
				dRp[1] = scaled(4.0 / rp.mag, KRp->K);
				dIp[1] = Z;
				dCp[1] = scaled(2.0 / rp.mag, KCp->K);
				dRC[1] = scaled(3.0 / rp.mag, KRC->K);
				dIC[1] = scaled(1.0 / rp.mag, KIC->K);
				for (std::size_t d = 0; d < PhaseDim; d++)
				{
					// quirk q10: the reference indexes the sub-kernel derivative array with the COMPLEX parameter
					// position (complex_kernel.cpp:534-541: `KRprimeDerivatives[iDim + iParam]`, iParam == 2), i.e. entry
					// d+2 of [mag, l_x, l_p, noise] -- l_p for d == 0 and the (zero) noise derivative for d == 1.
					const std::size_t ip_ = 2 + d, sub = 2 + d;
					dRp[ip_] = lincomb(1.0 / rp.l[d], KRp->K, std::numbers::sqrt2, (*KRp->dK)[sub]);
					dIp[ip_] = Z;
					dCp[ip_] = lincomb(2.0 / rp.l[d] - 3.0 * ROverC2[d] / 2.0, KCp->K, 1.0 / std::numbers::sqrt2 * ROverC[d], (*KCp->dK)[sub]);
					dRC[ip_] = lincomb(2.0 / rp.l[d] - ROverC2[d] / 2.0, KRC->K, 1.5 * ROverRC[d], lincomb(1.0, (*KRC->dK)[sub], -1.0 / rc_p.l[d], KRC->K));
					dIC[ip_] = lincomb(1.0 / rp.l[d] - ROverC2[d] / 2.0, KIC->K, ROverIC[d] / 2.0, lincomb(1.0, (*KIC->dK)[sub], -1.0 / ic_p.l[d], KIC->K));
				}
				// imaginary kernel parameters (complex_kernel.cpp:546-567)
				dRp[4] = Z;
				dIp[4] = scaled(4.0 / ip.mag, KIp->K);
				dCp[4] = scaled(2.0 / ip.mag, KCp->K);
				dRC[4] = scaled(1.0 / ip.mag, KRC->K);
				dIC[4] = scaled(3.0 / ip.mag, KIC->K);
				for (std::size_t d = 0; d < PhaseDim; d++)
				{
					// quirk q10 again (complex_kernel.cpp:558-564: index iDim + iParam - 3 with iParam == 5)
					const std::size_t ip_ = 5 + d, sub = 2 + d;
					dRp[ip_] = Z;
					dIp[ip_] = lincomb(1.0 / ip.l[d], KIp->K, std::numbers::sqrt2, (*KIp->dK)[sub]);
					dCp[ip_] = lincomb(2.0 / ip.l[d] - 3.0 * IOverC2[d] / 2.0, KCp->K, 1.0 / std::numbers::sqrt2 * IOverC[d], (*KCp->dK)[sub]);
					dRC[ip_] = lincomb(1.0 / ip.l[d] - IOverC2[d] / 2.0, KRC->K, IOverRC[d] / 2.0, lincomb(1.0, (*KRC->dK)[sub], -1.0 / rc_p.l[d], KRC->K));
					dIC[ip_] = lincomb(2.0 / ip.l[d] - IOverC2[d] / 2.0, KIC->K, 1.5 * IOverIC[d], lincomb(1.0, (*KIC->dK)[sub], -1.0 / ic_p.l[d], KIC->K));
				}
				dRp[7] = dIp[7] = dCp[7] = dRC[7] = dIC[7] = Z;
				std::array<double, NumComplexParams> du{};
				for (std::size_t p = 0; p < NumComplexParams; p++)
				{
					Mat K1d(N, N);
					CMat K2d(N, N);
					for (std::size_t i = 0; i < K1d.d.size(); i++)
					{
						K1d.d[i] = dRp[p].d[i] + dIp[p].d[i] + 2.0 * dCp[p].d[i];
						K2d.d[i] = dRp[p].d[i] - dIp[p].d[i] - 2.0i * (dRC[p].d[i] + dIC[p].d[i]);
					}
					double r = 2.0 * quad_adjoint(v, K1, DV[p]).real() + quad_adjoint(v, K1d, v).real()
						+ 2.0 * quad_transpose(v, K2, DV[p]).real() + quad_transpose(v, K2d, v).real();
					r *= gf / sq(rescale);
					du[p] = r;
				}
				dpurity = du;
			}
			dv = std::move(DV);
		}
	}

	/// gple/complex_kernel.h:192-204
	double get_magnitude() const
	{
		cplx s = 0;
		for (std::size_t i = 0; i < label.size(); i++)
		{
			s += std::conj(label[i]) * v[i];
		}
		const double w = s.real() / static_cast<double>(label.size());
		return w < 0 ? std::sqrt(-w) : std::sqrt(w);
	}
};

/// gple/complex_kernel.cpp:594-670
struct PredictiveComplexKernel
{
	std::unique_ptr<ComplexKernelBase> base;
	double rescale;
	CVec prediction, cutoff_prediction;
	Vec variance;
	std::optional<double> error;
	std::optional<std::array<double, NumComplexParams>> derror;

	PredictiveComplexKernel(const double* feat, const std::size_t M, const TrainingComplexKernel& k, const bool is_deriv, const cplx* test_label):
		rescale(k.rescale)
	{
		const Points T{feat, M};
		base = std::make_unique<ComplexKernelBase>(k.prm, T, k.X, false, is_deriv);
		const Mat& Ks = base->K;
		const CMat& Kts = base->Kt;
		const std::size_t N = k.X.n;
		CVec vc(N);
		for (std::size_t i = 0; i < N; i++)
		{
			vc[i] = std::conj(k.v[i]);
		}
		{
			// complex_kernel.cpp:608
			const CVec a = matvec(Ks, k.v), b = matvec(Kts, vc);
			prediction.resize(M);
			for (std::size_t m = 0; m < M; m++)
			{
				prediction[m] = a[m] + b[m];
			}
		}
		// complex_kernel.cpp:609-642
		variance.resize(M);
		const double prior = sq(k.prm.mag) * (sq(k.prm.sub_mag[0]) + sq(k.prm.sub_mag[1]) + sq(k.prm.noise));
		parallel_for(
			M,
			[&](const std::size_t m)
			{
				cplx t1 = 0, t2 = 0, t3 = 0, t4 = 0;
				for (std::size_t c = 0; c < N; c++)
				{
					const cplx *pc = k.P.col(c), *qc = k.Q.col(c);
					cplx s1 = 0, s2 = 0, s3 = 0, s4 = 0;
					for (std::size_t r = 0; r < N; r++)
					{
						s1 += Ks(m, r) * pc[r];
						s2 += Kts(m, r) * std::conj(pc[r]);
						s3 += Kts(m, r) * qc[r];
						s4 += Ks(m, r) * std::conj(qc[r]);
					}
					t1 += s1 * Ks(m, c);
					t2 += s2 * std::conj(Kts(m, c));
					t3 += s3 * Ks(m, c);
					t4 += s4 * std::conj(Kts(m, c));
				}
				variance[m] = (cplx(prior) - t1 - t2 - t3 - t4).real();
			},
			2
		);
		const Vec cf = cutoff_factor(prediction, variance); // complex_kernel.cpp:643
		cutoff_prediction.resize(M);
		for (std::size_t m = 0; m < M; m++)
		{
			cutoff_prediction[m] = prediction[m] * cf[m] / rescale;
		}
		if (test_label != nullptr)
		{
			CVec lbl(M);
			double e = 0;
			for (std::size_t m = 0; m < M; m++)
			{
				lbl[m] = test_label[m] * rescale;
				e += std::norm(prediction[m] - lbl[m]);
			}
			error = e;
			if (is_deriv)
			{
				// complex_kernel.cpp:648-667; Eigen's dot() conjugates its left operand
				CVec diff(M);
				for (std::size_t m = 0; m < M; m++)
				{
					diff[m] = cutoff_prediction[m] * rescale - lbl[m];
				}
				std::array<double, NumComplexParams> de{};
				for (std::size_t p = 0; p < NumComplexParams; p++)
				{
					CVec dvc(N);
					for (std::size_t i = 0; i < N; i++)
					{
						dvc[i] = std::conj((*k.dv)[p][i]);
					}
					const CVec a = matvec((*base->dK)[p], k.v), b = matvec(Ks, (*k.dv)[p]), c = matvec((*base->dKt)[p], vc), d = matvec(Kts, dvc);
					cplx s = 0;
					for (std::size_t m = 0; m < M; m++)
					{
						s += std::conj(diff[m]) * (a[m] + b[m] + c[m] + d[m]);
					}
					de[p] = 2.0 * s.real();
				}
				derror = de;
			}
		}
	}
};

} // namespace orc
