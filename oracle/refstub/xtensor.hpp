// TEST INFRASTRUCTURE ONLY -- oracle/refstub/xtensor.hpp
//
// Stand-in for the small part of xtensor the reference's hot path uses (gple/stdafx.h:56;
// evolve.cpp:21-25, 48-51, 200-372; pes.cpp:71, 156-167; `xt::arange` index ranges everywhere):
// fixed-shape tensors of rank <= 4 with at most 16 coefficients, strided views (`xt::view` with
// integer / `xt::all()` / `xt::newaxis()` slices), pointer adaptors, numpy-style broadcasting
// arithmetic evaluated eagerly, and `xt::arange(n)` as a random-access index range.  xtensor is an
// un-vendored dependency of the reference and is absent from this image; the UNMODIFIED reference
// sources compile against this header (oracle/Makefile.ref).  Row-major like xtensor's default.
#pragma once
#include <array>
#include <cassert>
#include <complex>
#include <cstddef>
#include <iterator>
#include <type_traits>
#include <utility>

namespace xt
{
constexpr std::size_t MaxRank = 4, MaxElems = 16;

struct Shape
{
	std::size_t n = 0;
	std::array<std::size_t, MaxRank> d{};
	std::size_t operator[](std::size_t i) const { return d[i]; }
	std::size_t size() const { return n; }
	std::size_t count() const
	{
		std::size_t c = 1;
		for (std::size_t i = 0; i < n; i++)
		{
			c *= d[i];
		}
		return c;
	}
};

template <std::size_t... N>
struct xshape
{
	static constexpr std::size_t rank = sizeof...(N);
	static constexpr std::size_t count = (std::size_t(1) * ... * N);
	static Shape get()
	{
		Shape s;
		s.n = rank;
		std::size_t i = 0;
		((s.d[i++] = N), ...);
		return s;
	}
};

/// non-owning description of a strided block of coefficients
template <typename T>
struct ndref
{
	T* base = nullptr;
	Shape shape;
	std::array<std::ptrdiff_t, MaxRank> strides{};
};

namespace detail
{
template <typename T>
struct is_complex: std::false_type
{
};
template <typename T>
struct is_complex<std::complex<T>>: std::true_type
{
};
template <typename T>
concept Scalar = std::is_arithmetic_v<std::remove_cvref_t<T>> || is_complex<std::remove_cvref_t<T>>::value;
template <typename A, typename B>
struct promote
{
	using type = std::conditional_t<is_complex<A>::value, A, std::conditional_t<is_complex<B>::value, B, decltype(std::declval<A>() * std::declval<B>())>>;
};
template <typename A, typename B>
using promote_t = typename promote<std::remove_cv_t<A>, std::remove_cv_t<B>>::type;
template <typename To, typename From>
inline To lift(const From& x)
{
	if constexpr (is_complex<To>::value && !is_complex<From>::value)
	{
		return To(static_cast<typename To::value_type>(x), 0);
	}
	else
	{
		return static_cast<To>(x);
	}
}
inline std::array<std::ptrdiff_t, MaxRank> row_major_strides(const Shape& s)
{
	std::array<std::ptrdiff_t, MaxRank> st{};
	std::ptrdiff_t acc = 1;
	for (std::size_t i = s.n; i-- > 0;)
	{
		st[i] = acc;
		acc *= static_cast<std::ptrdiff_t>(s.d[i]);
	}
	return st;
}
inline Shape broadcast(const Shape& a, const Shape& b)
{
	Shape r;
	r.n = a.n > b.n ? a.n : b.n;
	for (std::size_t i = 0; i < r.n; i++)
	{
		// align trailing dimensions
		const std::size_t da = i < r.n - a.n ? 1 : a.d[i - (r.n - a.n)];
		const std::size_t db = i < r.n - b.n ? 1 : b.d[i - (r.n - b.n)];
		assert(da == db || da == 1 || db == 1);
		r.d[i] = da == 1 ? db : da;
	}
	return r;
}
/// offset of the broadcast multi-index `idx` (of shape `to`) inside `r`
template <typename T>
inline std::ptrdiff_t offset_of(const ndref<T>& r, const Shape& to, const std::array<std::size_t, MaxRank>& idx)
{
	std::ptrdiff_t off = 0;
	const std::size_t lead = to.n - r.shape.n;
	for (std::size_t i = 0; i < r.shape.n; i++)
	{
		const std::size_t ii = r.shape.d[i] == 1 ? 0 : idx[i + lead];
		off += static_cast<std::ptrdiff_t>(ii) * r.strides[i];
	}
	return off;
}
inline void advance(std::array<std::size_t, MaxRank>& idx, const Shape& s)
{
	for (std::size_t i = s.n; i-- > 0;)
	{
		if (++idx[i] < s.d[i])
		{
			return;
		}
		idx[i] = 0;
	}
}
} // namespace detail

/// owning result of an expression (runtime shape, at most MaxElems coefficients)
template <typename T>
class ndarray
{
public:
	using value_type = T;
	Shape shp;
	std::array<T, MaxElems> buf{};
	ndref<const T> xref() const { return {buf.data(), shp, detail::row_major_strides(shp)}; }
	ndref<T> xref() { return {buf.data(), shp, detail::row_major_strides(shp)}; }
};

template <typename X>
concept XExpr = requires(const std::remove_cvref_t<X>& x) { x.xref(); };
template <typename X>
using xvalue_t = std::remove_cv_t<std::remove_pointer_t<decltype(std::declval<const std::remove_cvref_t<X>&>().xref().base)>>;

namespace detail
{
template <typename S, typename A, typename B, typename F>
inline ndarray<S> zip(const ndref<A>& a, const ndref<B>& b, F&& f)
{
	ndarray<S> r;
	r.shp = broadcast(a.shape, b.shape);
	const std::size_t total = r.shp.count();
	assert(total <= MaxElems);
	std::array<std::size_t, MaxRank> idx{};
	for (std::size_t l = 0; l < total; l++)
	{
		r.buf[l] = f(a.base[offset_of(a, r.shp, idx)], b.base[offset_of(b, r.shp, idx)]);
		advance(idx, r.shp);
	}
	return r;
}
template <typename S, typename A, typename F>
inline ndarray<S> map(const ndref<A>& a, F&& f)
{
	ndarray<S> r;
	r.shp = a.shape;
	const std::size_t total = r.shp.count();
	assert(total <= MaxElems);
	std::array<std::size_t, MaxRank> idx{};
	for (std::size_t l = 0; l < total; l++)
	{
		r.buf[l] = f(a.base[offset_of(a, r.shp, idx)]);
		advance(idx, r.shp);
	}
	return r;
}
/// dst (broadcast target) = src
template <typename D, typename S>
inline void assign(const ndref<D>& dst, const ndref<S>& src)
{
	const std::size_t total = dst.shape.count();
	std::array<std::size_t, MaxRank> idx{};
	[[maybe_unused]] const Shape b = broadcast(dst.shape, src.shape);
	assert(b.count() == total);
	for (std::size_t l = 0; l < total; l++)
	{
		dst.base[offset_of(dst, dst.shape, idx)] = lift<std::remove_cv_t<D>>(src.base[offset_of(src, dst.shape, idx)]);
		advance(idx, dst.shape);
	}
}
} // namespace detail

#define REFSTUB_XT_BINARY(OP)                                                                                                                              \
	template <XExpr A, XExpr B>                                                                                                                            \
	auto operator OP(const A& a, const B& b)                                                                                                               \
	{                                                                                                                                                      \
		using S = detail::promote_t<xvalue_t<A>, xvalue_t<B>>;                                                                                             \
		return detail::zip<S>(a.xref(), b.xref(), [](const auto& x, const auto& y) { return detail::lift<S>(x) OP detail::lift<S>(y); });                  \
	}                                                                                                                                                      \
	template <XExpr A, detail::Scalar Sc>                                                                                                                  \
	auto operator OP(const A& a, const Sc& s)                                                                                                              \
	{                                                                                                                                                      \
		using S = detail::promote_t<xvalue_t<A>, Sc>;                                                                                                      \
		return detail::map<S>(a.xref(), [&s](const auto& x) { return detail::lift<S>(x) OP detail::lift<S>(s); });                                         \
	}                                                                                                                                                      \
	template <detail::Scalar Sc, XExpr A>                                                                                                                  \
	auto operator OP(const Sc& s, const A& a)                                                                                                              \
	{                                                                                                                                                      \
		using S = detail::promote_t<xvalue_t<A>, Sc>;                                                                                                      \
		return detail::map<S>(a.xref(), [&s](const auto& x) { return detail::lift<S>(s) OP detail::lift<S>(x); });                                         \
	}
REFSTUB_XT_BINARY(+)
REFSTUB_XT_BINARY(-)
REFSTUB_XT_BINARY(*)
REFSTUB_XT_BINARY(/)
#undef REFSTUB_XT_BINARY

/// fixed-shape tensor (xt::xtensor_fixed<T, xt::xshape<...>>)
template <typename T, typename FSH>
class xtensor_fixed;
template <typename T, std::size_t... N>
class xtensor_fixed<T, xshape<N...>>
{
	std::array<T, xshape<N...>::count> buf;

public:
	using value_type = T;
	using shape_type = xshape<N...>;
	static constexpr std::size_t rank = sizeof...(N);
	xtensor_fixed() = default;
	xtensor_fixed(const xtensor_fixed&) = default;
	xtensor_fixed& operator=(const xtensor_fixed&) = default;
	template <XExpr E>
		requires(!std::is_same_v<std::remove_cvref_t<E>, xtensor_fixed>)
	xtensor_fixed(const E& e)
	{
		detail::assign(xref(), e.xref());
	}
	template <XExpr E>
		requires(!std::is_same_v<std::remove_cvref_t<E>, xtensor_fixed>)
	xtensor_fixed& operator=(const E& e)
	{
		// evaluate first: the right-hand side may alias this tensor
		const auto tmp = detail::map<xvalue_t<E>>(e.xref(), [](const auto& x) { return x; });
		detail::assign(xref(), tmp.xref());
		return *this;
	}
	ndref<const T> xref() const { return {buf.data(), shape_type::get(), detail::row_major_strides(shape_type::get())}; }
	ndref<T> xref() { return {buf.data(), shape_type::get(), detail::row_major_strides(shape_type::get())}; }
	template <typename... I>
	T& operator()(I... i)
	{
		static_assert(sizeof...(I) == rank);
		return buf[lin(static_cast<std::size_t>(i)...)];
	}
	template <typename... I>
	const T& operator()(I... i) const
	{
		static_assert(sizeof...(I) == rank);
		return buf[lin(static_cast<std::size_t>(i)...)];
	}
	T* data() { return buf.data(); }
	const T* data() const { return buf.data(); }
	Shape shape() const { return shape_type::get(); }
	std::size_t dimension() const { return rank; }
	template <typename E>
	xtensor_fixed& operator+=(const E& e)
	{
		return *this = *this + e;
	}
	template <typename E>
	xtensor_fixed& operator-=(const E& e)
	{
		return *this = *this - e;
	}
	template <typename E>
	xtensor_fixed& operator*=(const E& e)
	{
		return *this = *this * e;
	}

private:
	template <typename... I>
	static std::size_t lin(I... i)
	{
		constexpr std::array<std::size_t, rank> dims{N...};
		const std::array<std::size_t, rank> idx{i...};
		std::size_t l = 0;
		for (std::size_t a = 0; a < rank; a++)
		{
			assert(idx[a] < dims[a]);
			l = l * dims[a] + idx[a];
		}
		return l;
	}
};

template <typename T, typename SH>
ndarray<T> zeros(const SH&)
{
	ndarray<T> r;
	r.shp = SH::get();
	return r;
}

// ---- views
struct xall_tag
{
};
struct xnewaxis_tag
{
};
inline xall_tag all()
{
	return {};
}
inline xnewaxis_tag newaxis()
{
	return {};
}

/// strided view; owns a copy of the coefficients when it was taken from a temporary
template <typename T>
class xview
{
	using V = std::remove_cv_t<T>;
	T* base_;
	std::ptrdiff_t off_ = 0;
	Shape shp_;
	std::array<std::ptrdiff_t, MaxRank> str_{};
	bool owns_ = false;
	std::array<V, MaxElems> own_{};

	T* origin() const { return owns_ ? const_cast<T*>(own_.data()) : base_; }

public:
	using value_type = V;
	xview(T* base, std::ptrdiff_t off, const Shape& s, const std::array<std::ptrdiff_t, MaxRank>& st): base_(base), off_(off), shp_(s), str_(st) {}
	void take_ownership(const V* src, std::size_t count)
	{
		assert(count <= MaxElems);
		for (std::size_t i = 0; i < count; i++)
		{
			own_[i] = src[i];
		}
		owns_ = true;
		base_ = nullptr;
	}
	xview(const xview&) = default;
	ndref<const V> xref() const { return {origin() + off_, shp_, str_}; }
	ndref<T> xref() { return {origin() + off_, shp_, str_}; }
	/// pointer to the first coefficient of the underlying container (xtensor semantics: add data_offset())
	T* data() const { return origin(); }
	std::size_t data_offset() const { return static_cast<std::size_t>(off_); }
	const Shape& shape() const { return shp_; }
	std::size_t dimension() const { return shp_.n; }
	template <typename... I>
	T& operator()(I... i) const
	{
		const std::array<std::size_t, sizeof...(I)> idx{static_cast<std::size_t>(i)...};
		assert(sizeof...(I) == shp_.n);
		std::ptrdiff_t o = off_;
		for (std::size_t a = 0; a < sizeof...(I); a++)
		{
			assert(idx[a] < shp_.d[a]);
			o += static_cast<std::ptrdiff_t>(idx[a]) * str_[a];
		}
		return origin()[o];
	}
	// assignment writes through
	xview& operator=(const xview& o)
	{
		const auto tmp = detail::map<V>(o.xref(), [](const auto& x) { return x; });
		detail::assign(ndref<T>{origin() + off_, shp_, str_}, tmp.xref());
		return *this;
	}
	template <XExpr E>
	xview& operator=(const E& e)
	{
		const auto tmp = detail::map<xvalue_t<E>>(e.xref(), [](const auto& x) { return x; });
		detail::assign(ndref<T>{origin() + off_, shp_, str_}, tmp.xref());
		return *this;
	}
};

namespace detail
{
template <typename T, typename... S>
inline xview<T> make_view(const ndref<T>& src, const S&... slices)
{
	Shape shp;
	std::array<std::ptrdiff_t, MaxRank> str{};
	std::ptrdiff_t off = 0;
	std::size_t in = 0;
	auto one = [&](const auto& s)
	{
		using ST = std::remove_cvref_t<decltype(s)>;
		if constexpr (std::is_same_v<ST, xall_tag>)
		{
			shp.d[shp.n] = src.shape.d[in];
			str[shp.n] = src.strides[in];
			shp.n++;
			in++;
		}
		else if constexpr (std::is_same_v<ST, xnewaxis_tag>)
		{
			shp.d[shp.n] = 1;
			str[shp.n] = 0;
			shp.n++;
		}
		else
		{
			assert(static_cast<std::size_t>(s) < src.shape.d[in]);
			off += static_cast<std::ptrdiff_t>(s) * src.strides[in];
			in++;
		}
	};
	(one(slices), ...);
	// trailing dimensions not named by a slice are kept whole
	while (in < src.shape.n)
	{
		shp.d[shp.n] = src.shape.d[in];
		str[shp.n] = src.strides[in];
		shp.n++;
		in++;
	}
	return xview<T>(src.base, off, shp, str);
}
} // namespace detail

template <typename E, typename... S>
	requires XExpr<E>
auto view(E&& e, const S&... slices)
{
	if constexpr (std::is_lvalue_reference_v<E>)
	{
		return detail::make_view(e.xref(), slices...);
	}
	else
	{
		// a temporary: keep its coefficients alive inside the view
		const auto dense = detail::map<xvalue_t<E>>(std::as_const(e).xref(), [](const auto& x) { return x; });
		auto v = detail::make_view(dense.xref(), slices...);
		xview<xvalue_t<E>> owned(nullptr, static_cast<std::ptrdiff_t>(v.data_offset()), v.shape(), v.xref().strides);
		owned.take_ownership(dense.buf.data(), dense.shp.count());
		return owned;
	}
}

// ---- adaptors over foreign memory
template <typename T, typename SH>
xview<T> adapt(T* ptr, const SH&)
{
	const Shape s = SH::get();
	return xview<T>(ptr, 0, s, detail::row_major_strides(s));
}
template <typename T, std::size_t N, typename SH>
xview<const T> adapt(const std::array<T, N>& a, const SH&)
{
	const Shape s = SH::get();
	assert(s.count() == N);
	return xview<const T>(a.data(), 0, s, detail::row_major_strides(s));
}

// ---- arange: a random-access range of indices
class index_iterator
{
	std::size_t i = 0;

public:
	using iterator_category = std::random_access_iterator_tag;
	using value_type = std::size_t;
	using difference_type = std::ptrdiff_t;
	using pointer = const std::size_t*;
	using reference = std::size_t;
	index_iterator() = default;
	explicit index_iterator(std::size_t start): i(start) {}
	std::size_t operator*() const { return i; }
	std::size_t operator[](difference_type n) const { return i + static_cast<std::size_t>(n); }
	index_iterator& operator++()
	{
		++i;
		return *this;
	}
	index_iterator operator++(int)
	{
		index_iterator t = *this;
		++i;
		return t;
	}
	index_iterator& operator--()
	{
		--i;
		return *this;
	}
	index_iterator operator--(int)
	{
		index_iterator t = *this;
		--i;
		return t;
	}
	index_iterator& operator+=(difference_type n)
	{
		i = static_cast<std::size_t>(static_cast<difference_type>(i) + n);
		return *this;
	}
	index_iterator& operator-=(difference_type n)
	{
		i = static_cast<std::size_t>(static_cast<difference_type>(i) - n);
		return *this;
	}
	friend index_iterator operator+(index_iterator a, difference_type n) { return a += n; }
	friend index_iterator operator+(difference_type n, index_iterator a) { return a += n; }
	friend index_iterator operator-(index_iterator a, difference_type n) { return a -= n; }
	friend difference_type operator-(const index_iterator& a, const index_iterator& b) { return static_cast<difference_type>(a.i) - static_cast<difference_type>(b.i); }
	friend auto operator<=>(const index_iterator&, const index_iterator&) = default;
};
class index_range
{
	std::size_t lo, hi;

public:
	index_range(std::size_t a, std::size_t b): lo(a), hi(b) {}
	index_iterator begin() const { return index_iterator(lo); }
	index_iterator end() const { return index_iterator(hi); }
	index_iterator cbegin() const { return index_iterator(lo); }
	index_iterator cend() const { return index_iterator(hi); }
	std::size_t size() const { return hi - lo; }
};
template <typename I>
inline index_range arange(I n)
{
	return index_range(0, static_cast<std::size_t>(n));
}
template <typename I, typename J>
inline index_range arange(I a, J b)
{
	return index_range(static_cast<std::size_t>(a), static_cast<std::size_t>(b));
}
} // namespace xt
