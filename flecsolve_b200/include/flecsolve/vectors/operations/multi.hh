// Ops policy of vec::multi: fan every operation out over the component vectors.
//
// Reference: flecsolve/vectors/operations/multi.hh:36-277.  Element-wise calls visit the
// components in order; reductions queue one reduction per component and combine on get():
// dot = sum of component dots (:241-252), lp norms combine lp_norm_local of the components
// (:197-223), min/max/inf_norm take the extreme (:167-239).  Because the C ABI defers and
// fuses, the k component kernels of one call (and their k reductions) can share a launch.
#ifndef FLECSOLVE_B200_VECTORS_OPERATIONS_MULTI_HH
#define FLECSOLVE_B200_VECTORS_OPERATIONS_MULTI_HH

#include <algorithm>
#include <array>
#include <cmath>
#include <functional>
#include <tuple>

#include "flecsolve/util/future.hh"

namespace flecsolve::vec::ops {

template<class Data>
struct multi {
	using vec_data = Data;
	using config = typename Data::config;
	using scalar = typename config::scalar;
	using real = typename config::real;
	static constexpr std::size_t num_vecs = config::num_components;
	using indices = std::make_index_sequence<num_vecs>;

	// f(get<I>(first), get<I>(rest)...) for every component I
	template<class F, class... Tuples>
	static constexpr void each(F && f, Tuples &&... ts) {
		each_impl(f, indices{}, ts...);
	}
	template<class F, class... Tuples>
	static constexpr auto collect(F && f, Tuples &&... ts) {
		return collect_impl(f, indices{}, ts...);
	}

	template<class T>
	static void copy(const T & x, vec_data & z) {
		each([](auto & zc, const auto & xc) { zc.copy(xc); }, z.components, x.components);
	}
	static void zero(vec_data & x) {
		each([](auto & c) { c.zero(); }, x.components);
	}
	static void set_to_scalar(scalar a, vec_data & x) {
		each([a](auto & c) { c.set_scalar(a); }, x.components);
	}
	static void scale(scalar a, vec_data & x) {
		each([a](auto & c) { c.scale(a); }, x.components);
	}
	template<class T>
	static void scale(scalar a, const T & x, vec_data & y) {
		each([a](auto & yc, const auto & xc) { yc.scale(a, xc); }, y.components, x.components);
	}
	template<class T0, class T1>
	static void add(const T0 & x, const T1 & y, vec_data & z) {
		each([](auto & zc, const auto & xc, const auto & yc) { zc.add(xc, yc); }, z.components, x.components, y.components);
	}
	template<class T0, class T1>
	static void subtract(const T0 & x, const T1 & y, vec_data & z) {
		each([](auto & zc, const auto & xc, const auto & yc) { zc.subtract(xc, yc); }, z.components, x.components,
		     y.components);
	}
	template<class T0, class T1>
	static void multiply(const T0 & x, const T1 & y, vec_data & z) {
		each([](auto & zc, const auto & xc, const auto & yc) { zc.multiply(xc, yc); }, z.components, x.components,
		     y.components);
	}
	template<class T0, class T1>
	static void divide(const T0 & x, const T1 & y, vec_data & z) {
		each([](auto & zc, const auto & xc, const auto & yc) { zc.divide(xc, yc); }, z.components, x.components,
		     y.components);
	}
	template<class T>
	static void reciprocal(const T & x, vec_data & y) {
		each([](auto & yc, const auto & xc) { yc.reciprocal(xc); }, y.components, x.components);
	}
	template<class T0, class T1>
	static void linear_sum(scalar a, const T0 & x, scalar b, const T1 & y, vec_data & z) {
		each([a, b](auto & zc, const auto & xc, const auto & yc) { zc.linear_sum(a, xc, b, yc); }, z.components,
		     x.components, y.components);
	}
	template<class T0, class T1>
	static void axpy(scalar a, const T0 & x, const T1 & y, vec_data & z) {
		each([a](auto & zc, const auto & xc, const auto & yc) { zc.axpy(a, xc, yc); }, z.components, x.components,
		     y.components);
	}
	template<class T>
	static void axpby(scalar a, scalar b, const T & x, vec_data & z) {
		each([a, b](auto & zc, const auto & xc) { zc.axpby(a, b, xc); }, z.components, x.components);
	}
	template<class T>
	static void abs(const T & x, vec_data & y) {
		each([](auto & yc, const auto & xc) { yc.abs(xc); }, y.components, x.components);
	}
	template<class T>
	static void add_scalar(const T & x, scalar a, vec_data & y) {
		each([a](auto & yc, const auto & xc) { yc.add_scalar(xc, a); }, y.components, x.components);
	}
	static void set_random(vec_data & x, unsigned seed) {
		each([seed](auto & c) { c.set_random(seed); }, x.components);
	}

	// ---- reductions: one future per component, folded at get()
	template<class Futs, class Fold>
	static auto folded(Futs futs, Fold fold) {
		return future_transform{future_vector{std::move(futs)},
		                        [fold](auto && vals) { return std::apply(fold, std::forward<decltype(vals)>(vals)); }};
	}
	static auto min(const vec_data & x) {
		return folded(collect([](const auto & c) { return c.min(); }, x.components),
		              [](auto... v) { return std::min({static_cast<real>(v)...}); });
	}
	static auto max(const vec_data & x) {
		return folded(collect([](const auto & c) { return c.max(); }, x.components),
		              [](auto... v) { return std::max({static_cast<real>(v)...}); });
	}
	static auto inf_norm(const vec_data & x) {
		return folded(collect([](const auto & c) { return c.inf_norm(); }, x.components),
		              [](auto... v) { return std::max({static_cast<real>(v)...}); });
	}
	template<unsigned short p>
	static auto lp_norm(const vec_data & x) {
		auto parts = collect(
			[](const auto & c) { return std::remove_reference_t<decltype(c)>::ops::template lp_norm_local<p>(c.data); },
			x.components);
		return folded(std::move(parts), [](auto... v) {
			const real sum = (v + ...);
			if constexpr (p == 1)
				return sum;
			else if constexpr (p == 2)
				return std::sqrt(sum);
			else
				return std::pow(sum, 1. / p);
		});
	}
	template<class T>
	static auto dot(const vec_data & x, const T & y) {
		return folded(collect([](const auto & xc, const auto & yc) { return xc.dot(yc); }, x.components, y.components),
		              [](auto... v) { return (v + ...); });
	}

private:
	template<class F, std::size_t... I, class... Tuples>
	static constexpr void each_impl(F & f, std::index_sequence<I...>, Tuples &... ts) {
		auto one = [&](auto idx) { std::invoke(f, std::get<decltype(idx)::value>(ts)...); };
		(one(std::integral_constant<std::size_t, I>{}), ...);
	}
	template<class F, std::size_t... I, class... Tuples>
	static constexpr auto collect_impl(F & f, std::index_sequence<I...>, Tuples &... ts) {
		auto one = [&](auto idx) { return std::invoke(f, std::get<decltype(idx)::value>(ts)...); };
		return std::make_tuple(one(std::integral_constant<std::size_t, I>{})...);
	}
};

}
#endif
