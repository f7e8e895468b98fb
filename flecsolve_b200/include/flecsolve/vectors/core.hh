// vec::core<Data, Ops, Config>: the vector seam.
//
// Every solver, integrator and user callback talks to vectors through these members only; all
// work is forwarded to the static functions of the Ops policy.  Interface parity with the
// reference is line-by-line (flecsolve/vectors/core.hh:26-365): same member names, argument
// order and return conventions (reductions return an object with get()).
#ifndef FLECSOLVE_B200_VECTORS_CORE_HH
#define FLECSOLVE_B200_VECTORS_CORE_HH

#include <random>
#include <string_view>
#include <tuple>
#include <type_traits>
#include <utility>

#include "flecsolve/util/traits.hh"
#include "flecsolve/vectors/variable.hh"

namespace flecsolve::vec {

template<template<class> class Data, template<class> class Ops, class Config>
struct core;

namespace detail {
template<template<class> class D, template<class> class O, class C>
std::true_type derives_from_core(const core<D, O, C> &);
std::false_type derives_from_core(...);
}

template<template<class> class Data, template<class> class Ops, class Config>
struct core : Config {
	using config = Config;
	using data_t = Data<Config>;
	using ops = Ops<data_t>;
	using scalar = typename Config::scalar;
	using len_t = typename Config::len_t;
	using var_t = typename Config::var_t;
	static constexpr auto var = Config::var;

	template<class T>
	struct is_vector : decltype(detail::derives_from_core(std::declval<T>())) {};
	template<class T>
	static constexpr bool is_vector_v = is_vector<T>::value;
	template<class... V>
	using if_vectors = std::enable_if_t<(... && is_vector_v<V>), bool>;

	explicit core(data_t d) : data(std::move(d)) {}

	// ---- element-wise: destination is *this
	template<class V, if_vectors<V> = true>
	void copy(const V & src) { ops::copy(src.data, data); } // this = src
	void zero() { ops::zero(data); }
	void set_scalar(scalar val) { ops::set_to_scalar(val, data); }
	template<class V, if_vectors<V> = true>
	void scale(scalar alpha, const V & x) { ops::scale(alpha, x.data, data); } // this = alpha x
	void scale(scalar alpha) { ops::scale(alpha, data); } // this *= alpha
	template<class V1, class V2, if_vectors<V1, V2> = true>
	void add(const V1 & x, const V2 & y) { ops::add(x.data, y.data, data); }
	template<class V1, class V2, if_vectors<V1, V2> = true>
	void subtract(const V1 & x, const V2 & y) { ops::subtract(x.data, y.data, data); }
	template<class V1, class V2, if_vectors<V1, V2> = true>
	void multiply(const V1 & x, const V2 & y) { ops::multiply(x.data, y.data, data); }
	template<class V1, class V2, if_vectors<V1, V2> = true>
	void divide(const V1 & x, const V2 & y) { ops::divide(x.data, y.data, data); }
	template<class V, if_vectors<V> = true>
	void reciprocal(const V & x) { ops::reciprocal(x.data, data); }
	template<class V1, class V2, if_vectors<V1, V2> = true>
	void linear_sum(scalar alpha, const V1 & x, scalar beta, const V2 & y) { // this = alpha x + beta y
		ops::linear_sum(alpha, x.data, beta, y.data, data);
	}
	template<class V1, class V2, if_vectors<V1, V2> = true>
	void axpy(scalar alpha, const V1 & x, const V2 & y) { ops::axpy(alpha, x.data, y.data, data); } // this = alpha x + y
	template<class V, if_vectors<V> = true>
	void axpby(scalar alpha, scalar beta, const V & x) { ops::axpby(alpha, beta, x.data, data); } // this = alpha x + beta this
	template<class V, if_vectors<V> = true>
	void abs(const V & x) { ops::abs(x.data, data); }
	template<class V, if_vectors<V> = true>
	void add_scalar(const V & x, scalar alpha) { ops::add_scalar(x.data, alpha, data); }

	// ---- global reductions: return futures
	auto min() const { return ops::min(data); }
	auto max() const { return ops::max(data); }
	auto l1norm() const { return ops::template lp_norm<1>(data); }
	auto l2norm() const { return ops::template lp_norm<2>(data); }
	template<unsigned short p>
	auto lp_norm() const { return ops::template lp_norm<p>(data); }
	auto inf_norm() const { return ops::inf_norm(data); }
	template<class V, if_vectors<V> = true>
	auto dot(const V & x) const { return ops::dot(data, x.data); }
	auto global_size() const { return ops::global_size(data); }
	std::size_t local_size() const { return ops::local_size(data); }

	void set_random() {
		std::random_device rd;
		ops::set_random(data, rd());
	}
	void set_random(unsigned seed) { ops::set_random(data, seed); }
	void dump(std::string_view prefix) { ops::dump(prefix, data); }

	// ---- component selection: a single vector is its own (only) component
	template<auto ovar>
	constexpr decltype(auto) subset(variable_t<ovar>) const {
		static_assert(ovar == var.value, "variable does not name this vector");
		return *this;
	}
	template<auto ovar>
	constexpr decltype(auto) subset(variable_t<ovar>) {
		static_assert(ovar == var.value, "variable does not name this vector");
		return *this;
	}
	template<auto ovar>
	constexpr decltype(auto) subset(multivariable_t<ovar>) const {
		static_assert(ovar == var.value, "variable does not name this vector");
		return *this;
	}
	template<auto ovar>
	constexpr decltype(auto) subset(multivariable_t<ovar>) {
		static_assert(ovar == var.value, "variable does not name this vector");
		return *this;
	}

	scalar & operator[](len_t i) { return ops::retreive(data, i); }
	const scalar & operator[](len_t i) const { return ops::retreive(data, i); }

	template<class F>
	constexpr decltype(auto) apply(F && f) {
		if constexpr (config::num_components == 1)
			return std::forward<F>(f)(*this);
		else
			return std::apply(std::forward<F>(f), data);
	}

	data_t data;
};

template<template<class> class Data, template<class> class Ops, class Config>
bool operator==(const core<Data, Ops, Config> & a, const core<Data, Ops, Config> & b) {
	return a.data == b.data;
}
template<template<class> class Data, template<class> class Ops, class Config>
bool operator!=(const core<Data, Ops, Config> & a, const core<Data, Ops, Config> & b) {
	return a.data != b.data;
}

}
#endif
