// mat::sparse<Data, Ops, Config>: a matrix is an operator whose apply() is Ops::spmv.
// Reference: flecsolve/matrices/seq.hh:35-63.
#ifndef FLECSOLVE_B200_MATRICES_SPARSE_HH
#define FLECSOLVE_B200_MATRICES_SPARSE_HH

#include "flecsolve/operators/core.hh"
#include "flecsolve/vectors/traits.hh"

namespace flecsolve::mat {

template<template<class> class Data, template<class> class Ops, class Config>
struct sparse : op::base<> {
	using config = Config;
	using scalar = typename config::scalar;
	using size = typename config::size;
	using data_t = Data<Config>;
	using ops = Ops<data_t>;

	sparse() = default;
	sparse(data_t && d) : data{std::move(d)} {}

	template<class D, class R, std::enable_if_t<is_vector_v<D> && is_vector_v<R>, bool> = true>
	constexpr void apply(const D & x, R & y) const {
		mult(x, y);
	}
	template<class X, class Y, std::enable_if_t<is_vector_v<X> && is_vector_v<Y>, bool> = true>
	constexpr void mult(const X & x, Y & y) const {
		ops::spmv(x, data, y);
	}

	data_t data;
};

}
#endif
