// op::shell: wrap any callable (x, y) as an operator; op::I and make_identity copy x into y.
// Reference: flecsolve/operators/shell.hh:27-95.
#ifndef FLECSOLVE_B200_OPERATORS_SHELL_HH
#define FLECSOLVE_B200_OPERATORS_SHELL_HH

#include <utility>

#include "flecsolve/operators/core.hh"
#include "flecsolve/operators/handle.hh"

namespace flecsolve::op {

template<class F, class ivar_t, class ovar_t>
struct shell : base<std::nullptr_t, ivar_t, ovar_t> {
	constexpr shell(F fn, ivar_t, ovar_t) : f(std::move(fn)) {}

	template<class domain_vec, class range_vec>
	constexpr decltype(auto) apply(const domain_vec & x, range_vec & y) const {
		return f(x, y);
	}

protected:
	F f;
};

template<class F, auto I, auto O>
auto make_shell(F && f, variable_t<I>, variable_t<O>) {
	return core<shell<std::decay_t<F>, variable_t<I>, variable_t<O>>>(std::forward<F>(f), variable<I>, variable<O>);
}
template<class F, auto... I, auto... O>
auto make_shell(F && f, multivariable_t<I...>, multivariable_t<O...>) {
	return core<shell<std::decay_t<F>, multivariable_t<I...>, multivariable_t<O...>>>(
		std::forward<F>(f), multivariable<I...>, multivariable<O...>);
}
template<class F>
auto make_shell(F && f) {
	return make_shell(std::forward<F>(f), variable<anon_var::anonymous>, variable<anon_var::anonymous>);
}

template<class F, auto I, auto O>
auto make_shared_shell(F && f, variable_t<I>, variable_t<O>) {
	return make_shared<shell<std::decay_t<F>, variable_t<I>, variable_t<O>>>(std::forward<F>(f), variable<I>,
	                                                                         variable<O>);
}
template<class F, auto... I, auto... O>
auto make_shared_shell(F && f, multivariable_t<I...>, multivariable_t<O...>) {
	return make_shared<shell<std::decay_t<F>, multivariable_t<I...>, multivariable_t<O...>>>(
		std::forward<F>(f), multivariable<I...>, multivariable<O...>);
}
template<class F>
auto make_shared_shell(F && f) {
	return make_shared_shell(std::forward<F>(f), variable<anon_var::anonymous>, variable<anon_var::anonymous>);
}

namespace detail {
struct copy_through {
	template<class X, class Y>
	void operator()(const X & x, Y & y) const {
		y.copy(x);
	}
};
}

template<auto ivar, auto ovar>
auto make_identity(variable_t<ivar>, variable_t<ovar>) {
	return make_shared_shell(detail::copy_through{}, variable<ivar>, variable<ovar>);
}
template<auto... I, auto... O>
auto make_identity(multivariable_t<I...>, multivariable_t<O...>) {
	return make_shared_shell(detail::copy_through{}, multivariable<I...>, multivariable<O...>);
}

static inline const auto I = make_shared_shell(detail::copy_through{});

}
#endif
