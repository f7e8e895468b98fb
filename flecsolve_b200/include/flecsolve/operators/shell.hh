// op::shell: any callable (x, y) dressed as an operator, so lambdas can stand wherever an operator or
// a preconditioner is expected; op::I / make_identity are the copy-through operator the solvers use as
// "no preconditioner".  Factory names and overload sets as in the reference (flecsolve/operators/shell.hh:27-95):
// make_shell / make_shared_shell take the callable plus, optionally, the input and output variable tags
// (single variables or multivariables).
#ifndef FLECSOLVE_B200_OPERATORS_SHELL_HH
#define FLECSOLVE_B200_OPERATORS_SHELL_HH

#include <type_traits>
#include <utility>

#include "flecsolve/operators/core.hh"
#include "flecsolve/operators/handle.hh"

namespace flecsolve::op {

template<class Callable, class InputTag, class OutputTag>
struct shell : base<std::nullptr_t, InputTag, OutputTag> {
	constexpr shell(Callable c, InputTag, OutputTag) : call_(std::move(c)) {}

	template<class X, class Y>
	constexpr decltype(auto) apply(const X & x, Y & y) const {
		return call_(x, y);
	}

protected:
	Callable call_;
};

namespace detail {
template<class C, class In, class Out>
using shell_for = shell<std::decay_t<C>, In, Out>;

// the two ways a shell is handed out, chosen by the public factories below
template<class C, class In, class Out>
auto shell_by_value(C && c, In in, Out out) {
	return core<shell_for<C, In, Out>>(std::forward<C>(c), in, out);
}
template<class C, class In, class Out>
auto shell_shared(C && c, In in, Out out) {
	return make_shared<shell_for<C, In, Out>>(std::forward<C>(c), in, out);
}

inline constexpr auto anonymous_tag = variable_t<anon_var::anonymous>{};

// y = x
struct copy_through {
	template<class X, class Y>
	void operator()(const X & x, Y & y) const {
		y.copy(x);
	}
};
}

// ---- by value
template<class C, auto In, auto Out>
auto make_shell(C && c, variable_t<In> in, variable_t<Out> out) {
	return detail::shell_by_value(std::forward<C>(c), in, out);
}
template<class C, auto... In, auto... Out>
auto make_shell(C && c, multivariable_t<In...> in, multivariable_t<Out...> out) {
	return detail::shell_by_value(std::forward<C>(c), in, out);
}
template<class C>
auto make_shell(C && c) {
	return detail::shell_by_value(std::forward<C>(c), detail::anonymous_tag, detail::anonymous_tag);
}

// ---- as a shared handle
template<class C, auto In, auto Out>
auto make_shared_shell(C && c, variable_t<In> in, variable_t<Out> out) {
	return detail::shell_shared(std::forward<C>(c), in, out);
}
template<class C, auto... In, auto... Out>
auto make_shared_shell(C && c, multivariable_t<In...> in, multivariable_t<Out...> out) {
	return detail::shell_shared(std::forward<C>(c), in, out);
}
template<class C>
auto make_shared_shell(C && c) {
	return detail::shell_shared(std::forward<C>(c), detail::anonymous_tag, detail::anonymous_tag);
}

// ---- identity
template<auto In, auto Out>
auto make_identity(variable_t<In> in, variable_t<Out> out) {
	return detail::shell_shared(detail::copy_through{}, in, out);
}
template<auto... In, auto... Out>
auto make_identity(multivariable_t<In...> in, multivariable_t<Out...> out) {
	return detail::shell_shared(detail::copy_through{}, in, out);
}
static inline const auto I = make_shared_shell(detail::copy_through{});

}
#endif
