// op::base / op::core: the operator seam.
//
// An operator is a policy P with `apply(x, y)`; op::core<P> adds the public entry points the
// solvers call -- apply, operator() and residual -- each of which first narrows x and y to the
// operator's input/output variables (vec::multi::subset).  Interface as in the reference
// (flecsolve/operators/core.hh:32-161).
#ifndef FLECSOLVE_B200_OPERATORS_CORE_HH
#define FLECSOLVE_B200_OPERATORS_CORE_HH

#include <initializer_list>
#include <memory>
#include <type_traits>
#include <utility>

#include "flecsolve/operators/traits.hh"
#include "flecsolve/util/traits.hh"
#include "flecsolve/vectors/traits.hh"
#include "flecsolve/vectors/variable.hh"

namespace flecsolve::op {

template<class Params = std::nullptr_t,
         class ivar = variable_t<anon_var::anonymous>,
         class ovar = variable_t<anon_var::anonymous>>
struct base {
	static constexpr auto input_var = ivar{};
	static constexpr auto output_var = ovar{};
	using params_t = Params;

	template<class Head, class... Tail,
	         std::enable_if_t<!std::is_same_v<std::decay_t<Head>, base>, bool> = true,
	         std::enable_if_t<std::is_constructible_v<params_t, Head, Tail...> ||
	                              is_aggregate_initializable_v<params_t, Head, Tail...>,
	                          bool> = true>
	base(Head && h, Tail &&... t) : params{std::forward<Head>(h), std::forward<Tail>(t)...} {}

	template<class Head, class... Tail,
	         std::enable_if_t<std::is_constructible_v<params_t, std::initializer_list<Head>, Tail...>, bool> = true>
	base(std::initializer_list<Head> head, Tail &&... tail) : params{head, std::forward<Tail>(tail)...} {}

	template<class T = params_t, class = std::enable_if_t<std::is_null_pointer_v<T>>>
	base() : params{nullptr} {}

	params_t & get_params() { return params; }
	const params_t & get_params() const { return params; }

	template<op::label tag, class T>
	auto get_parameters(const T &) const {
		return nullptr;
	}
	template<class T>
	void reset(const T &) const {}

	auto & get_operator() { return *this; }
	const auto & get_operator() const { return *this; }

protected:
	params_t params;
};

// The public face of an operator: everything callers use (solvers, integrators, user code) is here, and all
// of it funnels into the policy's apply(x, y) after both vectors have been narrowed to the variables the
// operator declares (a plain vector is its own subset; a vec::multi hands out the matching components).
template<class Policy>
struct core : Policy {
	using policy_type = Policy;
	using params_t = typename Policy::params_t;
	static constexpr auto input_var = Policy::input_var;
	static constexpr auto output_var = Policy::output_var;

private:
	template<class First, class... Rest>
	static constexpr bool builds_policy_from =
		!std::is_same_v<std::decay_t<First>, core> && std::is_constructible_v<Policy, First, Rest...>;
	template<class V>
	using if_vector = std::enable_if_t<is_vector_v<V>, bool>;

public:
	// whatever constructs the policy constructs the operator (copying an operator is left to the compiler)
	template<class First, class... Rest, std::enable_if_t<builds_policy_from<First, Rest...>, bool> = true>
	core(First && first, Rest &&... rest) : Policy{std::forward<First>(first), std::forward<Rest>(rest)...} {}

	template<class Elem, class... Rest,
	         std::enable_if_t<std::is_constructible_v<Policy, std::initializer_list<Elem>, Rest...>, bool> = true>
	core(std::initializer_list<Elem> list, Rest &&... rest) : Policy{list, std::forward<Rest>(rest)...} {}

	// y = Op(x)
	template<class X, class Y, if_vector<X> = true, if_vector<Y> = true>
	decltype(auto) apply(const X & x, Y & y) const {
		decltype(auto) out = y.subset(output_var);
		decltype(auto) in = x.subset(input_var);
		return Policy::apply(in, out);
	}
	template<class X, class Y, if_vector<X> = true, if_vector<Y> = true>
	decltype(auto) operator()(const X & x, Y & y) const {
		return this->apply(x, y);
	}

	// r = b - Op(x): Op(x) lands in r, then r is subtracted from b in place (the aliased subtract of
	// the reference, operators/core.hh:133-142, so the device sees one fused statement after the SpMV)
	template<class B, class X, class R, if_vector<B> = true, if_vector<X> = true, if_vector<R> = true>
	void residual(const B & b, const X & x, R & r) const {
		this->apply(x, r);
		decltype(auto) rhs = b.subset(output_var);
		decltype(auto) res = r.subset(output_var);
		res.subtract(rhs, res);
	}
};

template<class P>
auto make(P && policy) {
	return core<std::decay_t<P>>(std::forward<P>(policy));
}
template<class P, class... Args>
auto make_shared1(Args &&... args) {
	return std::make_shared<core<P>>(std::forward<Args>(args)...);
}

template<class T>
struct is_operator : std::false_type {};
template<class P>
struct is_operator<core<P>> : std::true_type {};
template<class T>
inline constexpr bool is_operator_v = is_operator<T>::value;

}
#endif
