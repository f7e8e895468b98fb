// What a Krylov solver is built from: settings, work vectors, operator / preconditioner handles
// and the user's diagnostic callback; krylov_solver binds them into an op::core<solver>.
// Reference: flecsolve/solvers/krylov_parameters.hh:29-126.
#ifndef FLECSOLVE_B200_SOLVERS_KRYLOV_PARAMETERS_HH
#define FLECSOLVE_B200_SOLVERS_KRYLOV_PARAMETERS_HH

#include <memory>
#include <tuple>

#include "flecsolve/operators/handle.hh"
#include "flecsolve/operators/shell.hh"
#include "flecsolve/util/traits.hh"

namespace flecsolve {

// called every iteration with the current iterate and residual norm; true stops the solve
static inline const auto default_diagnostic = [](auto &&, double) { return false; };

template<class Op, class Precond, class Diag>
struct krylov_ops {
	op::handle<Op> A;
	op::handle<Precond> P;
	std::decay_t<Diag> diagnostic;
};

template<class Settings, class Work, class Op, class... Rest>
struct krylov_parameters {
	using work_t = Work;
	using workvec_t = typename std::remove_reference_t<Work>::value_type;
	using real = typename workvec_t::real;
	using scalar = typename workvec_t::scalar;
	using op_type = Op;
	using input_var_t = decltype(op_type::input_var);
	using output_var_t = decltype(op_type::output_var);

	template<class S, class W, class A, class P, class D>
	krylov_parameters(S && s, W && w, op::handle<A> a, op::handle<P> p, D && d)
		: work{std::forward<W>(w)}, settings{std::forward<S>(s)}, ops{a, p, std::forward<D>(d)} {}

	template<class S, class W, class A, class P>
	krylov_parameters(S && s, W && w, op::handle<A> a, op::handle<P> p)
		: krylov_parameters(std::forward<S>(s), std::forward<W>(w), a, p, default_diagnostic) {}

	template<class S, class W, class A>
	krylov_parameters(S && s, W && w, op::handle<A> a)
		: krylov_parameters(std::forward<S>(s), std::forward<W>(w), a, op::make_identity(A::input_var, A::output_var)) {}

	const auto & A() const { return ops.A.get(); }
	const auto & P() const { return ops.P.get(); }

	mutable Work work; // a solver is not re-entrant: apply() is const but scribbles on its work vectors
	Settings settings;
	krylov_ops<Op, Rest...> ops;
};

template<class S, class W, class A>
krylov_parameters(S &&, W &&, op::handle<A>)
	-> krylov_parameters<std::decay_t<S>, std::decay_t<W>, A,
                         typename decltype(op::make_identity(A::input_var, A::output_var))::type,
                         decltype(default_diagnostic)>;
template<class S, class W, class A, class P>
krylov_parameters(S &&, W &&, op::handle<A>, op::handle<P>)
	-> krylov_parameters<std::decay_t<S>, std::decay_t<W>, A, P, decltype(default_diagnostic)>;
template<class S, class W, class A, class P, class D>
krylov_parameters(S &&, W &&, op::handle<A>, op::handle<P>, D &&)
	-> krylov_parameters<std::decay_t<S>, std::decay_t<W>, A, P, D>;

template<template<class> class OpType, class Settings, class Workspace>
struct krylov_solver {
	using settings_type = Settings;

	template<class... Ops>
	auto bind(Ops &&... ops) && {
		return OpType(krylov_parameters(std::move(settings), std::move(workspace), std::forward<Ops>(ops)...));
	}
	template<class... Ops>
	auto bind(Ops &&... ops) & {
		return OpType(krylov_parameters(settings, workspace, std::forward<Ops>(ops)...));
	}
	template<class... Ops>
	auto operator()(Ops &&... ops) && {
		return op::make(std::move(*this).bind(std::forward<Ops>(ops)...));
	}
	template<class... Ops>
	auto operator()(Ops &&... ops) & {
		return op::make(bind(std::forward<Ops>(ops)...));
	}

	settings_type settings;
	Workspace workspace;
};

}
#endif
