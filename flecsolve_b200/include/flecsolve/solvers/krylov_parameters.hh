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

// the user's per-iteration hook: (current iterate, residual norm) -> true to stop the solve
static inline const auto default_diagnostic = [](auto &&, double) { return false; };

// the three things a Krylov solver is bound to; held by handle so solvers never copy operators
template<class Op, class Precond, class Diag>
struct krylov_ops {
	op::handle<Op> A;
	op::handle<Precond> P;
	std::decay_t<Diag> diagnostic;
};

namespace detail {
// the identity preconditioner an operator of these variables gets when none is given
template<class Op>
auto identity_like() {
	return op::make_identity(Op::input_var, Op::output_var);
}
template<class Op>
using identity_like_t = typename decltype(identity_like<Op>())::type;
}

// settings + work vectors + bound operators = the parameter block of op::cg / gmres / bicgstab ...
// Constructible from (settings, work, A), (settings, work, A, P) or (settings, work, A, P, diagnostic).
template<class Settings, class Work, class Op, class... Rest>
struct krylov_parameters {
	using work_t = Work;
	using workvec_t = typename std::remove_reference_t<Work>::value_type;
	using scalar = typename workvec_t::scalar;
	using real = typename workvec_t::real;
	using op_type = Op;
	using input_var_t = decltype(op_type::input_var);
	using output_var_t = decltype(op_type::output_var);

	template<class S, class W, class A, class Pc, class D>
	krylov_parameters(S && set, W && wrk, op::handle<A> a, op::handle<Pc> pc, D && diag)
		: work{std::forward<W>(wrk)}, settings{std::forward<S>(set)}, ops{a, pc, std::forward<D>(diag)} {}

	template<class S, class W, class A, class Pc>
	krylov_parameters(S && set, W && wrk, op::handle<A> a, op::handle<Pc> pc)
		: krylov_parameters(std::forward<S>(set), std::forward<W>(wrk), a, pc, default_diagnostic) {}

	template<class S, class W, class A>
	krylov_parameters(S && set, W && wrk, op::handle<A> a)
		: krylov_parameters(std::forward<S>(set), std::forward<W>(wrk), a, detail::identity_like<A>()) {}

	const auto & A() const { return ops.A.get(); }
	const auto & P() const { return ops.P.get(); }

	// apply() of a solver is const but scribbles on its work vectors: a solver object is not re-entrant
	mutable Work work;
	Settings settings;
	krylov_ops<Op, Rest...> ops;
};

template<class S, class W, class A, class Pc, class D>
krylov_parameters(S &&, W &&, op::handle<A>, op::handle<Pc>, D &&)
	-> krylov_parameters<std::decay_t<S>, std::decay_t<W>, A, Pc, D>;
template<class S, class W, class A, class Pc>
krylov_parameters(S &&, W &&, op::handle<A>, op::handle<Pc>)
	-> krylov_parameters<std::decay_t<S>, std::decay_t<W>, A, Pc, decltype(default_diagnostic)>;
template<class S, class W, class A>
krylov_parameters(S &&, W &&, op::handle<A>)
	-> krylov_parameters<std::decay_t<S>, std::decay_t<W>, A, detail::identity_like_t<A>, decltype(default_diagnostic)>;

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
