// Preconditioned conjugate gradients (op::cg) and flexible CG (op::fcg).
//
// Behavioural parity with the reference loops (flecsolve/solvers/cg.hh:28-244) is the point of
// this file: same operation order per iteration, same float tolerance arithmetic
// (terminate_tol = float(rtol) * |b|, |b| == 0 -> 1, strict `<`), convergence tested before the
// preconditioner, diagnostics called with the updated iterate, `iters == 0 => diverged_iters`.
// Every vector call below lands in the C ABI's deferred queue; one iteration becomes
//   { w = A p ; w.p }   { x += a p ; r -= a w ; |r|^2 }   { z = P r ; r.z }   { p = b p + z }
// = 4 kernels + 3 host reads of a scalar.
#ifndef FLECSOLVE_B200_SOLVERS_CG_HH
#define FLECSOLVE_B200_SOLVERS_CG_HH

#include <array>
#include <iostream>

#include "flecsolve/solvers/krylov_parameters.hh"
#include "flecsolve/solvers/solver_settings.hh"

namespace flecsolve::op {

namespace detail {
// shared prologue of cg/fcg: norms, initial residual, early exit.  Returns true when done.
template<class Op, class Settings, class B, class X, class R, class Real>
bool krylov_start(const Op & A, const Settings & settings, const B & b, X & x, R & r, solve_info & info,
                  Real & terminate_tol, Real & current_res) {
	Real b_norm = b.l2norm().get();
	if (b_norm == 0.0)
		b_norm = 1.0;
	terminate_tol = settings.rtol * b_norm;
	info.rhs_norm = b_norm;

	if (settings.use_zero_guess) {
		info.sol_norm_initial = 0;
		x.set_scalar(0.);
		r.copy(b);
	}
	else {
		info.sol_norm_initial = x.l2norm().get();
		A.residual(b, x, r);
	}
	current_res = r.l2norm().get();
	if (current_res < terminate_tol) {
		info.res_norm_initial = current_res;
		info.res_norm_final = current_res;
		info.status = solve_info::stop_reason::converged_rtol;
		return true;
	}
	return false;
}
}

template<class Params>
struct cg : base<Params, typename Params::input_var_t, typename Params::output_var_t> {
	using base_t = base<Params, typename Params::input_var_t, typename Params::output_var_t>;
	using base_t::params;
	using real = typename Params::real;
	using scalar = typename Params::scalar;

	cg(Params p) : base_t(std::move(p)) {}

	const auto & get_operator() const { return params.A(); }

	template<class DomainVec, class RangeVec>
	solve_info apply(const RangeVec & rhs, DomainVec & sol) const {
		using scalar = typename DomainVec::scalar;
		using real = typename DomainVec::real;
		using stop = solve_info::stop_reason;

		const auto & A = params.A();
		const auto & M = params.P();
		const auto & cfg = params.settings;
		auto & monitor = params.ops.diagnostic;
		// work vectors: residual, preconditioned residual, search direction, A * direction
		auto & w = params.work;
		auto &resid = w[0], &precond = w[1], &dir = w[2], &a_dir = w[3];

		solve_info out;
		real goal, resid_norm;
		if (detail::krylov_start(A, cfg, rhs, sol, resid, out, goal, resid_norm))
			return out;

		M.apply(resid, precond);
		scalar rho = precond.dot(resid).get();
		dir.copy(precond);

		// one iteration = three host reads: <Ap, p>, |r|, <r, z> -- each closes one fused kernel
		for (int it = 0; it < cfg.maxiter; ++it) {
			A.apply(dir, a_dir);
			const scalar curvature = a_dir.dot(dir).get();
			if (curvature <= 0.0)
				std::cerr << "PCG: negative curvature encountered!" << std::endl;
			const scalar step = rho / curvature;

			sol.axpy(step, dir, sol);
			resid.axpy(-step, a_dir, resid);
			resid_norm = resid.l2norm().get();

			// the callback sees the updated iterate and may end the solve; then the tolerance test
			const bool user_stop = monitor(sol, resid_norm);
			if (user_stop || resid_norm < goal) {
				out.iters = it + 1;
				out.status = user_stop ? stop::converged_user : stop::converged_rtol;
				break;
			}

			M.apply(resid, precond);
			const scalar rho_before = rho;
			rho = resid.dot(precond).get();
			dir.axpy(rho / rho_before, dir, precond); // direction = beta * direction + z
		}

		out.res_norm_final = resid_norm;
		out.sol_norm_final = sol.l2norm().get();
		if (out.iters == 0) // an exhausted iteration budget leaves iters at 0, as in the reference
			out.status = stop::diverged_iters;
		return out;
	}
};
template<class P>
cg(P) -> cg<P>;

template<class Params>
struct fcg : base<Params, typename Params::input_var_t, typename Params::output_var_t> {
	using base_t = base<Params, typename Params::input_var_t, typename Params::output_var_t>;
	using base_t::params;
	using real = typename Params::real;
	using scalar = typename Params::scalar;

	fcg(Params p) : base_t(std::move(p)) {}

	const auto & get_operator() const { return params.A(); }

	template<class DomainVec, class RangeVec>
	solve_info apply(const RangeVec & b, DomainVec & u) const {
		using scalar = typename DomainVec::scalar;
		using real = typename DomainVec::real;
		using stop = solve_info::stop_reason;

		solve_info info;
		const auto & A = params.A();
		const auto & P = params.P();
		auto & diagnostic = params.ops.diagnostic;
		const auto & settings = params.settings;
		auto & [r, v, w, q, d] = params.work;

		real terminate_tol, current_res;
		if (detail::krylov_start(A, settings, b, u, r, info, terminate_tol, current_res))
			return info;

		scalar rho = 0;
		for (int iter = 0; iter < settings.maxiter; iter++) {
			P.apply(r, v);
			A.apply(v, w);
			const scalar alpha = v.dot(r).get();
			const scalar beta = v.dot(w).get();
			if (iter == 0) {
				d.copy(v);
				q.copy(w);
				rho = beta;
			}
			else {
				const scalar gamma = v.dot(q).get();
				d.axpy((-gamma) / rho, d, v);
				q.axpy((-gamma) / rho, q, w);
				rho = beta - (gamma * gamma) / rho;
			}
			u.axpy(alpha / rho, d, u);
			r.axpy(-alpha / rho, q, r);

			current_res = r.l2norm().get();
			if (diagnostic(u, current_res)) {
				info.iters = iter + 1;
				info.status = stop::converged_user;
				break;
			}
			if (current_res < terminate_tol) {
				info.iters = iter + 1;
				info.status = stop::converged_rtol;
				break;
			}
		}
		info.res_norm_final = current_res;
		info.sol_norm_final = u.l2norm().get();
		if (info.iters == 0)
			info.status = stop::diverged_iters;
		return info;
	}
};
template<class P>
fcg(P) -> fcg<P>;

}

namespace flecsolve::cg {
static constexpr std::size_t nwork = 4;
using settings = solver_settings;
using options = solver_options;
static inline work_factory<nwork> make_work;

template<class Work>
struct solver : krylov_solver<op::cg, settings, Work> {
	using base_t = krylov_solver<op::cg, settings, Work>;
	template<class W>
	solver(const settings & set, W && w) : base_t{set, std::forward<W>(w)} {}
};
template<class W>
solver(const settings &, W &&) -> solver<std::decay_t<W>>;
}

namespace flecsolve::fcg {
static constexpr std::size_t nwork = 5;
using settings = solver_settings;
using options = solver_options;
static inline work_factory<nwork> make_work;

template<class Work>
struct solver : krylov_solver<op::fcg, settings, Work> {
	using base_t = krylov_solver<op::fcg, settings, Work>;
	template<class W>
	solver(const settings & set, W && w) : base_t{set, std::forward<W>(w)} {}
};
template<class W>
solver(const settings &, W &&) -> solver<std::decay_t<W>>;
}
#endif
