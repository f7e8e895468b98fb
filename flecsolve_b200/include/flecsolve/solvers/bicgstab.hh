// BiCGStab (op::bicgstab).
//
// Behavioural parity with flecsolve/solvers/bicgstab.hh:27-188: rho kept as a pair with
// beta = (rho1/rho0)(alpha/omega), breakdown restart when sqrt|rho| < eps |r~|, the early exit
// after the first half step tests |s| against settings.rtol (not the scaled tolerance) and
// reports iters = iter, omega == 0 stops with diverged_breakdown.
// Fused by the deferred queue into, per iteration:
//   { r~.res }  { p = -w v + p ; p = b p + res ; p^ = P p }  { v = A p^ ; r~.v }  { s = -a v + res ; |s|^2 }
//   { s^ = P s }  { t = A s^ ; t.t }  { t.s }  { x += a p^ ; x += w s^ ; res = -w t + s ; |res|^2 }
#ifndef FLECSOLVE_B200_SOLVERS_BICGSTAB_HH
#define FLECSOLVE_B200_SOLVERS_BICGSTAB_HH

#include <cmath>
#include <limits>
#include <stdexcept>

#include "flecsolve/solvers/krylov_parameters.hh"
#include "flecsolve/solvers/solver_settings.hh"

namespace flecsolve::op {

// The iteration is written as a prologue plus one `sweep` per iteration that reports whether the solve is
// over; the order of vector operations and host reads inside a sweep is the reference's (it decides both the
// rounding and which statements the queue can fuse).
template<class Params>
struct bicgstab : base<Params, typename Params::input_var_t, typename Params::output_var_t> {
	using base_t = base<Params, typename Params::input_var_t, typename Params::output_var_t>;
	using base_t::params;
	using real = typename Params::real;

	bicgstab(Params p) : base_t(std::move(p)) {}

	const auto & get_operator() const { return params.A(); }

	template<class DomainVec, class RangeVec>
	solve_info apply(const RangeVec & rhs, DomainVec & sol) const {
		using stop = solve_info::stop_reason;
		const auto & A = params.A();
		const auto & M = params.P();
		const auto & cfg = params.settings;
		auto & monitor = params.ops.diagnostic;
		// work vectors, in the reference's order: residual, shadow residual, direction, A*direction,
		// preconditioned direction, half-step residual, its preconditioned copy, A*that
		// (named references rather than a structured binding: the sweep lambda below captures them)
		auto & w = params.work;
		auto &resid = w[0], &shadow = w[1], &dir = w[2], &a_dir = w[3], &dir_pc = w[4], &half = w[5], &half_pc = w[6],
		     &a_half = w[7];

		solve_info out;
		real rhs_norm = rhs.l2norm().get();
		if (rhs_norm == 0.)
			rhs_norm = 1.;
		out.rhs_norm = rhs_norm;
		const real goal = cfg.rtol * rhs_norm; // float rtol times |b|, as everywhere in flecsolve

		if (!cfg.use_zero_guess) {
			out.sol_norm_initial = sol.l2norm().get();
			A.residual(rhs, sol, resid);
		}
		else {
			out.sol_norm_initial = 0;
			resid.copy(rhs);
			sol.set_scalar(0.);
		}

		real resid_norm = resid.l2norm().get();
		out.res_norm_initial = resid_norm;
		if (resid_norm < goal) {
			out.res_norm_final = resid_norm;
			out.status = stop::converged_rtol;
			return out;
		}

		// scalars of the recurrence; rho starts from the pair {2, 1} like the reference's vector
		struct {
			real alpha = 1.0, beta = 0.0, omega = 1.0, rho_prev = 2.0, rho = 1.0;
		} k;
		real shadow_norm = resid_norm;
		shadow.copy(resid);
		dir.zero();
		a_dir.zero();

		auto finish = [&out](int iterations, stop why) {
			out.iters = iterations;
			out.status = why;
			return true;
		};

		// one iteration; true when the solve is over
		auto sweep = [&](int it) -> bool {
			k.rho = shadow.dot(resid).get();
			if (std::sqrt(std::fabs(k.rho)) < std::numeric_limits<real>::epsilon() * shadow_norm) {
				// the shadow residual lost its grip on the residual: start over from the true residual
				// (rho_prev is deliberately left alone, as in the reference)
				A.residual(rhs, sol, resid);
				shadow.copy(resid);
				resid_norm = resid.l2norm().get();
				k.rho = shadow_norm = resid_norm;
				dir.copy(resid);
				++out.restarts;
				return false;
			}

			if (it > 0) {
				k.beta = (k.rho / k.rho_prev) * (k.alpha / k.omega);
				dir.axpy(-k.omega, a_dir, dir);
				dir.axpy(k.beta, dir, resid);
			}
			else
				dir.copy(resid);

			M.apply(dir, dir_pc);
			A.apply(dir_pc, a_dir);
			const real shadow_dot = shadow.dot(a_dir).get();
			if (shadow_dot == 0.)
				throw std::runtime_error("BiCGSTAB: encountered alpha = 0");
			k.alpha = k.rho / shadow_dot;

			// first half step; note the test against the bare rtol and the iteration count it reports
			half.axpy(-k.alpha, a_dir, resid);
			if (half.l2norm().get() < cfg.rtol) {
				sol.axpy(k.alpha, dir_pc, sol);
				return finish(it, stop::converged_rtol);
			}

			M.apply(half, half_pc);
			A.apply(half_pc, a_half);
			const real denom = a_half.dot(a_half).get();
			const real numer = a_half.dot(half).get();
			k.omega = denom == 0.0 ? real(0.0) : numer / denom;

			sol.axpy(k.alpha, dir_pc, sol);
			sol.axpy(k.omega, half_pc, sol);
			resid.axpy(-k.omega, a_half, half);
			resid_norm = resid.l2norm().get();

			if (monitor(sol, resid_norm))
				return finish(it + 1, stop::converged_user);
			if (resid_norm < goal)
				return finish(it + 1, stop::converged_rtol);
			if (k.omega == 0.0)
				return finish(it + 1, stop::diverged_breakdown);
			k.rho_prev = k.rho;
			return false;
		};

		for (int it = 0; it < cfg.maxiter; ++it)
			if (sweep(it))
				break;

		out.res_norm_final = resid_norm;
		out.sol_norm_final = sol.l2norm().get();
		if (out.iters == 0) // also what an exhausted iteration budget looks like
			out.status = stop::diverged_iters;
		return out;
	}
};
template<class P>
bicgstab(P) -> bicgstab<P>;

}

namespace flecsolve::bicgstab {
static constexpr std::size_t nwork = 8;
struct settings : solver_settings {};
struct options : solver_options {
	using settings_type = settings;
	options(const char * pre) : solver_options(pre) {}
	po::options_description operator()(settings_type & s) { return solver_options::operator()(s); }
};
static inline work_factory<nwork> make_work;

template<class Workspace>
struct solver : krylov_solver<op::bicgstab, settings, Workspace> {
	using base_t = krylov_solver<op::bicgstab, settings, Workspace>;
	template<class W>
	solver(const settings & s, W && w) : base_t{s, std::forward<W>(w)} {}
};
template<class W>
solver(const settings &, W &&) -> solver<std::decay_t<W>>;
}
#endif
