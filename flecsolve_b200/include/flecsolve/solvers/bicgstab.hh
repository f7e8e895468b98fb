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

template<class Params>
struct bicgstab : base<Params, typename Params::input_var_t, typename Params::output_var_t> {
	using base_t = base<Params, typename Params::input_var_t, typename Params::output_var_t>;
	using base_t::params;
	using real = typename Params::real;

	bicgstab(Params p) : base_t(std::move(p)) {}

	const auto & get_operator() const { return params.A(); }

	template<class DomainVec, class RangeVec>
	solve_info apply(const RangeVec & b, DomainVec & x) const {
		using stop = solve_info::stop_reason;
		solve_info info;
		const auto & A = params.A();
		const auto & P = params.P();
		auto & diagnostic = params.ops.diagnostic;
		const auto & settings = params.settings;
		auto & [res, r_tilde, p, v, p_hat, s, s_hat, t] = params.work;

		real b_norm = b.l2norm().get();
		if (b_norm == 0.)
			b_norm = 1.;
		const real terminate_tol = settings.rtol * b_norm;
		info.rhs_norm = b_norm;

		if (settings.use_zero_guess) {
			info.sol_norm_initial = 0;
			res.copy(b);
			x.set_scalar(0.);
		}
		else {
			info.sol_norm_initial = x.l2norm().get();
			A.residual(b, x, res);
		}

		real res_norm = res.l2norm().get();
		real r_tilde_norm = res_norm;
		info.res_norm_initial = res_norm;
		if (res_norm < terminate_tol) {
			info.status = stop::converged_rtol;
			info.res_norm_final = res_norm;
			return info;
		}

		real alpha = 1.0, beta = 0.0, omega = 1.0;
		real rho_old = 2.0, rho_new = 1.0; // the reference starts from std::vector<real> rho{2, 1.0}
		r_tilde.copy(res); // shadow residual: the initial residual
		p.zero();
		v.zero();

		for (int iter = 0; iter < settings.maxiter; iter++) {
			rho_new = r_tilde.dot(res).get();

			const real angle = std::sqrt(std::fabs(rho_new));
			if (angle < std::numeric_limits<real>::epsilon() * r_tilde_norm) {
				// r~ has become orthogonal to the residual: restart from the true residual
				A.residual(b, x, res);
				r_tilde.copy(res);
				res_norm = res.l2norm().get();
				rho_new = r_tilde_norm = res_norm;
				p.copy(res);
				++info.restarts;
				continue;
			}

			if (iter == 0) {
				p.copy(res);
			}
			else {
				beta = (rho_new / rho_old) * (alpha / omega);
				p.axpy(-omega, v, p);
				p.axpy(beta, p, res);
			}

			P.apply(p, p_hat);
			A.apply(p_hat, v);

			alpha = r_tilde.dot(v).get();
			if (alpha == 0.)
				throw std::runtime_error("BiCGSTAB: encountered alpha = 0");
			alpha = rho_new / alpha;

			s.axpy(-alpha, v, res);
			const real s_norm = s.l2norm().get();
			if (s_norm < settings.rtol) { // early convergence on the half step
				x.axpy(alpha, p_hat, x);
				info.iters = iter;
				info.status = stop::converged_rtol;
				break;
			}

			P.apply(s, s_hat);
			A.apply(s_hat, t);

			const real t_sqnorm = t.dot(t).get();
			const real t_dot_s = t.dot(s).get();
			omega = (t_sqnorm == 0.0) ? 0.0 : t_dot_s / t_sqnorm;

			x.axpy(alpha, p_hat, x);
			x.axpy(omega, s_hat, x);
			res.axpy(-omega, t, s);

			res_norm = res.l2norm().get();
			if (diagnostic(x, res_norm)) {
				info.status = stop::converged_user;
				info.iters = iter + 1;
				break;
			}
			if (res_norm < terminate_tol) {
				info.status = stop::converged_rtol;
				info.iters = iter + 1;
				break;
			}
			if (omega == 0.0) {
				info.iters = iter + 1;
				info.status = stop::diverged_breakdown;
				break;
			}
			rho_old = rho_new;
		}

		info.res_norm_final = res_norm;
		info.sol_norm_final = x.l2norm().get();
		if (info.iters == 0)
			info.status = stop::diverged_iters;
		return info;
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
