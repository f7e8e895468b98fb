// Restarted GMRES with modified Gram-Schmidt and Givens rotations (op::gmres).
//
// Behavioural parity with flecsolve/solvers/gmres.hh:39-361: 104 work vectors
// (3 + krylov_dim_bound + 1), right preconditioning by default, the residual estimate is
// |dw[k+1]| of the rotated least-squares right-hand side, the diagnostic sees the not yet
// corrected x, restart only when k reaches max_krylov_dim before the last iteration, Givens
// rotations after Bindel-Demmel-Kahan-Marques with the reference's sign convention.
// The Hessenberg matrix is column-major with leading dimension max_dim + 1.
// With the deferred queue one MGS step {v -= h q_j ; v.q_{j+1}} is a single kernel.
#ifndef FLECSOLVE_B200_SOLVERS_GMRES_HH
#define FLECSOLVE_B200_SOLVERS_GMRES_HH

#include <algorithm>
#include <cmath>
#include <istream>
#include <limits>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "flecsolve/solvers/krylov_parameters.hh"
#include "flecsolve/solvers/solver_settings.hh"

namespace flecsolve::gmres {

static constexpr int krylov_dim_bound = 100;
static constexpr std::size_t nwork = (krylov_dim_bound + 1) + 3;
enum class precond_side { left, right };

inline std::istream & operator>>(std::istream & in, precond_side & s) {
	std::string tok;
	in >> tok;
	if (tok == "left")
		s = precond_side::left;
	else if (tok == "right")
		s = precond_side::right;
	else
		in.setstate(std::ios_base::failbit);
	return in;
}
inline std::ostream & operator<<(std::ostream & os, const precond_side & s) {
	return os << (s == precond_side::left ? "left" : "right");
}

struct settings : solver_settings {
	int max_krylov_dim;
	precond_side pre_side;
	bool restart;

	void validate() {
		if (max_krylov_dim < 0)
			max_krylov_dim = maxiter;
		if (max_krylov_dim > krylov_dim_bound)
			throw std::invalid_argument("GMRES: max_krylov_dim is larger than bound");
		if (!restart && maxiter > max_krylov_dim)
			throw std::invalid_argument(
				"GMRES: maxiters must be less than or equal to max_krylov_dim when not using restart");
	}
};

}

namespace flecsolve::op {

template<class Params>
struct gmres : base<Params, typename Params::input_var_t, typename Params::output_var_t> {
	using base_t = base<Params, typename Params::input_var_t, typename Params::output_var_t>;
	using real = typename Params::real;
	using base_t::params;

	gmres(Params p) : base_t(std::move(p)) { reset(); }

	const auto & get_operator() const { return params.A(); }

	void reset() {
		params.settings.validate();
		const int max_dim = std::min(params.settings.max_krylov_dim, params.settings.maxiter);
		ld = static_cast<std::size_t>(max_dim) + 1;
		hess.assign(ld * ld, 0.0);
		cosvec.assign(ld, 0.0);
		sinvec.assign(ld, 0.0);
		dwvec.assign(ld, 0.0);
		dyvec.assign(ld, 0.0);
	}

	template<class DomainVec, class RangeVec>
	solve_info apply(const RangeVec & b, DomainVec & x) const {
		using namespace ::flecsolve::gmres;
		using stop = solve_info::stop_reason;
		const auto & A = params.A();
		const auto & P = params.P();
		auto & diagnostic = params.ops.diagnostic;
		const auto & settings = params.settings;
		auto & work = params.work;
		const bool right = settings.pre_side == precond_side::right;

		solve_info info;
		auto & res = work[0];
		auto & z = work[1];
		auto & v = work[2];
		auto * basis = work.data() + 3;

		auto b_norm = b.l2norm().get();
		info.rhs_norm = b_norm;
		if (b_norm < std::numeric_limits<real>::epsilon()) // zero rhs: converge relative to 1
			b_norm = 1.0;
		const real terminate_tol = settings.rtol * b_norm;

		if (settings.use_zero_guess)
			x.set_scalar(0.);

		// (preconditioned) residual into `res`
		auto residual_into_res = [&](bool zero_guess) {
			auto & target = right ? res : basis[0];
			if (zero_guess)
				target.copy(b);
			else
				A.residual(b, x, target);
			if (!right)
				P.apply(basis[0], res);
		};
		residual_into_res(settings.use_zero_guess);

		const real beta = res.l2norm().get();
		info.res_norm_initial = beta;
		if (beta < terminate_tol) {
			info.res_norm_final = beta;
			info.status = stop::converged_rtol;
			return info;
		}

		res.scale(1.0 / beta);
		basis[0].copy(res);
		dwvec[0] = beta; // rhs of the least-squares problem: beta e_1
		auto v_norm = beta;

		int k = 0;
		for (int iter = 0; iter < settings.maxiter; iter++) {
			if (right) {
				P.apply(basis[k], z);
				A.apply(z, v);
			}
			else {
				A.apply(basis[k], z);
				P.apply(z, v);
			}

			orthogonalize(v, basis, k + 1);
			v_norm = H(k + 1, k);
			if (v_norm != 0.0)
				v.scale(1.0 / v_norm);
			basis[k + 1].copy(v);

			for (int i = 0; i < k; i++)
				rotate(i, k);
			if (v_norm != 0.0) {
				make_rotation(k);
				rotate(k, k);
				H(k + 1, k) = 0.0; // exactly zero, not round-off
				const real g = dwvec[k];
				dwvec[k] = cosvec[k] * g;
				dwvec[k + 1] = -sinvec[k] * g;
			}
			v_norm = std::fabs(dwvec[k + 1]);
			++k;

			if (diagnostic(x, v_norm)) {
				info.status = stop::converged_user;
				info.iters = iter + 1;
				break;
			}
			if (v_norm < terminate_tol) {
				info.status = stop::converged_rtol;
				info.iters = iter + 1;
				break;
			}

			if (k == settings.max_krylov_dim && iter != settings.maxiter - 1) {
				back_solve(k - 1);
				correct(k - 1, P, basis, z, v, x);
				residual_into_res(false);
				const real betar = res.l2norm().get();
				res.scale(1.0 / betar);
				basis[0].copy(res);
				dwvec[0] = betar;
				++info.restarts;
				k = 0;
			}
		}

		if (k > 0) {
			back_solve(k - 1);
			correct(k - 1, P, basis, z, v, x);
		}

		info.res_norm_final = v_norm;
		info.sol_norm_final = x.l2norm().get();
		if (info.iters == 0)
			info.status = stop::diverged_iters;
		return info;
	}

protected:
	real & H(int i, int j) const { return hess[static_cast<std::size_t>(i) + static_cast<std::size_t>(j) * ld]; }

	// x += (P) sum_i y_i q_i
	template<class Op, class Basis, class W, class T>
	void correct(int nr, Op & P, Basis * basis, W & z, W & v, T & x) const {
		if (params.settings.pre_side == ::flecsolve::gmres::precond_side::right) {
			z.set_scalar(0.0);
			for (int i = 0; i <= nr; i++)
				z.axpy(dyvec[i], basis[i], z);
			P.apply(z, v);
			x.axpy(1.0, v, x);
		}
		else {
			for (int i = 0; i <= nr; i++)
				x.axpy(dyvec[i], basis[i], x);
		}
	}

	// modified Gram-Schmidt against basis[0..k), filling column k-1 of the Hessenberg matrix
	template<class T, class Basis>
	void orthogonalize(T & v, Basis * basis, int k) const {
		for (int j = 0; j < k; j++) {
			const double h_jk = v.dot(basis[j]).get();
			v.axpy(-h_jk, basis[j], v);
			H(j, k - 1) = h_jk;
		}
		H(k, k - 1) = v.l2norm().get();
	}

	void rotate(int i, int k) const {
		const real a = H(i, k), b = H(i + 1, k), c = cosvec[i], s = sinvec[i];
		H(i, k) = c * a + s * b;
		H(i + 1, k) = -s * a + c * b;
	}

	// rotation zeroing H(k+1, k): Bindel, Demmel, Kahan, Marques, "On computing Givens rotations
	// reliably and efficiently", algorithm 1
	void make_rotation(int k) const {
		const real f = H(k, k), g = H(k + 1, k);
		real c, s;
		if (g == 0.0) {
			c = 1.0;
			s = 0.0;
		}
		else if (f == 0.0) {
			c = 0.0;
			s = (g < 0.0) ? -1.0 : 1.0;
		}
		else {
			real r = std::sqrt(f * f + g * g);
			r = 1.0 / r;
			c = std::fabs(f) * r;
			s = std::copysign(g * r, f);
		}
		cosvec[k] = c;
		sinvec[k] = s;
	}

	void back_solve(int nr) const {
		dyvec[nr] = dwvec[nr] / H(nr, nr);
		for (int k = nr - 1; k >= 0; k--) {
			dyvec[k] = dwvec[k];
			for (int i = k + 1; i <= nr; i++)
				dyvec[k] -= H(k, i) * dyvec[i];
			dyvec[k] = dyvec[k] / H(k, k);
		}
	}

	mutable std::size_t ld = 0;
	mutable std::vector<real> hess, sinvec, cosvec, dwvec, dyvec;
};
template<class P>
gmres(P) -> gmres<P>;
}

namespace flecsolve::gmres {

// INI keys beside solver_options': max-krylov-dim (-1), pre-side (right), restart (false); reference :406-422
struct options : solver_options {
	using settings_type = settings;
	options(const char * pre) : solver_options(pre) {}
	po::options_description operator()(settings_type & s) {
		auto desc = solver_options::operator()(s);
		desc.add_options()
			(label("max-krylov-dim").c_str(), po::value<int>(&s.max_krylov_dim)->default_value(-1), "maximum krylov dimension")
			(label("pre-side").c_str(), po::value<precond_side>(&s.pre_side)->default_value(precond_side::right),
			 "preconditioner side")
			(label("restart").c_str(), po::value<bool>(&s.restart)->default_value(false), "should restart");
		return desc;
	}
};

static inline work_factory<nwork> make_work;

template<class Work>
struct solver : krylov_solver<op::gmres, settings, Work> {
	using base_t = krylov_solver<op::gmres, settings, Work>;
	template<class W>
	solver(const settings & s, W && w) : base_t{s, std::forward<W>(w)} {}
};
template<class W>
solver(const settings &, W &&) -> solver<std::decay_t<W>>;

}
#endif
