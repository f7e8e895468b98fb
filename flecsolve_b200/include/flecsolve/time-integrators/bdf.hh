// Variable-step BDF1-6 / Crank-Nicolson integrator with predictor-based step control.
//
// Behavioural parity with flecsolve/time-integrators/bdf.hh:53-946: the same public protocol
//     ti.advance(dt, first_step, u, unew);  ok = ti.check_solution();
//     if (ok) { ti.update(); swap(u, unew); }   dt = ti.get_next_dt(ok);
// the same source-term coefficients (incl. the BDF6 a6 expression exactly as the reference writes it,
// bdf.hh:431-432), predictors (forward Euler, AB2, leapfrog), time-derivative estimates, scaled
// local-truncation-error norm, PI / deadbeat step controllers and reject accounting.
// Everything the integrator does to vectors goes through vec::core, i.e. into the C ABI's deferred
// queue: the source-term chains f = a1 u_n ; f += a2 u_{n-1} ; ... become one fused kernel.
#ifndef FLECSOLVE_B200_TIME_INTEGRATORS_BDF_HH
#define FLECSOLVE_B200_TIME_INTEGRATORS_BDF_HH

#include <algorithm>
#include <array>
#include <cmath>
#include <functional>
#include <iostream>
#include <limits>
#include <vector>

#include "flecsolve/solvers/solver_settings.hh"
#include "flecsolve/time-integrators/base.hh"
#include "flecsolve/time-integrators/bdf_parameters.hh"

namespace flecsolve::time_integrator::bdf {

enum workvec : std::size_t {
	rhs,
	source,
	current_sol,
	scratch,
	previous_function,
	current_function,
	scratch_function,
	old_td,
	predict,
	time_deriv,
	nwork
};

inline constexpr std::size_t work_size = workvec::nwork + 6; // + history of up to six solutions (BDF6)

template<std::size_t Version = 0>
using topo_work = topo_work_base<work_size, Version>;

static inline work_factory<work_size> make_work;

template<class O, class W, class S>
struct integrator : base<parameters<O, W, S>> {
	using P = parameters<O, W, S>;
	using base<P>::params;
	using base<P>::current_dt;
	using base<P>::old_dt;
	using base<P>::current_time;
	using base<P>::assert_can_advance;
	using base<P>::integrator_step;

	integrator(P p)
		: base<P>(std::move(p)), prev(memory_size(params.integrator), std::ref(params.work)), solver_success{false},
		  total_steprejects{0} {
		params.validate();
		getvec<workvec::scratch>().zero();
	}

	// one implicit step of size dt from curr into out (not accepted until update())
	template<class Curr, class Out>
	void advance(double dt, bool first, Curr & curr, Out & out) {
		if (!(curr != out))
			throw std::logic_error("BDF integrator: curr cannot be the same as out");
		assert_can_advance();
		prev.seat(std::ref(params.work));

		prev[0].solution.copy(curr);
		current_dt = dt;
		set_initial_guess(first);

		auto & rhs = getvec<workvec::rhs>();
		auto & sol = getvec<workvec::current_sol>();
		rhs.scale(-1., getvec<workvec::source>());
		auto info = params.get_solver().apply(rhs, sol);
		solver_success = info.success();
		out.copy(sol);
	}

	double get_time_operator_scaling() const { return gamma; }

	bool check_solution() {
		bool passed = false;
		if (solver_success) {
			if (params.calculate_time_trunc_error)
				calculate_temporal_truncation_error();
			if (params.timestep_strategy == strategy::truncation_error) {
				passed = (time_trunc_err_est <= 1.0);
				if (passed) {
					prev_successive_rejects = (current_steprejects > 1);
					current_steprejects = 0;
				}
				else
					++current_steprejects;
			}
			else
				passed = true;
		}
		else
			++current_steprejects;
		if (!passed)
			++total_steprejects;
		return passed;
	}

	void update() {
		current_time += current_dt;
		new_time = current_time;
		if (params.use_predictor) {
			if (params.predictor == predictor::ab2)
				std::swap(getvec<workvec::old_td>(), getvec<workvec::current_function>());
			estimate_time_derivative();
		}
		prev.push(getvec<workvec::current_sol>(), current_dt);
		++integrator_step;
	}

	double get_next_dt(bool good_solution) {
		const double tmp_dt = current_dt;
		if (params.timestep_strategy == strategy::truncation_error) {
			if (good_solution) {
				current_dt = estimate_dt_with_truncation_error_estimates(current_dt, good_solution);
				if (params.max_dt < current_dt)
					current_dt = params.max_dt;
			}
			else if (solver_success)
				current_dt = estimate_dt_with_truncation_error_estimates(current_dt, good_solution);
			else
				current_dt = params.dt_cut_lower_bound * current_dt;

			evaluate_predictor();
			auto & op = params.get_operator();
			if (!op.is_valid(getvec<workvec::predict>())) {
				for (int i = 0; i < 10; ++i) { // predictor pre-check events
					current_dt = 0.5 * current_dt;
					evaluate_predictor();
					(void)op.is_valid(getvec<workvec::predict>());
				}
			}
		}
		else if (good_solution) {
			if (params.timestep_strategy == strategy::constant)
				current_dt = params.initial_dt;
			else if (params.timestep_strategy == strategy::final_constant) {
				if (ramp_interval < params.number_of_time_intervals) {
					current_dt = params.initial_dt + static_cast<double>(ramp_interval) /
					                                     static_cast<double>(params.number_of_time_intervals) *
					                                     (params.max_dt - params.initial_dt);
					++ramp_interval;
				}
				else
					current_dt = params.max_dt;
			}
			else if (params.timestep_strategy == strategy::limit_relative_change)
				current_dt = estimate_dynamical_time_scale(current_dt);
		}
		else
			current_dt = 0.;

		if (good_solution)
			old_dt = tmp_dt; // only once it has been used
		current_dt = std::min(std::min(current_dt, params.max_dt), params.final_time - current_time);
		return current_dt;
	}

	int num_step_rejects() const { return total_steprejects; }

protected:
	// CN/BE for the first step, then BDF of increasing order until enough history exists
	method get_current_integrator() const {
		if (integrator_step == 0)
			return params.starting_integrator;
		if ((integrator_step + 1) < memory_size(params.integrator))
			return static_cast<method>(integrator_step + 1);
		return params.integrator;
	}

	void set_initial_guess(bool first) {
		first_step = first;
		if (get_current_integrator() == method::cn) { // f(u_n) is part of the CN source term
			auto & scratch = getvec<workvec::scratch>();
			scratch.copy(prev[0].solution);
			params.get_operator().apply_rhs(scratch, getvec<workvec::previous_function>());
		}
		auto & sol = getvec<workvec::current_sol>();
		if (params.use_predictor) {
			auto & predict = getvec<workvec::predict>();
			if (first_step && !params.use_initial_predictor)
				predict.copy(prev[0].solution); // constant extrapolation
			else
				evaluate_predictor();
			sol.copy(predict);
		}
		else
			sol.copy(prev[0].solution);
		compute_source_term();
	}

	// source = sum_j a_j u_{n+1-j}, gamma = a_0: the step solves (I - gamma F) u_{n+1} = -source
	void compute_source_term() {
		auto & f = getvec<workvec::source>();
		const method m = get_current_integrator();
		if (m == method::be) {
			f.scale(-1., prev[0].solution);
			gamma = current_dt;
		}
		else if (m == method::cn) {
			const double eps = 0.0;
			const double a = 0.5 * current_dt + eps, b = 0.5 * current_dt - eps;
			f.axpy(b, getvec<workvec::previous_function>(), prev[0].solution); // u_n + (dt/2 - eps) f(u_n)
			f.scale(-1.);
			gamma = a;
		}
		else {
			const int k = memory_size(m);
			double h[6] = {}, n[6] = {}, a[7] = {};
			h[0] = current_dt;
			for (int i = 1; i < k; ++i)
				h[i] = h[i - 1] + prev[i - 1].dt;
			// n_j = product of the other h, taken cyclically from h_{j+1}; alpha = sum n_j
			double all = h[0];
			for (int i = 1; i < k; ++i)
				all = all * h[i];
			for (int j = 0; j < k; ++j) {
				double prod = h[(j + 1) % k];
				for (int t = 2; t < k; ++t)
					prod = prod * h[(j + t) % k];
				n[j] = prod;
			}
			if (k == 2)
				alpha = h[0] + h[1];
			else {
				alpha = n[0] + n[1];
				for (int j = 2; j < k; ++j)
					alpha = alpha + n[j];
			}
			a[0] = all / alpha;
			for (int j = 0; j < k; ++j) {
				double den = alpha;
				for (int i = 0; i < k; ++i) {
					if (i == j)
						continue;
					// the reference's a6 of BDF6 has (h1 - h5) where the pattern gives (h1 - h6); kept
					const double hj = (k == 6 && j == 5 && i == 0) ? h[4] : h[j];
					den = den * (h[i] - hj);
				}
				const double num = (k == 2) ? h[1 - j] * h[1 - j] : n[j] * n[j];
				a[j + 1] = -num / den;
			}
			f.scale(a[1], prev[0].solution);
			for (int j = 1; j < k; ++j)
				f.axpy(a[j + 1], prev[j].solution, f);
			gamma = a[0];
		}
		params.get_operator().set_scaling(gamma);
	}

	// Gresho & Sani p. 267 (CN, BE) and p. 805 (BDF2 formula, used for every BDF order)
	void estimate_time_derivative() {
		auto & td = getvec<workvec::time_deriv>();
		auto & sol = getvec<workvec::current_sol>();
		switch (get_current_integrator()) {
		case method::cn:
			td.axpy(-1., prev[0].solution, sol);
			td.linear_sum(2. / current_dt, td, -1., getvec<workvec::previous_function>());
			break;
		case method::be:
			td.axpy(-1., prev[0].solution, sol);
			td.scale(1. / current_dt);
			break;
		default: {
			const double dtt = current_dt / old_dt;
			const double c0 = (2. * dtt + 1.) / ((dtt + 1.) * current_dt);
			const double c1 = -(dtt + 1.) / current_dt;
			const double c2 = (dtt * dtt) / ((dtt + 1.) * current_dt);
			td.linear_sum(c0, sol, c1, prev[0].solution);
			td.axpy(c2, prev[1].solution, td);
		}
		}
	}

	void evaluate_predictor() {
		switch (params.integrator) {
		case method::cn:
			if (params.predictor == predictor::ab2)
				evaluate_ab2_predictor();
			else
				std::cerr << "Crack-Nicolson currently only supports ab2 predictor" << std::endl;
			break;
		case method::be:
			evaluate_forward_euler_predictor();
			break;
		default:
			if (first_step) {
				if (memory_size(params.starting_integrator) == 1)
					evaluate_forward_euler_predictor();
				else
					std::cerr << "ERROR: starting integrator must be CN or BE" << std::endl;
			}
			else if (params.predictor == predictor::leapfrog)
				evaluate_leapfrog_predictor();
			else
				std::cerr << "BDF2-6 currently only supports the leapfrog predcictor" << std::endl;
		}
		if (!params.get_operator().is_valid(getvec<workvec::predict>()))
			evaluate_forward_euler_predictor(); // constant extrapolation in time
	}

	void evaluate_ab2_predictor() {
		const double r = current_dt / old_dt;
		auto & predict = getvec<workvec::predict>();
		predict.linear_sum(current_dt * (2. + r) / 2., getvec<workvec::current_function>(), -current_dt * r / 2.,
		                   getvec<workvec::old_td>());
		predict.add(predict, prev[0].solution);
	}

	void evaluate_leapfrog_predictor() {
		const double r = current_dt / old_dt;
		const double a = r * r, b = 1. - a, c = (1. + r) * current_dt;
		auto & predict = getvec<workvec::predict>();
		predict.linear_sum(a, prev[1].solution, b, prev[0].solution);
		predict.axpy(c, getvec<workvec::time_deriv>(), predict);
	}

	void evaluate_forward_euler_predictor() {
		if (first_step) { // du/dt = f(u_0)
			auto & scratch = getvec<workvec::scratch>();
			auto & curr_func = getvec<workvec::current_function>();
			scratch.copy(prev[0].solution);
			params.get_operator().apply_rhs(scratch, curr_func);
			getvec<workvec::time_deriv>().copy(curr_func);
		}
		auto & predict = getvec<workvec::predict>();
		predict.copy(prev[0].solution);
		predict.axpy(current_dt, getvec<workvec::time_deriv>(), predict);
	}

	// step control that limits the relative change of the solution
	double estimate_dynamical_time_scale(double curr_dt) {
		if (params.integrator == method::cn)
			std::cerr << "Not implemented" << std::endl;
		auto & scratch = getvec<workvec::scratch>();
		scratch.add_scalar(prev[0].solution, std::numeric_limits<double>::epsilon());
		scratch.divide(prev[1].solution, scratch);
		scratch.add_scalar(scratch, -1);
		scratch.abs(scratch);

		std::vector<double> change;
		getvec<workvec::scratch_function>().apply([&](const auto &... v) {
			if (params.time_trunc_err_norm == vec::norm_type::inf)
				(change.push_back(v.inf_norm().get()), ...);
			else
				(change.push_back(v.l2norm().get()), ...);
		});
		const double actual = *std::max_element(change.begin(), change.end());
		double factor = (curr_dt > 0.5) ? 1.05 : 1.1;
		factor = (curr_dt > 10.) ? 1.03 : factor;
		const double cfl_new_dt = std::sqrt(params.target_relative_change / actual) * curr_dt;
		return std::min(cfl_new_dt, factor * curr_dt);
	}

	double estimate_dt_with_truncation_error_estimates(double dt, bool good_solution) {
		const double p = static_cast<double>(order(get_current_integrator()));
		const double eps = std::numeric_limits<double>::epsilon();
		const double safety = 0.8;
		const double err = std::max(time_trunc_err_est, eps);
		double factor = 0.;
		// PI control only while steps are being accepted; any rejection falls back to deadbeat
		if (params.use_pi_controller && good_solution && !prev_successive_rejects) {
			switch (params.pi_controller_type) {
			case controller::H211b: {
				const double b = 4.;
				factor = std::pow(safety / err, 1. / (b * (p + 1.)));
				factor *= std::pow(safety / std::max(prev_time_trunc_err_est, eps), 1. / (b * (p + 1.)));
				factor *= std::pow(alpha, -0.25);
				break;
			}
			case controller::pc4_7:
				factor = std::pow(safety / err, 0.4 / (p + 1.));
				factor *= std::pow(time_err_est_ratio, 0.7 / (p + 1.)) * alpha;
				break;
			case controller::pc11:
				factor = std::pow(safety / err, 1.0 / (p + 1.));
				factor *= std::pow(time_err_est_ratio, 1. / (p + 1.)) * alpha;
				break;
			case controller::deadbeat:
				factor = std::pow(safety / err, 1. / (p + 1.));
				break;
			}
		}
		else
			factor = std::pow(safety / err, 1. / (p + 1.));

		if (params.control_timestep_variation && !params.use_pi_controller && std::fabs(1. - factor) < 0.1)
			factor = 1.;
		return dt * factor;
	}

	double calculate_LTE_scaling_factor() {
		if (first_step) {
			if (params.starting_integrator == method::be)
				return 0.5;
			if (params.starting_integrator == method::cn)
				return 1.5 * current_dt; // Trompert & Verwer; the estimator is coarse
			return 0.;
		}
		switch (params.predictor) {
		case predictor::leapfrog:
			return (1. + alpha) / (2. + 3. * alpha);
		case predictor::ab2:
			return 2. / (6. - 1. / ((alpha + 1) * (alpha + 1)));
		}
		return 0.;
	}

	// norms[c] = || (x - y) / (rtol |scale| + atol) || per component
	template<class X, class Y, std::size_t N = X::num_components>
	void calculate_scaled_LTE_norm(X & x, Y & y, std::array<double, N> & norms) {
		auto & scratch = getvec<workvec::scratch>();
		if (params.time_error_scaling == error_scaling::fixed_resolution) {
			scratch.scale(params.time_rtol, x);
			scratch.add_scalar(scratch, params.time_atol);
		}
		else {
			scratch.zero();
			std::size_t i = 0;
			vec::apply([&](const auto & sol_c, auto & scratch_c) { scratch_c.add_scalar(sol_c, params.problem_scales[i++]); },
			           x, scratch);
			scratch.scale(params.time_rtol);
		}
		auto & diff = getvec<workvec::scratch_function>();
		diff.subtract(x, y);
		diff.divide(diff, scratch);
		std::size_t i = 0;
		vec::apply(
			[&](const auto & comp) {
				norms[i++] = params.time_trunc_err_norm == vec::norm_type::inf ? comp.inf_norm().get() : comp.l2norm().get();
			},
			diff);
	}

	void calculate_temporal_truncation_error() {
		auto & sol = getvec<workvec::current_sol>();
		constexpr auto nc = std::remove_reference_t<decltype(sol)>::num_components;
		if ((integrator_step > 0) || (first_step && params.use_initial_predictor)) {
			alpha = current_dt / old_dt;
			const double error_factor = calculate_LTE_scaling_factor();
			std::array<double, nc> err_norm;
			err_norm.fill(0.);
			auto & y = (first_step && params.starting_integrator == method::cn) ? prev[0].solution
			                                                                     : getvec<workvec::predict>();
			calculate_scaled_LTE_norm(sol, y, err_norm);
			double worst = err_norm[0] * error_factor;
			for (std::size_t i = 1; i < nc; ++i)
				worst = std::max(worst, err_norm[i] * error_factor);
			prev_time_trunc_err_est = time_trunc_err_est;
			time_trunc_err_est = std::max(std::numeric_limits<double>::epsilon(), worst);
			time_err_est_ratio = prev_time_trunc_err_est / time_trunc_err_est;
		}
		else { // very first step without an initial predictor: pretend the estimate was met
			prev_time_trunc_err_est = 1.;
			time_trunc_err_est = 1.;
			time_err_est_ratio = 1.;
		}
	}

	template<workvec wvec>
	auto & getvec() {
		return std::get<wvec>(params.work);
	}

	// ring of the last `len` accepted solutions living in the tail of the work array;
	// prev[0] is the most recent one
	struct history {
		using vec_t = typename std::remove_reference_t<W>::value_type;
		struct value_type {
			vec_t & solution;
			double & dt;
		};

		history(std::size_t mlen, std::reference_wrapper<W> w) : pos(0), len(mlen), work(w) { timesteps.resize(mlen); }

		value_type operator[](std::size_t back) { return at((pos + len - 1 - (back % len)) % len); }

		template<class V>
		void push(const V & v, double dt) {
			auto slot = at(pos);
			slot.solution.copy(v);
			slot.dt = dt;
			pos = (pos + 1) % len;
		}
		void seat(std::reference_wrapper<W> w) { work = w; }

	private:
		value_type at(std::size_t i) { return {work.get()[workvec::nwork + i], timesteps[i]}; }
		std::vector<double> timesteps;
		std::size_t pos, len;
		std::reference_wrapper<W> work;
	};

	bool prev_successive_rejects = false;
	int current_steprejects = 0;
	double new_time = 0;
	double time_trunc_err_est = 0, prev_time_trunc_err_est = 0, alpha = 0, time_err_est_ratio = 0;
	bool first_step = true;
	double gamma = 0;
	int ramp_interval = 1; // final_constant strategy (a function-local static in the reference)
	history prev;
	bool solver_success;
	int total_steprejects;
};
template<class O, class W, class S>
integrator(parameters<O, W, S>) -> integrator<O, W, S>;

}
#endif
