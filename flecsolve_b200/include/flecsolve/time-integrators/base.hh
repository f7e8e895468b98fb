// State every time integrator shares: current time, step size, step counter.
// Reference: flecsolve/time-integrators/base.hh:24-76, parameters.hh:32-74.
#ifndef FLECSOLVE_B200_TIME_INTEGRATORS_BASE_HH
#define FLECSOLVE_B200_TIME_INTEGRATORS_BASE_HH

#include <limits>
#include <stdexcept>
#include <utility>

#include "flecsolve/operators/handle.hh"
#include "flecsolve/util/config.hh"

namespace flecsolve::time_integrator {

struct base_settings {
	double initial_time;
	double final_time;
	int max_steps;
	double max_dt = std::numeric_limits<double>::max();
	double min_dt = std::numeric_limits<double>::min();
	double initial_dt = 0;
};

// INI keys <prefix>.initial-time / final-time / max-steps (required), max-dt, min-dt, initial-dt
// (reference time-integrators/parameters.hh:39-56)
struct base_options : with_label {
	using settings_type = base_settings;
	explicit base_options(const char * pre) : with_label(pre) {}

	po::options_description operator()(settings_type & s) {
		po::options_description desc;
		desc.add_options()
			(label("initial-time").c_str(), po::value<double>(&s.initial_time)->required(), "initial time")
			(label("final-time").c_str(), po::value<double>(&s.final_time)->required(), "final time")
			(label("max-steps").c_str(), po::value<int>(&s.max_steps)->required(), "maximum number of steps")
			(label("max-dt").c_str(), po::value<double>(&s.max_dt)->default_value(std::numeric_limits<double>::max()), "")
			(label("min-dt").c_str(), po::value<double>(&s.min_dt)->default_value(std::numeric_limits<double>::min()), "")
			(label("initial-dt").c_str(), po::value<double>(&s.initial_dt)->default_value(0), "");
		return desc;
	}
};

// settings + the right-hand-side operator + work vectors
template<class S, class O, class W>
struct parameters : S {
	template<class Work>
	parameters(const S & s, op::handle<O> o, Work && w) : S(s), op(o), work(std::forward<Work>(w)) {}

	auto & get_operator() { return op.get(); }

	op::handle<O> op;
	std::decay_t<W> work;
};

template<class P>
struct base {
	base(P && p)
		: params(std::move(p)), current_time(params.initial_time), current_dt(params.initial_dt),
		  old_dt(params.initial_dt), integrator_step(0), max_integrator_steps(params.max_steps) {}

	double get_current_time() const { return current_time; }
	double get_final_time() const { return params.final_time; }
	int get_current_step() const { return integrator_step; }
	double get_current_dt() const { return current_dt; }
	bool fixed_dt() const { return params.use_fixed_dt; }
	bool steps_remaining() const { return integrator_step < max_integrator_steps; }

	void update() {
		current_time += current_dt;
		++integrator_step;
	}

protected:
	void assert_can_advance() {
		if (!(steps_remaining() && current_time < params.final_time))
			throw std::logic_error("Time integrator: already finished integrating");
	}

	P params;
	double current_time;
	double current_dt;
	double old_dt;
	int integrator_step;
	int max_integrator_steps;
};

template<class P>
struct implicit : base<P> {
	using base<P>::params;
	implicit(P && p) : base<P>(std::move(p)) {}
	auto & get_solver() { return params.solver; }
	const auto & get_solver() const { return params.solver; }
};

}
#endif
