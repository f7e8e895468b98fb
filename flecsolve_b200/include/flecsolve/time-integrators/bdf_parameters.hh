// Settings of the variable-step BDF / Crank-Nicolson integrator.
// Reference: flecsolve/time-integrators/bdf_parameters.hh:29-171, bdf_parameters.cc (enum
// spellings), bdf.cc (memory_size, order).  Defaults are those of the reference's option table.
#ifndef FLECSOLVE_B200_TIME_INTEGRATORS_BDF_PARAMETERS_HH
#define FLECSOLVE_B200_TIME_INTEGRATORS_BDF_PARAMETERS_HH

#include <cctype>
#include <istream>
#include <stdexcept>
#include <string>
#include <vector>

#include "flecsolve/time-integrators/base.hh"
#include "flecsolve/vectors/util.hh"

namespace flecsolve::time_integrator::bdf {

enum class method { cn, be, bdf2, bdf3, bdf4, bdf5, bdf6 };
enum class predictor { ab2, leapfrog };
enum class strategy { truncation_error, constant, final_constant, limit_relative_change };
enum class controller { H211b, pc4_7, pc11, deadbeat };
enum class error_scaling { fixed_resolution, fixed_scaling };

// number of previous solutions a method needs (CN and BE: one)
inline int memory_size(method m) {
	const auto v = static_cast<int>(m);
	return v == 0 ? 1 : v;
}
inline short order(method m) {
	return m == method::cn ? 2 : (m == method::be ? 1 : static_cast<short>(m));
}

namespace detail {
// lower: the reference lower-cases predictor / strategy tokens before comparing (bdf_parameters.cc:9-46)
template<class E>
std::istream & parse(std::istream & in, E & out, std::initializer_list<std::pair<const char *, E>> table, bool lower = false) {
	std::string tok;
	in >> tok;
	if (lower)
		for (auto & ch : tok)
			ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
	for (const auto & [name, value] : table)
		if (tok == name) {
			out = value;
			return in;
		}
	in.setstate(std::ios_base::failbit);
	return in;
}
}
inline std::istream & operator>>(std::istream & in, predictor & p) {
	return detail::parse(in, p, {{"ab2", predictor::ab2}, {"leapfrog", predictor::leapfrog}}, true);
}
inline std::istream & operator>>(std::istream & in, strategy & s) {
	return detail::parse(in, s,
	                     {{"truncation-error", strategy::truncation_error},
	                      {"constant", strategy::constant},
	                      {"final-constant", strategy::final_constant},
	                      {"limit-relative-change", strategy::limit_relative_change}},
	                     true);
}
inline std::istream & operator>>(std::istream & in, method & m) {
	return detail::parse(in, m,
	                     {{"BE", method::be},
	                      {"CN", method::cn},
	                      {"BDF2", method::bdf2},
	                      {"BDF3", method::bdf3},
	                      {"BDF4", method::bdf4},
	                      {"BDF5", method::bdf5},
	                      {"BDF6", method::bdf6}});
}
inline std::istream & operator>>(std::istream & in, controller & c) {
	return detail::parse(in, c,
	                     {{"PC.4.7", controller::pc4_7},
	                      {"H211b", controller::H211b},
	                      {"PC11", controller::pc11},
	                      {"Deadbeat", controller::deadbeat}});
}
inline std::istream & operator>>(std::istream & in, error_scaling & e) {
	return detail::parse(in, e,
	                     {{"fixed-scaling", error_scaling::fixed_scaling},
	                      {"fixed-resolution", error_scaling::fixed_resolution}});
}

struct settings : base_settings {
	bool use_predictor = true;
	bool use_initial_predictor = true;
	bool has_source_term = false;
	bool combine_timestep_estimators = false;
	bdf::predictor predictor = bdf::predictor::leapfrog;
	bdf::strategy timestep_strategy = bdf::strategy::truncation_error;
	double dt_cut_lower_bound = 0.58754407; // Emmrich 2008
	double dt_growth_upper_bound = 1.702;
	int number_of_time_intervals = 100;
	bdf::method integrator = bdf::method::bdf2;
	bdf::method starting_integrator = bdf::method::cn;
	bool calculate_time_trunc_error = false;
	vec::norm_type time_trunc_err_norm = vec::norm_type::inf;
	double target_relative_change = 0;
	bool use_pi_controller = true;
	bdf::controller pi_controller_type = bdf::controller::pc4_7;
	bool control_timestep_variation = false;
	bdf::error_scaling time_error_scaling = bdf::error_scaling::fixed_scaling;
	double time_rtol = 1e-9, time_atol = 1e-15;
	std::vector<double> problem_scales;

	// consistency rules of the reference (bdf_parameters.hh:72-114)
	void validate() {
		auto require = [](bool ok, const char * msg) {
			if (!ok)
				throw std::invalid_argument(msg);
		};
		require(memory_size(starting_integrator) == 1, "Starting integrator must be CN or BE");
		if (use_predictor)
			calculate_time_trunc_error = true;
		if (timestep_strategy == strategy::truncation_error) {
			calculate_time_trunc_error = true;
			use_predictor = true;
		}
		else {
			combine_timestep_estimators = false;
			control_timestep_variation = false;
			use_pi_controller = false;
		}
		if (calculate_time_trunc_error) {
			use_predictor = true;
			if (time_error_scaling == error_scaling::fixed_scaling)
				require(!problem_scales.empty(), "Problem scales must be specified if using fixed time error scaling");
		}
		if (use_predictor) {
			if (memory_size(integrator) > 1)
				require(predictor == bdf::predictor::leapfrog, "Only the leapfrog predictor is supported for BDF2-6");
			else if (integrator == method::cn)
				require(predictor == bdf::predictor::ab2, "Valid option for Crank-Nicolson predictor is only ab2 currently");
		}
	}
};

// INI keys of the reference's option table (bdf_parameters.hh:116-151), after base_options'
struct options : base_options {
	using settings_type = settings;
	explicit options(const char * pre) : base_options(pre) {}

	po::options_description operator()(settings_type & s) {
		auto desc = base_options::operator()(s);
		desc.add_options()
			(label("use-predictor").c_str(), po::value<bool>(&s.use_predictor)->default_value(true), "use a predictor")
			(label("has-source-term").c_str(), po::value<bool>(&s.has_source_term)->default_value(false), "has source term")
			(label("use-initial-predictor").c_str(), po::value<bool>(&s.use_initial_predictor)->default_value(true),
			 "use an initial predictor")
			(label("predictor").c_str(), po::value<bdf::predictor>(&s.predictor)->required(), "ab2 or leapfrog")
			(label("timestep-selection-strategy").c_str(), po::value<strategy>(&s.timestep_strategy)->required(),
			 "truncation-error, constant, final-constant or limit-relative-change")
			(label("dt-cut-lower-bound").c_str(), po::value<double>(&s.dt_cut_lower_bound)->default_value(0.58754407), "")
			(label("dt-growth-upper-bound").c_str(), po::value<double>(&s.dt_growth_upper_bound)->default_value(1.702), "")
			(label("number-of-time-intervals").c_str(), po::value<int>(&s.number_of_time_intervals)->default_value(100),
			 "final-constant strategy")
			(label("integrator").c_str(), po::value<bdf::method>(&s.integrator)->required(), "BE, CN or BDF2..BDF6")
			(label("starting-integrator").c_str(),
			 po::value<bdf::method>(&s.starting_integrator)->required()->notifier([](const bdf::method & m) {
				 if (memory_size(m) != 1)
					 throw std::invalid_argument("BE or CN must be used for starting integrator");
			 }),
			 "one step integrator (BE or CN)")
			(label("calculate-time-trunc-error").c_str(), po::value<bool>(&s.calculate_time_trunc_error), "")
			(label("time-trunc-error-norm").c_str(), po::value<vec::norm_type>(&s.time_trunc_err_norm), "inf, l1 or l2")
			(label("target-relative-change").c_str(), po::value<double>(&s.target_relative_change), "")
			(label("use-pi-controller").c_str(), po::value<bool>(&s.use_pi_controller)->default_value(true), "")
			(label("pi-controller-type").c_str(),
			 po::value<bdf::controller>(&s.pi_controller_type)->default_value(bdf::controller::pc4_7, "PC.4.7"),
			 "H211b, PC.4.7, PC11 or Deadbeat")
			(label("control-timestep-variation").c_str(), po::value<bool>(&s.control_timestep_variation)->default_value(false), "")
			(label("time-error-scaling").c_str(),
			 po::value<bdf::error_scaling>(&s.time_error_scaling)->default_value(error_scaling::fixed_scaling, "fixed-scaling"),
			 "fixed-scaling or fixed-resolution")
			(label("trunc-error-rtol").c_str(), po::value<double>(&s.time_rtol)->default_value(1e-9), "")
			(label("trunc-error-atol").c_str(), po::value<double>(&s.time_atol)->default_value(1e-15), "")
			(label("combine-timestep-estimators").c_str(), po::value<bool>(&s.combine_timestep_estimators)->default_value(false), "")
			(label("problem-scales").c_str(), po::value<std::vector<double>>(&s.problem_scales)->multitoken(), "");
		return desc;
	}
};

template<class Op, class Work, class Solver>
struct parameters : time_integrator::parameters<settings, Op, Work> {
	using base = time_integrator::parameters<settings, Op, Work>;

	template<class O, class W, class S>
	parameters(const settings & s, op::handle<O> o, W && w, op::handle<S> slv)
		: base(s, o, std::forward<W>(w)), solver(slv) {}

	auto & get_solver() { return solver.get(); }

protected:
	op::handle<Solver> solver;
};
template<class O, class W, class S>
parameters(const settings &, op::handle<O>, W &&, op::handle<S>) -> parameters<O, W, S>;

}
#endif
