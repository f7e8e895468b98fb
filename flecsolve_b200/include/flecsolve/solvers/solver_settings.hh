// Solver settings, results and work-vector factories.
//
// Reference: flecsolve/solvers/solver_settings.hh:31-200.  Field names and types are kept
// (note rtol/atol and the norms in solve_info are float).  `solver_options` binds them to the
// INI keys <prefix>.maxiter / rtol / atol / use-zero-guess (reference :37-57) through util/config.hh.
#ifndef FLECSOLVE_B200_SOLVERS_SOLVER_SETTINGS_HH
#define FLECSOLVE_B200_SOLVERS_SOLVER_SETTINGS_HH

#include <array>
#include <ostream>
#include <tuple>
#include <type_traits>

#include "flecsolve/util/config.hh"
#include "flecsolve/vectors/multi.hh"
#include "flecsolve/vectors/topo_view.hh"

namespace flecsolve {

struct solver_settings {
	int maxiter;
	float rtol;
	float atol;
	bool use_zero_guess;
};

struct solver_options : with_label {
	using settings_type = solver_settings;
	solver_options(const char * pre) : with_label(pre) {}

	po::options_description operator()(settings_type & settings) {
		po::options_description desc;
		desc.add_options()
			(label("maxiter").c_str(), po::value<int>(&settings.maxiter)->required(), "maximum number of iterations")
			(label("rtol").c_str(), po::value<float>(&settings.rtol)->default_value(0), "relative tolerance")
			(label("atol").c_str(), po::value<float>(&settings.atol)->default_value(0), "absolute tolerance")
			(label("use-zero-guess").c_str(), po::value<bool>(&settings.use_zero_guess)->required(), "use zero initial guess");
		return desc;
	}
};

struct solve_stats {};

struct solve_info {
	enum class stop_reason {
		converged_atol,
		converged_rtol,
		converged_user,
		diverged_dtol,
		diverged_iters,
		diverged_breakdown,
		unknown
	};
	stop_reason status = stop_reason::unknown;
	int iters = 0;
	int restarts = 0;
	// zero until a solver fills them in (the reference leaves them indeterminate; not every solver sets every field)
	float res_norm_initial = 0, res_norm_final = 0;
	float sol_norm_initial = 0, sol_norm_final = 0;
	float rhs_norm = 0;

	bool success() const {
		return status == stop_reason::converged_atol || status == stop_reason::converged_rtol ||
		       status == stop_reason::converged_user;
	}
};

inline std::ostream & operator<<(std::ostream & os, const solve_info::stop_reason & r) {
	static constexpr const char * text[] = {"Converged to absolute tolerance",
	                                        "Converged to residual tolerance",
	                                        "Converged to user tolerance",
	                                        "Diverged: reached divergence tolerance",
	                                        "Diverged: exceeded maximum iterations",
	                                        "Diverged due to breakdown",
	                                        "Status unknown"};
	return os << text[static_cast<int>(r)];
}

template<std::size_t V>
struct version_t {
	static constexpr std::size_t value = V;
};
template<std::size_t V>
inline version_t<V> version;

// NumWork static field definitions per (vector type, count, version, component index): the work
// vectors of a solver are fields like any other, defined once and reused by every solve.
template<class Vec, std::size_t NumWork, std::size_t Version, std::size_t MVIndex = 0>
struct topo_solver_state {
	using field_def = typename Vec::data_t::field_definition;
	using topo_t = typename Vec::data_t::topo_t;
	static inline const std::array<field_def, NumWork> defs = {};

	static std::array<Vec, NumWork> get_work(const Vec & rhs) {
		return build(rhs.data.topo(), std::make_index_sequence<NumWork>());
	}

private:
	template<std::size_t... I>
	static std::array<Vec, NumWork> build(typename topo_t::topology & slot, std::index_sequence<I...>) {
		if constexpr (std::is_same_v<typename Vec::var_t, anon_var>)
			return {vec::make(defs[I](slot))...};
		else
			return {vec::make(Vec::var, defs[I](slot))...};
	}
};

template<std::size_t NumWork, std::size_t Version>
struct topo_work_base {
	template<class Vec>
	static auto get(const Vec & rhs) {
		return topo_solver_state<Vec, NumWork, Version>::get_work(rhs);
	}

	// multi-vector: one state per component, transposed into NumWork multi-vectors
	template<class... Vecs>
	static auto get(const vec::multi<Vecs...> & rhs) {
		auto per_component = states(rhs.data.components, std::index_sequence_for<Vecs...>());
		return transpose(per_component, std::make_index_sequence<NumWork>());
	}

private:
	template<class Tuple, std::size_t... C>
	static auto states(const Tuple & comps, std::index_sequence<C...>) {
		return std::make_tuple(
			topo_solver_state<std::remove_cv_t<std::remove_reference_t<std::tuple_element_t<C, Tuple>>>, NumWork, Version,
		                      C>::get_work(std::get<C>(comps))...);
	}
	template<std::size_t W, class States>
	static auto column(States & st) {
		return std::apply([](auto &... arr) { return vec::multi(std::move(arr[W])...); }, st);
	}
	template<class States, std::size_t... W>
	static auto transpose(States & st, std::index_sequence<W...>) {
		return std::array{column<W>(st)...};
	}
};

template<std::size_t nwork>
struct work_factory {
	template<class Vec>
	constexpr auto operator()(Vec & b) const {
		return topo_work_base<nwork, 0>::get(b);
	}
	template<class Vec, std::size_t Ver>
	constexpr auto operator()(Vec & b, version_t<Ver>) const {
		return topo_work_base<nwork, Ver>::get(b);
	}
};

}
#endif
