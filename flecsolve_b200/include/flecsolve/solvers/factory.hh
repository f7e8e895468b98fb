// krylov_factory: the Krylov solver a configuration file names,
//     [linear-solver]            type = cg | gmres | bicgstab | cg-device | cg-sr
//     [linear-solver.options]    maxiter = ...   (the chosen solver's own option table)
// built over the caller's operator / preconditioner / diagnostic handles:
//     auto s   = read_config(file, krylov_factory::options("linear-solver"));
//     auto slv = krylov_factory::make_shared(s, u, A, P, diagnostic);     // u: any vector of the solve space
//
// Same surface as the reference (flecsolve/solvers/factory.hh:28-89: krylov_target, krylov_registry<T>,
// krylov_factory_policy, krylov_factory = op::factory<policy>).  `cg-device` is the extra target of
// SURVEY 8(f) N1 (solvers/cg_device.hh), `cg-sr` its single-reduction sibling (solvers/cg_sr.hh).  The reference's fourth target, nka, is a nonlinear accelerator
// outside the scoped solve loop: its name is rejected as an invalid value rather than mapped to something else.
#ifndef FLECSOLVE_B200_SOLVERS_FACTORY_HH
#define FLECSOLVE_B200_SOLVERS_FACTORY_HH

#include <array>
#include <istream>
#include <ostream>
#include <string>
#include <string_view>
#include <utility>

#include "flecsolve/operators/factory.hh"
#include "flecsolve/solvers/bicgstab.hh"
#include "flecsolve/solvers/cg.hh"
#include "flecsolve/solvers/cg_device.hh"
#include "flecsolve/solvers/cg_sr.hh"
#include "flecsolve/solvers/gmres.hh"

namespace flecsolve {

enum class krylov_target { cg, gmres, bicgstab, cg_device, cg_sr };

namespace detail {
// spelling of each target in configuration files, indexed by the enum value
inline constexpr std::array<std::string_view, 5> krylov_target_names{"cg", "gmres", "bicgstab", "cg-device", "cg-sr"};

// what the operator factory needs to know about one solver family: its settings and option types and
// how to bind it; WorkFactory is the family's `make_work` object type
template<template<class> class Bound, class Settings, class Options, class WorkFactory>
struct krylov_family {
	using settings = Settings;
	using options = Options;
	using workgen = WorkFactory;

	// like_this: a vector of the space the solver works in; the work vectors are made to match it
	template<class Vec, class... Handles>
	static auto make(const Settings & with, const Vec & like_this, Handles &&... handles) {
		auto unbound = Bound(with, WorkFactory{}(like_this));
		return std::move(unbound)(std::forward<Handles>(handles)...);
	}
};
}

template<krylov_target>
struct krylov_registry; // one specialisation per target, below

#define FLECSOLVE_KRYLOV_FAMILY(TARGET, NS)                                                                    \
	template<>                                                                                                 \
	struct krylov_registry<krylov_target::TARGET>                                                              \
		: detail::krylov_family<NS::solver, NS::settings, NS::options, std::decay_t<decltype(NS::make_work)>> {}
FLECSOLVE_KRYLOV_FAMILY(cg, cg);
FLECSOLVE_KRYLOV_FAMILY(gmres, gmres);
FLECSOLVE_KRYLOV_FAMILY(bicgstab, bicgstab);
FLECSOLVE_KRYLOV_FAMILY(cg_device, cg_device);
FLECSOLVE_KRYLOV_FAMILY(cg_sr, cg_sr);
#undef FLECSOLVE_KRYLOV_FAMILY

inline std::ostream & operator<<(std::ostream & os, krylov_target t) {
	return os << detail::krylov_target_names[static_cast<std::size_t>(t)];
}

inline std::istream & operator>>(std::istream & is, krylov_target & t) {
	std::string word;
	is >> word;
	for (std::size_t k = 0; k < detail::krylov_target_names.size(); ++k)
		if (word == detail::krylov_target_names[k]) {
			t = static_cast<krylov_target>(k);
			return is;
		}
	is.setstate(std::ios_base::failbit); // "nka" ends here too
	return is;
}

struct krylov_factory_policy {
	using target = krylov_target;
	using targets = includes<target::cg, target::gmres, target::bicgstab, target::cg_device, target::cg_sr>;
	template<target T>
	using registry = krylov_registry<T>;
};

using krylov_factory = op::factory<krylov_factory_policy>;

}
#endif
