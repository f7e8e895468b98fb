// krylov_factory: the Krylov solver named by "type = cg | gmres | bicgstab | cg-device" in a
// configuration file, built over the caller's operator / preconditioner / diagnostic handles.
//
// Reference: flecsolve/solvers/factory.hh:28-89.  `cg-device` is the extra target of SURVEY 8(f) N1
// (solvers/cg_device.hh); the reference's fourth target, nka, is a nonlinear accelerator outside
// the scoped solve loop and is reported as unavailable rather than silently mapped to something else.
#ifndef FLECSOLVE_B200_SOLVERS_FACTORY_HH
#define FLECSOLVE_B200_SOLVERS_FACTORY_HH

#include <istream>
#include <string>

#include "flecsolve/operators/factory.hh"
#include "flecsolve/solvers/bicgstab.hh"
#include "flecsolve/solvers/cg.hh"
#include "flecsolve/solvers/cg_device.hh"
#include "flecsolve/solvers/gmres.hh"

namespace flecsolve {

enum class krylov_target { cg, gmres, bicgstab, cg_device };

template<krylov_target T>
struct krylov_registry {};

template<template<class> class Solver, class Settings, class Options, class Workgen>
struct krylov_opreg {
	using settings = Settings;
	using options = Options;
	using workgen = Workgen;
	// v: any vector of the space the solver works in (work vectors are made like it)
	template<class V, class... Args>
	static auto make(const settings & s, const V & v, Args &&... args) {
		return Solver(s, workgen{}(v))(std::forward<Args>(args)...);
	}
};

template<>
struct krylov_registry<krylov_target::cg> : krylov_opreg<cg::solver, cg::settings, cg::options, decltype(cg::make_work)> {};
template<>
struct krylov_registry<krylov_target::gmres>
	: krylov_opreg<gmres::solver, gmres::settings, gmres::options, decltype(gmres::make_work)> {};
template<>
struct krylov_registry<krylov_target::bicgstab>
	: krylov_opreg<bicgstab::solver, bicgstab::settings, bicgstab::options, decltype(bicgstab::make_work)> {};
template<>
struct krylov_registry<krylov_target::cg_device>
	: krylov_opreg<cg_device::solver, cg_device::settings, cg_device::options, decltype(cg_device::make_work)> {};

inline std::istream & operator>>(std::istream & in, krylov_target & reg) {
	std::string tok;
	in >> tok;
	if (tok == "cg")
		reg = krylov_target::cg;
	else if (tok == "gmres")
		reg = krylov_target::gmres;
	else if (tok == "bicgstab")
		reg = krylov_target::bicgstab;
	else if (tok == "cg-device")
		reg = krylov_target::cg_device;
	else
		in.setstate(std::ios_base::failbit); // includes "nka"
	return in;
}

struct krylov_factory_policy {
	using target = krylov_target;
	using targets = includes<target::cg, target::gmres, target::bicgstab, target::cg_device>;
	template<target V>
	using registry = krylov_registry<V>;
};

using krylov_factory = op::factory<krylov_factory_policy>;

}
#endif
