// Run-time choice between operator types that share an interface (the solver named by
// "<prefix>.type" in a configuration file).
//
// Reference: flecsolve/operators/factory.hh:27-226.  A policy P names an enum of targets and a
// registry<target> per target {settings, options, make(settings, args...)}; factory<P> then offers
//   settings  = {optional<target>, variant<settings of every target>}
//   options   : reads "<prefix>.type", and once it is known the chosen target's own options under
//               "<prefix>.options.*" (read_config's fixed-point loop supplies the second pass)
//   make / make_shared(settings, args...) -> op::core over a variant of the concrete operators
#ifndef FLECSOLVE_B200_OPERATORS_FACTORY_HH
#define FLECSOLVE_B200_OPERATORS_FACTORY_HH

#include <optional>
#include <stdexcept>
#include <variant>

#include "flecsolve/operators/core.hh"
#include "flecsolve/operators/handle.hh"
#include "flecsolve/util/config.hh"

namespace flecsolve::op {

// the operator the factory hands out: forwards to whichever alternative was built
template<class Variant>
struct factory_prod {
	static constexpr auto input_var = std::variant_alternative_t<0, Variant>::input_var;
	static constexpr auto output_var = std::variant_alternative_t<0, Variant>::output_var;
	using params_t = std::nullptr_t;

	template<class D, class R>
	decltype(auto) apply(const D & x, R & y) const {
		return std::visit([&](auto && p) { return p.apply(x, y); }, var);
	}
	auto & get_operator() {
		return std::visit([](auto && p) -> auto & { return p.get_operator(); }, var);
	}
	const auto & get_operator() const {
		return std::visit([](auto && p) -> const auto & { return p.get_operator(); }, var);
	}

	Variant var;
};

template<class P>
struct factory {
	using target = typename P::target;
	using targets = typename P::targets;

private:
	template<class T>
	struct unpack;
	template<auto... V>
	struct unpack<includes<V...>> {
		using settings_types = std::variant<typename P::template registry<V>::settings...>;

		static settings_types default_settings(target r) {
			settings_types s;
			((r == V ? void(s = typename P::template registry<V>::settings{}) : void()), ...);
			return s;
		}
		static po::options_description describe(const std::string & pre, target r, settings_types & s) {
			po::options_description desc;
			(
				[&] {
					using reg = typename P::template registry<V>;
					if (r == V)
						if (auto * mine = std::get_if<typename reg::settings>(&s))
							desc.add(typename reg::options(pre.c_str())(*mine));
				}(),
				...);
			return desc;
		}
		template<class... Args>
		static auto build(target r, const settings_types & s, Args &&... args) {
			using var_t = std::variant<decltype(P::template registry<V>::make(
				std::declval<const typename P::template registry<V>::settings &>(), std::forward<Args>(args)...))...>;
			std::optional<var_t> out;
			(
				[&] {
					using reg = typename P::template registry<V>;
					if (r == V)
						out.emplace(reg::make(std::get<typename reg::settings>(s), std::forward<Args>(args)...));
				}(),
				...);
			if (!out)
				throw std::invalid_argument("operator factory: unknown target");
			return std::move(out).value();
		}
	};
	using impl = unpack<targets>;

public:
	using settings_types = typename impl::settings_types;

	struct settings {
		std::optional<target> target_id;
		settings_types target_settings;
	};

	struct options : with_label {
		using settings_type = settings;
		explicit options(const char * pre) : with_label(pre) {}

		po::options_description operator()(settings_type & s) {
			po::options_description desc;
			desc.add_options()(label("type").c_str(),
			                   po::value<target>()->required()->notifier([&s](const target & t) { s.target_id.emplace(t); }),
			                   "operator type");
			if (s.target_id.has_value()) {
				s.target_settings = impl::default_settings(*s.target_id);
				desc.add(impl::describe(label("options"), *s.target_id, s.target_settings));
			}
			return desc;
		}
	};

	static settings_types make_settings(target r) { return impl::default_settings(r); }

	template<class... Args>
	static auto make_policy(const settings & s, Args &&... args) {
		if (!s.target_id)
			throw std::invalid_argument("operator factory: no type was configured");
		return impl::build(*s.target_id, s.target_settings, std::forward<Args>(args)...);
	}
	template<class... Args>
	static auto make(const settings & s, Args &&... args) {
		auto var = make_policy(s, std::forward<Args>(args)...);
		return op::core<factory_prod<decltype(var)>>(factory_prod<decltype(var)>{std::move(var)});
	}
	template<class... Args>
	static auto make_shared(const settings & s, Args &&... args) {
		auto var = make_policy(s, std::forward<Args>(args)...);
		return ::flecsolve::op::make_shared<factory_prod<decltype(var)>>(factory_prod<decltype(var)>{std::move(var)});
	}
};

}
#endif
