// vec::multi<Vecs...>: a tuple of component vectors that behaves like one vector.
// Reference: flecsolve/vectors/multi.hh:30-193 (component access by index or by variable tag,
// subset(variable / multivariable) returning a component or a narrower multi).
#ifndef FLECSOLVE_B200_VECTORS_MULTI_HH
#define FLECSOLVE_B200_VECTORS_MULTI_HH

#include <tuple>
#include <type_traits>

#include "flecsolve/vectors/core.hh"
#include "flecsolve/vectors/operations/multi.hh"
#include "flecsolve/vectors/traits.hh"
#include "flecsolve/vectors/variable.hh"

namespace flecsolve::vec {

template<class... Vecs>
struct multi_config {
	using first = std::remove_reference_t<std::tuple_element_t<0, std::tuple<Vecs...>>>;
	using scalar = typename first::scalar; // scalar type of the first component
	using len_t = typename first::len_t;
	using real = typename num_traits<scalar>::real;
	using var_t = typename first::var_t;
	static constexpr auto var = multivariable<std::remove_reference_t<Vecs>::var.value...>;
	using storage_type = std::tuple<Vecs...>;
	static constexpr std::size_t num_components = sizeof...(Vecs);
};

namespace data {
template<class Config>
struct multi {
	using config = Config;
	template<class... V>
	multi(V &&... v) : components(std::forward<V>(v)...) {}
	typename Config::storage_type components;
};
}

template<class... Vecs>
struct multi : core<data::multi, ops::multi, multi_config<Vecs...>> {
	using base = core<data::multi, ops::multi, multi_config<Vecs...>>;
	using base::data;
	using base::var;
	using var_t = typename base::var_t;
	using data_t = typename base::data_t;

	template<class Head, class... Tail,
	         std::enable_if_t<(... && std::is_same_v<typename std::remove_reference_t<Head>::var_t,
	                                                  typename std::remove_reference_t<Tail>::var_t>),
	                          bool> = true>
	multi(Head && head, Tail &&... tail) : base{data_t{std::forward<Head>(head), std::forward<Tail>(tail)...}} {}

	template<std::size_t I>
	constexpr auto & get() & { return std::get<I>(data.components); }
	template<std::size_t I>
	constexpr const auto & get() const & { return std::get<I>(data.components); }

	// first component whose variable tag equals `v`
	template<var_t v>
	constexpr decltype(auto) getvar() const { return std::get<index_of<v>()>(data.components); }
	template<var_t v>
	constexpr decltype(auto) getvar() { return std::get<index_of<v>()>(data.components); }

	template<var_t v>
	constexpr decltype(auto) subset(variable_t<v>) const { return getvar<v>(); }
	template<var_t v>
	constexpr decltype(auto) subset(variable_t<v>) { return getvar<v>(); }
	template<var_t... vs>
	constexpr decltype(auto) subset(multivariable_t<vs...>) const {
		if constexpr (sizeof...(vs) == 1)
			return (getvar<vs>(), ...);
		else
			return multi<decltype(getvar<vs>())...>(getvar<vs>()...);
	}
	template<var_t... vs>
	constexpr decltype(auto) subset(multivariable_t<vs...>) {
		if constexpr (sizeof...(vs) == 1)
			return (getvar<vs>(), ...);
		else
			return multi<decltype(getvar<vs>())...>(getvar<vs>()...);
	}

private:
	template<var_t v>
	static constexpr std::size_t index_of() {
		constexpr bool match[] = {(std::remove_reference_t<Vecs>::var == variable<v>)...};
		for (std::size_t i = 0; i < sizeof...(Vecs); ++i)
			if (match[i])
				return i;
		return sizeof...(Vecs); // out of range: std::get fails to compile, as intended
	}
};
template<class H, class... T>
multi(H &&, T &&...) -> multi<H, T...>;

template<class... A, class... B>
bool operator==(const multi<A...> & a, const multi<B...> & b) {
	return a.data.components == b.data.components;
}
template<class... A, class... B>
bool operator!=(const multi<A...> & a, const multi<B...> & b) {
	return a.data.components != b.data.components;
}

template<class... Vecs, std::enable_if_t<(... && is_vector_v<std::decay_t<Vecs>>), bool> = true>
auto make(Vecs &&... vecs) {
	return multi<Vecs...>(std::forward<Vecs>(vecs)...);
}

}

namespace std {
template<class... Vecs>
struct tuple_size<flecsolve::vec::multi<Vecs...>> : integral_constant<size_t, sizeof...(Vecs)> {};
template<size_t I, class... Vecs>
struct tuple_element<I, flecsolve::vec::multi<Vecs...>> {
	using type = tuple_element_t<I, tuple<Vecs...>>;
};
}
#endif
