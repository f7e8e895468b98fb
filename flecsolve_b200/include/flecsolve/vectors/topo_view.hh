// vec::topo_view: a vector that views one field of a topology; vec::make(...) factories.
// Reference: flecsolve/vectors/topo_view.hh:26-82 (same factory overload set, minus the
// FleCSI layout parameter which has no device meaning).
#ifndef FLECSOLVE_B200_VECTORS_TOPO_VIEW_HH
#define FLECSOLVE_B200_VECTORS_TOPO_VIEW_HH

#include <tuple>

#include "flecsolve/vectors/core.hh"
#include "flecsolve/vectors/data/topo_view.hh"
#include "flecsolve/vectors/operations/topo_view.hh"
#include "flecsolve/vectors/variable.hh"

namespace flecsolve::vec {

template<auto V, class Scalar, class Topo, typename Topo::index_space Space>
struct topo_view_config {
	using scalar = Scalar;
	using real = typename num_traits<scalar>::real;
	using len_t = std::size_t;
	static constexpr auto var = variable<V>;
	using var_t = decltype(V);
	static constexpr std::size_t num_components = 1;
	using topo_t = Topo;
	static constexpr typename Topo::index_space space = Space;
};

template<auto V, class Scalar, class Topo, typename Topo::index_space Space>
using topo_view = core<data::topo_view, ops::topo_view, topo_view_config<V, Scalar, Topo, Space>>;

template<auto V, class T, class Topo, typename Topo::index_space Space>
auto make(variable_t<V>, ::flecsolve::data::field_reference<T, Topo, Space> ref) {
	using vec_t = topo_view<V, T, Topo, Space>;
	return vec_t{data::topo_view<typename vec_t::config>{ref}};
}

template<class T, class Topo, typename Topo::index_space Space>
auto make(::flecsolve::data::field_reference<T, Topo, Space> ref) {
	return make(variable<anon_var::anonymous>, ref);
}

namespace detail {
template<class T, class = void>
struct is_topology : std::false_type {};
template<class T>
struct is_topology<T, std::void_t<decltype(std::declval<T &>().store), decltype(std::declval<T &>().mat)>> : std::true_type {};
}

// vec::make(var, topo)(defs...) -> one vector, or a tuple of vectors
template<auto V, class Topology, std::enable_if_t<detail::is_topology<Topology>::value, bool> = true>
auto make(variable_t<V> var, Topology & topo) {
	return [&topo, var](auto &... fd) {
		if constexpr (sizeof...(fd) == 1)
			return (make(var, fd(topo)), ...);
		else
			return std::tuple(make(var, fd(topo))...);
	};
}

template<class Topology, std::enable_if_t<detail::is_topology<Topology>::value, bool> = true>
auto make(Topology & topo) {
	return make(variable<anon_var::anonymous>, topo);
}

}
#endif
