// Small type traits shared by the header layer.  The names are the ones flecsolve user code and the other
// headers rely on (reference: flecsolve/util/traits.hh:25-103); the implementations are this repo's.
#ifndef FLECSOLVE_B200_UTIL_TRAITS_HH
#define FLECSOLVE_B200_UTIL_TRAITS_HH

#include <complex>
#include <functional>
#include <memory>
#include <tuple>
#include <type_traits>
#include <utility>

namespace flecsolve {

namespace detail {
// is T some instance Template<...> (type parameters only)?
template<class T, template<class...> class Template>
struct instance_of : std::false_type {};
template<template<class...> class Template, class... Args>
struct instance_of<Template<Args...>, Template> : std::true_type {};

// can `Struct{args...}` be written?
template<class Always, class Struct, class... Args>
struct brace_constructible : std::false_type {};
template<class Struct, class... Args>
struct brace_constructible<std::void_t<decltype(Struct{std::declval<Args>()...})>, Struct, Args...> : std::true_type {};

// a single argument that is the struct itself (copy / move), which aggregate initialisation must not swallow
template<class Struct, class... Args>
inline constexpr bool is_self_v = false;
template<class Struct, class Arg>
inline constexpr bool is_self_v<Struct, Arg> = std::is_same_v<std::decay_t<Arg>, Struct>;
}

// handles: how operators and solvers hold one another (operators/handle.hh)
template<class T>
using is_reference_wrapper = detail::instance_of<T, std::reference_wrapper>;
template<class T>
inline constexpr bool is_reference_wrapper_v = is_reference_wrapper<T>::value;

template<class T>
using is_smart_ptr = std::disjunction<detail::instance_of<T, std::unique_ptr>, detail::instance_of<T, std::shared_ptr>>;
template<class T>
inline constexpr bool is_smart_ptr_v = is_smart_ptr<T>::value;

// customisation point for user types (left empty here, specialised by vector / operator headers)
template<class T>
struct traits {};

// scalar <-> real pairing.  The device back end computes in real fp64 only (the reference itself disables
// complex vectors under CUDA, vectors/test/flecsi_vector.cc:398-400); the complex case is kept for host code.
template<class Scalar>
struct num_traits {
	static constexpr bool is_complex = false;
	using scalar = Scalar;
	using real = Scalar;
};
template<class Real>
struct num_traits<std::complex<Real>> {
	static constexpr bool is_complex = true;
	using scalar = std::complex<Real>;
	using real = Real;
};

// CRTP access to the most derived type
template<class Derived>
struct with_derived {
	constexpr const Derived & derived() const { return *static_cast<const Derived *>(this); }
	constexpr Derived & derived() { return *static_cast<Derived *>(this); }
};

template<class Struct, class... Args>
using is_direct_list_initializable = detail::brace_constructible<void, Struct, Args...>;
template<class Struct, class... Args>
inline constexpr bool is_direct_list_initializable_v = is_direct_list_initializable<Struct, Args...>::value;

// an aggregate that can be brace-initialised from Args..., the copy case excluded
// (op::base uses it to accept `params{a, b, c}` for plain settings structs)
template<class Struct, class... Args>
using is_aggregate_initializable =
	std::bool_constant<std::is_aggregate_v<Struct> && is_direct_list_initializable_v<Struct, Args...> &&
                       !detail::is_self_v<Struct, Args...>>;
template<class Struct, class... Args>
inline constexpr bool is_aggregate_initializable_v = is_aggregate_initializable<Struct, Args...>::value;

}
#endif
