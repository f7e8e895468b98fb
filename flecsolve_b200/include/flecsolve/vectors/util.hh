// vec::apply (visit the component vectors of one or several vectors) and norm selection.
// Reference: flecsolve/vectors/util.hh:22-41, util.cc (norm names "inf", "l1", "l2").
#ifndef FLECSOLVE_B200_VECTORS_UTIL_HH
#define FLECSOLVE_B200_VECTORS_UTIL_HH

#include <cctype>
#include <istream>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>

namespace flecsolve::vec {

namespace detail {
template<std::size_t I, class F, class... Vecs>
constexpr void apply_component(F & f, Vecs &... vecs) {
	f(std::get<I>(vecs.data.components)...);
}
template<class F, std::size_t... I, class... Vecs>
constexpr void apply_components(F & f, std::index_sequence<I...>, Vecs &... vecs) {
	(apply_component<I>(f, vecs...), ...);
}
}

// f(component_0 of every vector), f(component_1 of every vector), ...; a single vector is its own
// only component
template<class F, class... Vecs>
constexpr void apply(F && f, Vecs &&... vecs) {
	using first = std::tuple_element_t<0, std::tuple<std::remove_reference_t<Vecs>...>>;
	static_assert((... && (first::num_components == std::remove_reference_t<Vecs>::num_components)),
	              "vec::apply: vectors must have the same number of components");
	if constexpr (first::num_components == 1)
		f(vecs...);
	else
		detail::apply_components(f, std::make_index_sequence<first::num_components>(), vecs...);
}

enum class norm_type { inf, l2, l1 };

inline std::istream & operator>>(std::istream & in, norm_type & n) {
	std::string tok;
	in >> tok;
	for (auto & ch : tok) // the reference lower-cases the token (vectors/util.cc:13-16)
		ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
	if (tok == "inf")
		n = norm_type::inf;
	else if (tok == "l1")
		n = norm_type::l1;
	else if (tok == "l2")
		n = norm_type::l2;
	else
		in.setstate(std::ios_base::failbit);
	return in;
}

}
#endif
