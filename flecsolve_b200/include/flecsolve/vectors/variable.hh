// Compile-time variable tags: which physical variable a vector (or a component of a vec::multi) stands
// for, and which variables an operator reads and writes.  Vocabulary of the reference
// (flecsolve/vectors/variable.hh:26-77): variable<V>, multivariable<Vs...>, anon_var::anonymous for
// vectors that carry no name; tags compare equal iff their values do.
#ifndef FLECSOLVE_B200_VECTORS_VARIABLE_HH
#define FLECSOLVE_B200_VECTORS_VARIABLE_HH

#include <cstddef>
#include <limits>
#include <type_traits>

namespace flecsolve {

// the tag of vectors and operators that do not name their variable
enum class anon_var : std::size_t { anonymous = std::numeric_limits<std::size_t>::max() };

// optional printable name of a tag value; specialise for your own enumerators
template<auto Tag>
struct variable_name {
	static constexpr const char * value = "";
};

// ---- one variable
template<auto Tag>
struct variable_t {
	static constexpr auto value = Tag;
	static constexpr const char * name = variable_name<Tag>::value;

	template<auto Other>
	constexpr bool operator==(variable_t<Other>) const {
		return Tag == Other;
	}
	template<auto Other>
	constexpr bool operator!=(variable_t<Other>) const {
		return !(Tag == Other);
	}
};
template<auto Tag>
inline variable_t<Tag> variable{};

template<class T>
inline constexpr bool is_variable_v = false;
template<auto Tag>
inline constexpr bool is_variable_v<variable_t<Tag>> = true;
template<class T>
struct is_variable : std::bool_constant<is_variable_v<T>> {};

// ---- an ordered set of variables (operators on vec::multi)
template<auto... Tags>
struct multivariable_t {
	template<auto... Others>
	constexpr bool operator==(multivariable_t<Others...>) const {
		return ((Tags == Others) && ...);
	}
	template<auto... Others>
	constexpr bool operator!=(multivariable_t<Others...>) const {
		return ((Tags != Others) || ...);
	}
};
template<auto... Tags>
inline multivariable_t<Tags...> multivariable{};

}
#endif
