// STUB (test infrastructure): declarations of the Boost.program_options names the reference's
// option structs mention.  Nothing parses anything: refcheck fills the settings structs directly.
#pragma once
#include <string>
#include <vector>

namespace boost::program_options {

struct value_semantic {};
template<class T>
struct typed_value : value_semantic {
	typed_value * required() { return this; }
	typed_value * default_value(const T &) { return this; }
	template<class U>
	typed_value * default_value(const U &) { return this; }
	template<class U>
	typed_value * default_value(const U &, const char *) { return this; }
	template<class F>
	typed_value * notifier(F &&) { return this; }
	typed_value * multitoken() { return this; }
};
template<class T>
typed_value<T> * value(T *) {
	static typed_value<T> v;
	return &v;
}
template<class T>
typed_value<T> * value() {
	static typed_value<T> v;
	return &v;
}

struct options_description;
struct options_description_easy_init {
	options_description_easy_init & operator()(const char *, const value_semantic *, const char *) { return *this; }
	options_description_easy_init & operator()(const char *, const char *) { return *this; }
};
struct options_description {
	options_description() = default;
	explicit options_description(const std::string &) {}
	options_description_easy_init add_options() { return {}; }
	options_description & add(const options_description &) { return *this; }
};

struct option {};
struct parsed_options {
	std::vector<option> options;
};
struct variables_map {};
enum collect_unrecognized_mode { include_positional, exclude_positional };
parsed_options parse_config_file(const char *, const options_description &, bool);
void store(const parsed_options &, variables_map &);
void notify(variables_map &);
std::vector<std::string> collect_unrecognized(const std::vector<option> &, collect_unrecognized_mode);

}
