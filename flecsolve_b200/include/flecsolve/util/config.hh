// INI configuration: option descriptions bound to settings structs, and read_config().
//
// Reference: flecsolve/util/config.hh:24-91 builds on boost::program_options; Boost is not a
// dependency here, so `flecsolve::po` below is a small stand-in with the handful of calls the
// reference's option structs make (options_description::add_options()(name, value<T>(&field)
// ->required() / ->default_value(v) / ->notifier(f) / ->multitoken(), help), add(),
// parse_config_file, store, notify, collect_unrecognized) and the same file grammar:
//   # comment      [section] or [section.sub]      name = value
// an option is addressed as "section.name".  Values are converted with operator>> (so the
// enum spellings are the ones the reference's stream operators accept), booleans take
// true/false, yes/no, on/off, 1/0.
//
// read_config keeps the reference's fixed-point loop (config.hh:50-83): settings may add options
// once a "type" has been read (operators/factory.hh), so the file is parsed until the set of
// unrecognised keys stops changing, then once more strictly.
#ifndef FLECSOLVE_B200_UTIL_CONFIG_HH
#define FLECSOLVE_B200_UTIL_CONFIG_HH

#include <algorithm>
#include <cctype>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

namespace flecsolve {

namespace util {
template<auto V>
struct constant {
	static constexpr auto value = V;
};
template<auto... V>
struct constants {};
}

namespace po {

struct error : std::runtime_error {
	using std::runtime_error::runtime_error;
};

namespace detail {
inline std::string trim(const std::string & s) {
	std::size_t a = 0, b = s.size();
	while (a < b && std::isspace(static_cast<unsigned char>(s[a])))
		++a;
	while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1])))
		--b;
	return s.substr(a, b - a);
}

template<class T>
struct is_vector : std::false_type {};
template<class T, class A>
struct is_vector<std::vector<T, A>> : std::true_type {};

template<class T>
T convert(const std::string & name, const std::string & text) {
	if constexpr (std::is_same_v<T, std::string>)
		return text;
	else if constexpr (std::is_same_v<T, bool>) {
		std::string t = text;
		std::transform(t.begin(), t.end(), t.begin(), [](unsigned char c) { return std::tolower(c); });
		if (t == "true" || t == "yes" || t == "on" || t == "1" || t.empty())
			return true;
		if (t == "false" || t == "no" || t == "off" || t == "0")
			return false;
		throw error("the argument ('" + text + "') for option '" + name + "' is invalid");
	}
	else {
		std::istringstream in(text);
		T v{};
		in >> v;
		std::string rest;
		if (in.fail() || (in >> rest))
			throw error("the argument ('" + text + "') for option '" + name + "' is invalid");
		return v;
	}
}
}

struct value_semantic {
	virtual ~value_semantic() = default;
	virtual void parse(const std::string & name, const std::string & text) = 0; // may be called once per occurrence
	virtual void finish(const std::string & name) = 0; // defaults, store to the bound field, notifier, required check
	virtual void reset() = 0;
};

template<class T>
struct typed_value : value_semantic {
	explicit typed_value(T * field) : field_(field) {}

	typed_value * operator->() { return this; }
	typed_value * required() {
		required_ = true;
		return this;
	}
	typed_value * default_value(const T & v) {
		default_ = v;
		return this;
	}
	typed_value * default_value(const T & v, const std::string &) { return default_value(v); }
	template<class F>
	typed_value * notifier(F f) {
		notify_ = std::move(f);
		return this;
	}
	typed_value * multitoken() { return this; }

	void parse(const std::string & name, const std::string & text) override {
		if constexpr (detail::is_vector<T>::value) {
			if (!value_)
				value_.emplace();
			std::istringstream in(text);
			std::string tok;
			while (in >> tok)
				value_->push_back(detail::convert<typename T::value_type>(name, tok));
		}
		else if (!value_) // first occurrence wins, like boost::program_options::store
			value_ = detail::convert<T>(name, text);
	}
	void finish(const std::string & name) override {
		if (!value_ && default_)
			value_ = default_;
		if (!value_) {
			if (required_)
				throw error("the option '" + name + "' is required but missing");
			return;
		}
		if (field_)
			*field_ = *value_;
		if (notify_)
			notify_(*value_);
	}
	void reset() override { value_.reset(); }

private:
	T * field_;
	std::optional<T> value_, default_;
	std::function<void(const T &)> notify_;
	bool required_ = false;
};

template<class T>
typed_value<T> * value(T * field = nullptr) {
	return new typed_value<T>(field);
}

struct option_description {
	std::string name, help;
	std::shared_ptr<value_semantic> semantic;
};

struct options_description {
	struct easy_init {
		options_description * owner;
		easy_init & operator()(const char * name, value_semantic * s, const char * help = "") {
			owner->options.push_back({name, help, std::shared_ptr<value_semantic>(s)});
			return *this;
		}
	};
	easy_init add_options() { return easy_init{this}; }
	options_description & add(const options_description & other) {
		options.insert(options.end(), other.options.begin(), other.options.end());
		return *this;
	}
	const option_description * find(const std::string & name) const {
		for (const auto & o : options)
			if (o.name == name)
				return &o;
		return nullptr;
	}
	std::vector<option_description> options;
};

struct option {
	std::string string_key, value;
	bool unregistered = false;
};

struct parsed_options {
	std::vector<option> options;
	const options_description * description = nullptr;
};

inline parsed_options parse_config_file(const char * fname, const options_description & desc, bool allow_unregistered = false) {
	std::ifstream in(fname);
	if (!in)
		throw error(std::string("can not read options configuration file '") + fname + "'");
	parsed_options out;
	out.description = &desc;
	std::string line, section;
	while (std::getline(in, line)) {
		const auto hash = line.find('#');
		if (hash != std::string::npos)
			line.erase(hash);
		line = detail::trim(line);
		if (line.empty())
			continue;
		if (line.front() == '[' && line.back() == ']') {
			section = detail::trim(line.substr(1, line.size() - 2));
			continue;
		}
		const auto eq = line.find('=');
		if (eq == std::string::npos)
			throw error("the options configuration file contains an invalid line '" + line + "'");
		option o;
		o.string_key = (section.empty() ? std::string() : section + ".") + detail::trim(line.substr(0, eq));
		o.value = detail::trim(line.substr(eq + 1));
		o.unregistered = desc.find(o.string_key) == nullptr;
		if (o.unregistered && !allow_unregistered)
			throw error("unrecognised option '" + o.string_key + "'");
		out.options.push_back(std::move(o));
	}
	return out;
}

// values seen so far, keyed by option name; the descriptions keep the typed copies
struct variables_map {
	std::map<std::string, std::string> seen;
	const options_description * description = nullptr;
	std::size_t count(const std::string & k) const { return seen.count(k); }
};

inline void store(const parsed_options & parsed, variables_map & vm) {
	vm.description = parsed.description;
	for (const auto & o : parsed.options) {
		if (o.unregistered)
			continue;
		parsed.description->find(o.string_key)->semantic->parse(o.string_key, o.value);
		vm.seen.emplace(o.string_key, o.value);
	}
}

inline void notify(variables_map & vm) {
	if (vm.description)
		for (const auto & o : vm.description->options)
			o.semantic->finish(o.name);
}

enum collect_unrecognized_mode { include_positional, exclude_positional };
inline std::vector<std::string> collect_unrecognized(const std::vector<option> & options, collect_unrecognized_mode) {
	std::vector<std::string> out;
	for (const auto & o : options)
		if (o.unregistered) {
			out.push_back(o.string_key);
			out.push_back(o.value);
		}
	return out;
}

}

struct with_label {
	with_label(const char * pre) : prefix(pre) {}
	void set_prefix(const char * pre) { prefix = pre; }
	const std::string & get_prefix() const { return prefix; }

protected:
	std::string prefix;
	std::string label(const char * suf) { return {prefix + "." + suf}; }
};

template<auto V>
struct null_settings {};

template<auto V>
struct null_options : with_label {
	using settings_type = null_settings<V>;
	explicit null_options(const char * pre) : with_label(pre) {}
	auto operator()(null_settings<V> &) { return po::options_description{}; }
};

template<auto... V>
using includes = util::constants<V...>;

// Options: callable objects `options_description operator()(settings_type &)`.
template<class... Options>
auto read_config(const char * fname, Options &&... ops) {
	std::tuple<typename std::decay_t<Options>::settings_type...> settings;

	std::vector<std::string> prev_opts{"-1"};
	int depth = 0;
	const int depth_limit = 50;
	bool done = false;
	while (!done) {
		po::options_description desc;
		std::apply([&](auto &... spack) { (desc.add(ops(spack)), ...); }, settings);

		po::variables_map vm;
		po::parsed_options parsed = po::parse_config_file(fname, desc, true);
		po::store(parsed, vm);
		po::notify(vm);

		std::vector<std::string> opts = po::collect_unrecognized(parsed.options, po::exclude_positional);
		if (prev_opts == opts)
			done = true;
		if (depth++ >= depth_limit)
			done = true;

		if (done)
			po::parse_config_file(fname, desc, false); // strict pass: unknown keys are errors
		else
			prev_opts = opts;
	}

	if constexpr (sizeof...(ops) == 1)
		return std::get<0>(settings);
	else
		return settings;
}

}
#endif
