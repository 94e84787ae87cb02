// oracle/shim/boost/program_options.hpp -- TEST INFRASTRUCTURE ONLY.
//
// Stand-in for the Boost.program_options calls made by the unmodified reference src/main.cpp:11,
// 15, 35-52, 54-170 (headers are absent from this image).  Flag parsing only -- no arithmetic.
// Accepts "--name value" and "--name=value"; unknown options and bad values throw, like Boost.
#ifndef FARMS_ORACLE_SHIM_BOOST_PO
#define FARMS_ORACLE_SHIM_BOOST_PO

#include <map>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost { namespace program_options {

struct value_semantic { virtual ~value_semantic() {} };
template <class T> struct typed_value : value_semantic {};
template <class T> typed_value<T> *value() { return new typed_value<T>(); }

struct option_spec { std::string name; bool takes_value; std::string help; };

class options_description;
class options_easy_init {
public:
  explicit options_easy_init(options_description *o) : owner_(o) {}
  options_easy_init &operator()(const char *name, const char *help);
  options_easy_init &operator()(const char *name, const value_semantic *v, const char *help);
private:
  options_description *owner_;
};

class options_description {
public:
  explicit options_description(const std::string &caption) : caption_(caption) {}
  options_easy_init add_options() { return options_easy_init(this); }
  std::string caption_;
  std::vector<option_spec> opts_;
};

inline options_easy_init &options_easy_init::operator()(const char *name, const char *help) {
  owner_->opts_.push_back(option_spec{name, false, help});
  return *this;
}
inline options_easy_init &options_easy_init::operator()(const char *name, const value_semantic *v,
                                                        const char *help) {
  delete v;
  owner_->opts_.push_back(option_spec{name, true, help});
  return *this;
}

inline std::ostream &operator<<(std::ostream &os, const options_description &d) {
  os << d.caption_ << ":\n";
  for (size_t i = 0; i < d.opts_.size(); i++)
    os << "  --" << d.opts_[i].name << (d.opts_[i].takes_value ? " arg" : "") << "\t" << d.opts_[i].help << "\n";
  return os;
}

struct parsed_options { std::vector<std::pair<std::string, std::string> > kv; };

inline parsed_options parse_command_line(int argc, char *argv[], const options_description &d) {
  parsed_options p;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a.size() < 3 || a[0] != '-' || a[1] != '-')
      throw std::runtime_error("too many positional options have been specified on the command line");
    std::string name = a.substr(2), val;
    bool has_eq = false;
    size_t eq = name.find('=');
    if (eq != std::string::npos) { val = name.substr(eq + 1); name = name.substr(0, eq); has_eq = true; }
    const option_spec *spec = 0;
    for (size_t k = 0; k < d.opts_.size(); k++) if (d.opts_[k].name == name) spec = &d.opts_[k];
    if (!spec) throw std::runtime_error("unrecognised option '--" + name + "'");
    if (spec->takes_value && !has_eq) {
      if (i + 1 >= argc) throw std::runtime_error("the required argument for option '--" + name + "' is missing");
      val = argv[++i];
    }
    p.kv.push_back(std::make_pair(name, val));
  }
  return p;
}

class variable_value {
public:
  variable_value() {}
  explicit variable_value(const std::string &s) : s_(s) {}
  template <class T> T as() const {
    std::istringstream is(s_);
    T v; is >> v;
    if (is.fail()) throw std::runtime_error("the argument ('" + s_ + "') for option is invalid");
    return v;
  }
private:
  std::string s_;
};
template <> inline std::string variable_value::as<std::string>() const { return s_; }

class variables_map {
public:
  size_t count(const std::string &k) const { return m_.count(k); }
  const variable_value &operator[](const std::string &k) const { return m_.find(k)->second; }
  std::map<std::string, variable_value> m_;
};

inline void store(const parsed_options &p, variables_map &vm) {
  for (size_t i = 0; i < p.kv.size(); i++)
    if (!vm.m_.count(p.kv[i].first)) vm.m_[p.kv[i].first] = variable_value(p.kv[i].second);
}
inline void notify(variables_map &) {}

}}  // namespace boost::program_options

#endif
