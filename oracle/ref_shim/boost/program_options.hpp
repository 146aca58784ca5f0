// TEST INFRASTRUCTURE. A minimal stand-in for boost::program_options (Boost is not installed here), covering exactly what
// the reference's main.cpp:54-101 uses, so that the UNMODIFIED main.cpp can be compiled into oracle/_ref/coalref_main* and
// its output writer (main.cpp:232-240, 288-340) byte-compared with the host driver's (tests/test_host_driver_input.py).
// Supported: options_description(caption[, width]), add_options()(name[,short], value<T>(&v)->default_value(..)/required(), text),
// add(), parse_command_line ("--name value", "--name=value", "-x value"), store, notify, variables_map::count, operator<<.
#pragma once
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost { namespace program_options {

struct value_base {
    bool is_required = false, has_default = false;
    virtual ~value_base() {}
    virtual void parse(const std::string& s) = 0;
    virtual void apply_default() = 0;
};
template <class T> struct typed_value : value_base {
    T* target; T def;
    explicit typed_value(T* t) : target(t), def() {}
    typed_value* default_value(const T& v) { def = v; has_default = true; return this; }
    typed_value* default_value(const T& v, const std::string&) { return default_value(v); }
    typed_value* required() { is_required = true; return this; }
    void parse(const std::string& s) override { std::istringstream is(s); T v; if (!(is >> v)) throw std::runtime_error("the argument ('" + s + "') for an option is invalid"); *target = v; }
    void apply_default() override { *target = def; }
};
template <> inline void typed_value<std::string>::parse(const std::string& s) { *target = s; }
template <class T> typed_value<T>* value(T* t) { return new typed_value<T>(t); }

struct option { std::string long_name, short_name, text; std::shared_ptr<value_base> val; };

class options_description;
class options_adder {
    options_description* owner;
  public:
    explicit options_adder(options_description* o) : owner(o) {}
    options_adder& operator()(const char* name, const char* text);
    options_adder& operator()(const char* name, value_base* v, const char* text);
};
class options_description {
  public:
    std::string caption; std::vector<option> opts;
    explicit options_description(const std::string& c = "", unsigned = 80) : caption(c) {}
    options_adder add_options() { return options_adder(this); }
    options_description& add(const options_description& o) { for (const option& x : o.opts) opts.push_back(x); return *this; }
    const option* find(const std::string& key, bool is_short) const {
        for (const option& o : opts) if ((is_short ? o.short_name : o.long_name) == key) return &o;
        return nullptr;
    }
};
inline void split_name(const char* name, option& o) {
    std::string n(name); const size_t c = n.find(',');
    o.long_name = n.substr(0, c); if (c != std::string::npos) o.short_name = n.substr(c + 1);
}
inline options_adder& options_adder::operator()(const char* name, const char* text) { option o; split_name(name, o); o.text = text; owner->opts.push_back(o); return *this; }
inline options_adder& options_adder::operator()(const char* name, value_base* v, const char* text) { option o; split_name(name, o); o.text = text; o.val.reset(v); owner->opts.push_back(o); return *this; }
inline std::ostream& operator<<(std::ostream& os, const options_description& d) {
    os << d.caption << "\n";
    for (const option& o : d.opts) os << "  --" << o.long_name << (o.short_name.empty() ? "" : " [ -" + o.short_name + " ]") << "  " << o.text << "\n";
    return os;
}

struct parsed_options { const options_description* desc; std::vector<std::pair<const option*, std::string>> found; };
inline parsed_options parse_command_line(int argc, char* const argv[], const options_description& desc) {
    parsed_options p; p.desc = &desc;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i], v; bool has_v = false;
        const option* o = nullptr;
        if (a.rfind("--", 0) == 0) {
            const size_t eq = a.find('=');
            if (eq != std::string::npos) { v = a.substr(eq + 1); has_v = true; a = a.substr(0, eq); }
            o = desc.find(a.substr(2), false);
        } else if (a.size() == 2 && a[0] == '-') o = desc.find(a.substr(1), true);
        if (!o) throw std::runtime_error("unrecognised option '" + a + "'");
        if (o->val && !has_v) { if (i + 1 >= argc) throw std::runtime_error("the required argument for option '--" + o->long_name + "' is missing"); v = argv[++i]; }
        p.found.push_back(std::make_pair(o, v));
    }
    return p;
}
class variables_map {
  public:
    std::map<std::string, std::string> seen; const options_description* desc = nullptr;
    size_t count(const std::string& k) const { return seen.count(k); }
};
inline void store(const parsed_options& p, variables_map& vm) {
    vm.desc = p.desc;
    for (auto& f : p.found) { vm.seen[f.first->long_name] = f.second; }
}
inline void notify(variables_map& vm) {
    for (const option& o : vm.desc->opts) {
        if (!o.val) continue;
        auto it = vm.seen.find(o.long_name);
        if (it != vm.seen.end()) o.val->parse(it->second);
        else if (o.val->has_default) o.val->apply_default();
        else if (o.val->is_required) throw std::runtime_error("the option '--" + o.long_name + "' is required but missing");
    }
}
}}  // namespace boost::program_options
