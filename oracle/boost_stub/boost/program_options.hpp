// Stand-in for <boost/program_options.hpp> -- TEST INFRASTRUCTURE ONLY.
//
// Boost is not installed in this image, and /root/reference/traj2diffusion.cpp needs it for nothing but
// its command line (traj2diffusion.cpp:394-478).  This header implements, from Boost's documented
// behaviour, exactly the handful of calls that main() makes, so that oracle/Makefile can compile the
// UNMODIFIED reference source into oracle/_ref/traj2diffusion_ref and use it as the end-to-end oracle
// of the MSD program (readers, imaging, MSD loop, slope, printf formats are all the reference's own
// object code).  What this stub does NOT pin: the wording of option-parsing errors and the layout of
// --help, which come from the real Boost in a real build.
//
// Semantics kept: long/short names ("traj,t"), "--opt value", "--opt=value", "-o value", "-ovalue";
// required(); default_value() (a defaulted option has count()==1, as in Boost); store-to pointers are
// written by notify(); unknown options, missing values and bad values throw std::exception subclasses.
#pragma once
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost {
namespace program_options {

struct error : std::logic_error {
    explicit error(const std::string& w) : std::logic_error(w) {}
};

struct value_semantic {
    bool is_required = false;
    bool has_default = false;
    virtual ~value_semantic() {}
    virtual void parse(const std::string& opt, const std::string& text) = 0;  // text -> held value
    virtual void apply_default() = 0;
    virtual void publish() = 0;  // copy held value to the store-to pointer
    virtual std::string default_text() const = 0;
};

template <class T>
struct typed_value : value_semantic {
    T held{};
    T dflt{};
    T* store_to;
    explicit typed_value(T* p) : store_to(p) {}
    typed_value* required()
    {
        is_required = true;
        return this;
    }
    typed_value* default_value(const T& v)
    {
        dflt = v;
        has_default = true;
        return this;
    }
    void parse(const std::string& opt, const std::string& text) override { parse_impl(opt, text, held); }
    void apply_default() override { held = dflt; }
    void publish() override
    {
        if (store_to) *store_to = held;
    }
    std::string default_text() const override
    {
        std::ostringstream os;
        os << dflt;
        return os.str();
    }

private:
    static void parse_impl(const std::string&, const std::string& text, std::string& out) { out = text; }
    template <class U>
    static void parse_impl(const std::string& opt, const std::string& text, U& out)
    {
        std::istringstream is(text);
        U v;
        char extra;
        if (!(is >> v) || (is >> extra))
            throw error("the argument ('" + text + "') for option '--" + opt + "' is invalid");
        out = v;
    }
};

template <class T>
typed_value<T>* value()
{
    return new typed_value<T>(nullptr);
}
template <class T>
typed_value<T>* value(T* p)
{
    return new typed_value<T>(p);
}

struct option_description {
    std::string long_name, short_name, text;
    std::shared_ptr<value_semantic> sem;  // null for a flag such as --help
};

class options_description;
class options_description_easy_init {
public:
    explicit options_description_easy_init(options_description* o) : owner(o) {}
    options_description_easy_init& operator()(const char* name, const char* text);
    options_description_easy_init& operator()(const char* name, value_semantic* s, const char* text);

private:
    options_description* owner;
};

class options_description {
public:
    options_description(const std::string& caption, unsigned = 80, unsigned = 40) : caption_(caption) {}
    options_description_easy_init add_options() { return options_description_easy_init(this); }
    const option_description* find(const std::string& name, bool is_short) const
    {
        for (const auto& o : opts_)
            if ((is_short ? o.short_name : o.long_name) == name) return &o;
        return nullptr;
    }
    std::vector<option_description> opts_;
    std::string caption_;
};

inline options_description_easy_init& options_description_easy_init::operator()(const char* name, const char* text)
{
    return (*this)(name, nullptr, text);
}
inline options_description_easy_init& options_description_easy_init::operator()(const char* name, value_semantic* s,
                                                                                 const char* text)
{
    option_description d;
    std::string n(name);
    const size_t comma = n.find(',');
    d.long_name = n.substr(0, comma);
    if (comma != std::string::npos) d.short_name = n.substr(comma + 1);
    d.text = text;
    d.sem.reset(s);
    owner->opts_.push_back(d);
    return *this;
}

// Layout rule of Boost's options_description printer: first column = longest "  name arg" + 1 (at least 24).
inline std::ostream& operator<<(std::ostream& os, const options_description& d)
{
    std::vector<std::string> left;
    size_t width = 23;
    for (const auto& o : d.opts_) {
        std::string l = "  ";
        if (!o.short_name.empty()) l += "-" + o.short_name + " [ --" + o.long_name + " ]";
        else l += "--" + o.long_name;
        l += ' ';
        if (o.sem) {
            l += "arg";
            if (o.sem->has_default) l += " (=" + o.sem->default_text() + ")";
        }
        if (l.size() > width) width = l.size();
        left.push_back(l);
    }
    ++width;
    os << d.caption_ << ":\n";
    for (size_t i = 0; i < d.opts_.size(); ++i) {
        os << left[i];
        if (!d.opts_[i].text.empty()) {
            for (size_t pad = left[i].size(); pad < width; ++pad) os.put(' ');
            os << d.opts_[i].text;
        }
        os << "\n";
    }
    return os;
}

struct variable_value {
    std::shared_ptr<value_semantic> sem;
    bool defaulted = false;
    template <class T>
    const T& as() const
    {
        auto* tv = dynamic_cast<typed_value<T>*>(sem.get());
        if (!tv) throw error("boost::bad_any_cast: failed conversion using boost::any_cast");
        return tv->held;
    }
};

class variables_map : public std::map<std::string, variable_value> {
public:
    size_t count(const std::string& k) const { return std::map<std::string, variable_value>::count(k); }
    const variable_value& operator[](const std::string& k) const
    {
        static const variable_value empty;
        auto it = find(k);
        return it == end() ? empty : it->second;
    }
    std::vector<std::string> required_missing;  // filled by store(), reported by notify()
};

struct parsed_options {
    const options_description* desc;
    std::vector<std::pair<const option_description*, std::string>> items;
};

inline parsed_options parse_command_line(int argc, char** argv, const options_description& desc)
{
    parsed_options out;
    out.desc = &desc;
    for (int i = 1; i < argc; ++i) {
        const std::string tok = argv[i];
        const option_description* od = nullptr;
        std::string val;
        bool have_val = false;
        if (tok.size() > 2 && tok[0] == '-' && tok[1] == '-') {
            const size_t eq = tok.find('=');
            const std::string name = tok.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            od = desc.find(name, false);
            if (!od) throw error("unrecognised option '--" + name + "'");
            if (eq != std::string::npos) {
                val = tok.substr(eq + 1);
                have_val = true;
            }
        } else if (tok.size() >= 2 && tok[0] == '-') {
            od = desc.find(tok.substr(1, 1), true);
            if (!od) throw error("unrecognised option '" + tok + "'");
            if (tok.size() > 2) {
                val = tok.substr(2);
                have_val = true;
            }
        } else {
            throw error("too many positional options have been specified on the command line");
        }
        if (od->sem && !have_val) {
            if (i + 1 >= argc) throw error("the required argument for option '--" + od->long_name + "' is missing");
            val = argv[++i];
        }
        out.items.push_back(std::make_pair(od, val));
    }
    return out;
}

inline void store(const parsed_options& po, variables_map& vm)
{
    for (const auto& it : po.items) {
        const option_description* od = it.first;
        if (vm.count(od->long_name) && !vm[od->long_name].defaulted)
            throw error("option '--" + od->long_name + "' cannot be specified more than once");
        variable_value v;
        v.sem = od->sem;
        if (od->sem) od->sem->parse(od->long_name, it.second);
        static_cast<std::map<std::string, variable_value>&>(vm)[od->long_name] = v;
    }
    for (const auto& od : po.desc->opts_) {
        if (!od.sem || vm.count(od.long_name)) continue;
        if (od.sem->has_default) {
            variable_value v;
            v.sem = od.sem;
            v.defaulted = true;
            od.sem->apply_default();
            static_cast<std::map<std::string, variable_value>&>(vm)[od.long_name] = v;
        } else if (od.sem->is_required) {
            vm.required_missing.push_back(od.long_name);
        }
    }
}

inline void notify(variables_map& vm)
{
    if (!vm.required_missing.empty())
        throw error("the option '--" + vm.required_missing.front() + "' is required but missing");
    for (auto& kv : vm)
        if (kv.second.sem) kv.second.sem->publish();
}

}  // namespace program_options
}  // namespace boost
