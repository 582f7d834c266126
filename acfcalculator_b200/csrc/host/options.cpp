#include "options.hpp"

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <sstream>

namespace acfhost {

// Exact name first; otherwise a unique prefix among the documented options (boost's allow_guessing).  Options that only
// this program knows (hidden aliases) must be spelled out, so that the reference's own abbreviations keep their meaning.
// Several candidates: boost's ambiguous_option error, candidates in alphabetical order.
const OptionSpec* Options::find_long(const std::string& name, bool allow_prefix, const std::string& token) const
{
    std::vector<const OptionSpec*> hits;
    for (const auto& s : specs_) {
        if (name == s.long_name) return &s;
        if (allow_prefix && !s.hidden && !name.empty() && std::strncmp(s.long_name, name.c_str(), name.size()) == 0)
            hits.push_back(&s);
    }
    if (hits.size() == 1) return hits[0];
    if (hits.size() > 1) {
        std::vector<std::string> names;
        for (const auto* h : hits) names.push_back(std::string("'--") + h->long_name + "'");
        std::sort(names.begin(), names.end());
        std::string msg = "option '" + token + "' is ambiguous and matches ";
        for (size_t i = 0; i < names.size(); ++i) {
            if (i + 1 == names.size()) msg += "and ";
            msg += names[i];
            if (i + 1 < names.size()) msg += ", ";
        }
        throw OptionError(msg);
    }
    return nullptr;
}

const OptionSpec* Options::find_short(char c) const
{
    for (const auto& s : specs_)
        if (s.short_name != '\0' && s.short_name == c) return &s;
    return nullptr;
}

// lexical_cast-like conversions: the whole token must be consumed, no surrounding blanks
static bool to_double(const std::string& t, double* out)
{
    if (t.empty() || t[0] == ' ' || t[0] == '\t') return false;
    errno = 0;
    char* end = nullptr;
    const double v = std::strtod(t.c_str(), &end);
    if (end == t.c_str() || *end != '\0') return false;
    if (t.find_first_of("xXpP") != std::string::npos) return false;  // no hex floats in lexical_cast
    // a finite literal that overflows is refused ("1e400"); "inf", "-infinity", "nan" are values; underflow is accepted
    const size_t first = (t[0] == '-' || t[0] == '+') ? 1 : 0;
    if (std::isinf(v) && !(first < t.size() && (t[first] == 'i' || t[first] == 'I'))) return false;
    *out = v;
    return true;
}

static bool to_int(const std::string& t, int* out)
{
    if (t.empty()) return false;
    size_t i = (t[0] == '-' || t[0] == '+') ? 1 : 0;
    if (i == t.size()) return false;
    for (size_t j = i; j < t.size(); ++j)
        if (t[j] < '0' || t[j] > '9') return false;
    errno = 0;
    const long v = std::strtol(t.c_str(), nullptr, 10);
    if (errno != 0 || v > 2147483647L || v < -2147483647L - 1) return false;
    *out = (int)v;
    return true;
}

void Options::parse(int argc, const char* const* argv)
{
    struct Seen {
        const OptionSpec* spec;
        std::string value;
    };
    std::vector<Seen> seen;
    bool only_positional = false;
    for (int i = 1; i < argc; ++i) {
        const std::string tok = argv[i];
        if (only_positional || tok.size() < 2 || tok[0] != '-') continue;  // positionals are ignored
        if (tok == "--") {
            only_positional = true;
            continue;
        }
        const OptionSpec* spec = nullptr;
        std::string value;
        bool have_value = false;
        if (tok[1] == '-') {
            const size_t eq = tok.find('=');
            const std::string name = tok.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            if (eq != std::string::npos && eq + 1 == tok.size())  // checked while the token is split, before any look-up
                throw OptionError("the argument for option '--" + name + "' should follow immediately after the equal sign");
            spec = find_long(name, true, tok);
            if (!spec) throw OptionError("unrecognised option '" + tok + "'");
            if (eq != std::string::npos) {
                value = tok.substr(eq + 1);
                have_value = true;
                if (spec->kind == OptionSpec::Flag)
                    throw OptionError(std::string("option '--") + spec->long_name + "' does not take any arguments");
            }
        } else {
            // sticky short options: a flag may be followed by more short options in the same token ("-habc" is -h and
            // -a with the value "bc"); if the tail does not start with a known option the whole token is unknown
            size_t pos = 1;
            spec = find_short(tok[pos]);
            while (spec && spec->kind == OptionSpec::Flag && pos + 1 < tok.size()) {
                seen.push_back({spec, ""});
                spec = find_short(tok[++pos]);
            }
            if (!spec) throw OptionError("unrecognised option '" + tok + "'");
            if (spec->kind != OptionSpec::Flag && pos + 1 < tok.size()) {
                value = tok.substr(pos + 1);
                have_value = true;
            }
        }
        if (spec->kind == OptionSpec::Flag) {
            seen.push_back({spec, ""});
            continue;
        }
        if (!have_value) {
            // the next token is the value -- unless it is, literally, the short form of a known option
            const bool next_is_short_option =
                i + 1 < argc && std::strlen(argv[i + 1]) == 2 && argv[i + 1][0] == '-' && find_short(argv[i + 1][1]);
            if (i + 1 >= argc || next_is_short_option)
                throw OptionError(std::string("the required argument for option '--") + spec->long_name + "' is missing");
            // boost looks at the candidate value with its option parsers first; "--name=" trips the long-option one
            const std::string cand = argv[i + 1];
            if (cand.size() > 3 && cand[0] == '-' && cand[1] == '-' && cand.find('=') == cand.size() - 1)
                throw OptionError(std::string("the argument for option '--") + spec->long_name +
                                  "' should follow immediately after the equal sign");
            value = argv[++i];
        }
        seen.push_back({spec, value});
    }
    // store phase
    for (const auto& s : seen) {
        const std::string key = s.spec->long_name;
        if (values_.count(key)) throw OptionError("option '--" + key + "' cannot be specified more than once");
        double d;
        int n;
        if ((s.spec->kind == OptionSpec::Double && !to_double(s.value, &d)) ||
            (s.spec->kind == OptionSpec::Int && !to_int(s.value, &n)))
            throw OptionError(s.value.empty() ? "the argument for option '--" + key + "' is invalid"
                                              : "the argument ('" + s.value + "') for option '--" + key + "' is invalid");
        values_[key] = s.value;
    }
}

void Options::notify() const
{
    std::vector<std::string> req;
    for (const auto& s : specs_)
        if (s.required) req.push_back(s.long_name);
    std::sort(req.begin(), req.end());
    for (const auto& r : req)
        if (!values_.count(r)) throw OptionError("the option '--" + r + "' is required but missing");
}

double Options::get_double(const std::string& n, double dflt) const
{
    auto it = values_.find(n);
    double v = dflt;
    if (it != values_.end()) to_double(it->second, &v);
    return v;
}

int Options::get_int(const std::string& n, int dflt) const
{
    auto it = values_.find(n);
    int v = dflt;
    if (it != values_.end()) to_int(it->second, &v);
    return v;
}

std::string Options::help_text() const
{
    // boost::program_options layout: two spaces, "-x [ --name ] arg (=dflt)", padded to the widest entry + 1
    std::vector<std::string> left;
    size_t width = 0;
    for (const auto& s : specs_) {
        if (s.hidden) {
            left.push_back("");
            continue;
        }
        std::string l = std::string("  -") + s.short_name + " [ --" + s.long_name + " ]";
        if (s.kind != OptionSpec::Flag) {
            l += " arg";
            if (s.default_text) l += std::string(" (=") + s.default_text + ")";
        }
        left.push_back(l);
        if (l.size() > width) width = l.size();
    }
    std::ostringstream os;
    os << "Allowed options:\n";
    for (size_t i = 0; i < specs_.size(); ++i) {
        if (specs_[i].hidden) continue;
        os << left[i] << std::string(width + 1 - left[i].size(), ' ') << specs_[i].description << "\n";
    }
    return os.str();
}

}  // namespace acfhost
