// Command-line parsing for the ACFcalculator front-end, without Boost.
//
// Reproduces the observable behaviour of the boost::program_options set-up in
// /root/reference/ACFcalculator.cpp:554-586 (option names, defaults, help text, error strings, the order
// in which errors surface, exit status), as probed from the shipped binary:
//   parse phase  : unknown option, missing argument            -> what() on stderr, exit 1
//   store phase  : repeated option, value that does not convert -> what() on stderr, exit 1
//   --help       : usage on stdout, exit 1 (checked after store, before the required-option check)
//   notify phase : first missing required option, alphabetical by long name (acf, input, output)
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace acfhost {

struct OptionSpec {
    const char* long_name;
    char short_name;
    enum Kind { Flag, String, Double, Int } kind;
    bool required;
    const char* default_text;  // as printed in the help ("(=0)"), nullptr if none
    const char* description;
    bool hidden;               // accepted but not listed in --help (aliases that only the docs mention)
};

class OptionError : public std::runtime_error {
public:
    explicit OptionError(const std::string& m) : std::runtime_error(m) {}
};

class Options {
public:
    explicit Options(const std::vector<OptionSpec>& specs) : specs_(specs) {}
    // parse + store; throws OptionError with boost's wording
    void parse(int argc, const char* const* argv);
    // required-option check; throws OptionError
    void notify() const;
    bool count(const std::string& long_name) const { return values_.count(long_name) != 0; }
    std::string get_string(const std::string& n) const { return values_.at(n); }
    double get_double(const std::string& n, double dflt) const;
    int get_int(const std::string& n, int dflt) const;
    std::string help_text() const;

private:
    const OptionSpec* find_long(const std::string& name, bool allow_prefix, const std::string& token) const;
    const OptionSpec* find_short(char c) const;
    std::vector<OptionSpec> specs_;
    std::map<std::string, std::string> values_;
};

}  // namespace acfhost
