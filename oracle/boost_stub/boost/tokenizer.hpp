// Stand-in for <boost/tokenizer.hpp> -- TEST INFRASTRUCTURE ONLY (see program_options.hpp in this directory).
// Implements what /root/reference/ACFcalculator.cpp:385-396 uses: char_separator<char>{dropped, kept} with Boost's
// default drop_empty_tokens policy -- a dropped delimiter ends a token and vanishes, a kept delimiter ends a token and
// is itself returned as a one-character token, empty tokens are never returned -- and a tokenizer that can be walked
// with a range-for.
#pragma once
#include <string>
#include <vector>

namespace boost {

template <class Char>
class char_separator {
public:
    explicit char_separator(const Char* dropped, const Char* kept = "") : dropped_(dropped), kept_(kept) {}
    bool is_dropped(Char c) const { return dropped_.find(c) != std::basic_string<Char>::npos; }
    bool is_kept(Char c) const { return kept_.find(c) != std::basic_string<Char>::npos; }

private:
    std::basic_string<Char> dropped_, kept_;
};

template <class Separator>
class tokenizer {
public:
    typedef std::vector<std::string>::const_iterator iterator;
    tokenizer(const std::string& text, const Separator& sep)
    {
        std::string cur;
        for (char c : text) {
            if (sep.is_kept(c)) {  // checked first, as Boost does
                if (!cur.empty()) tokens_.push_back(cur);
                cur.clear();
                tokens_.push_back(std::string(1, c));
            } else if (sep.is_dropped(c)) {
                if (!cur.empty()) tokens_.push_back(cur);
                cur.clear();
            } else {
                cur.push_back(c);
            }
        }
        if (!cur.empty()) tokens_.push_back(cur);
    }
    iterator begin() const { return tokens_.begin(); }
    iterator end() const { return tokens_.end(); }

private:
    std::vector<std::string> tokens_;
};

}  // namespace boost
