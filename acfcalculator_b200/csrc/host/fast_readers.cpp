#include "fast_readers.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <thread>

namespace acfhost {

namespace {

struct Piece {
    const char* begin;
    const char* end;
    std::vector<double> values;
    bool ok = true;
};

// strtod on a NUL-terminated copy of [p, p+len): what atof(std::string(...).c_str()) evaluates
inline double parse_prefix(const char* p, size_t len, bool* converted, size_t* used)
{
    char buf[128];
    if (len >= sizeof(buf)) len = sizeof(buf) - 1;
    memcpy(buf, p, len);
    buf[len] = '\0';
    char* endp = nullptr;
    const double v = strtod(buf, &endp);
    if (converted) *converted = endp != buf;
    if (used) *used = (size_t)(endp - buf);
    return v;
}

// first token of the general reader must fully match [-]?[0-9]*\.?[0-9]+   (ACFcalculator.cpp:384,:400)
inline bool is_numeric_token(const char* p, size_t len)
{
    size_t i = 0;
    if (i < len && p[i] == '-') ++i;
    size_t d1 = 0;
    while (i < len && p[i] >= '0' && p[i] <= '9') {
        ++i;
        ++d1;
    }
    if (i == len) return d1 > 0;  // digits only
    if (p[i] != '.') return false;
    ++i;
    size_t d2 = 0;
    while (i < len && p[i] >= '0' && p[i] <= '9') {
        ++i;
        ++d2;
    }
    return i == len && d2 > 0;
}

bool parse_line(const char* l, size_t len, FastFormat fmt, int field, double p_factor, Piece& pc)
{
    if (len == 0) return false;  // the reference dies on a blank line: let the faithful reader do that
    switch (fmt) {
        case FAST_NAMD: {
            if (l[0] == '#') return true;
            if (field < 1) return false;
            const size_t begin = (size_t)(23 * (field - 1) + 15);
            if (begin > len) return false;  // substr would throw -> the reference stops reading here
            size_t w = len - begin;
            if (w > 21) w = 21;
            pc.values.push_back(parse_prefix(l + begin, w, nullptr, nullptr));
            return true;
        }
        case FAST_CHARMM: {
            if (l[0] == '#') return true;
            pc.values.push_back(parse_prefix(l, len, nullptr, nullptr));
            return true;
        }
        case FAST_GROMACS: {
            if (l[0] == '#' || l[0] == '@') return true;
            if (field < 1 || len >= 127) return false;
            size_t pos = 0;
            double v = 0.0;
            for (int i = 1; i <= field; ++i) {
                bool conv = false;
                size_t used = 0;
                v = parse_prefix(l + pos, len - pos, &conv, &used);
                if (!conv) return false;  // fewer columns than `field`: stale-value quirk, faithful reader
                pos += used;
            }
            pc.values.push_back(v * 10.0);
            return true;
        }
        case FAST_GENERAL: {
            // tokens: blanks separate and vanish, every tab is a token of its own (ACFcalculator.cpp:385)
            size_t i = 0;
            int index = 0;
            bool first_checked = false;
            while (i < len) {
                if (l[i] == ' ') {
                    ++i;
                    continue;
                }
                size_t j = i;
                if (l[i] == '\t')
                    j = i + 1;
                else
                    while (j < len && l[j] != ' ' && l[j] != '\t') ++j;
                if (!first_checked) {
                    if (!is_numeric_token(l + i, j - i)) return true;  // header / comment line: skipped
                    first_checked = true;
                }
                if (index == field) {
                    pc.values.push_back(parse_prefix(l + i, j - i, nullptr, nullptr) * p_factor);
                    return true;
                }
                ++index;
                i = j;
            }
            return false;  // no token at all (all blanks) or field beyond the last token: undefined upstream
        }
    }
    return false;
}

void parse_piece(Piece* pc, FastFormat fmt, int field, double p_factor)
{
    const char* p = pc->begin;
    pc->values.reserve((size_t)(pc->end - pc->begin) / 24 + 16);
    while (p < pc->end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(pc->end - p));
        const char* e = nl ? nl : pc->end;
        if (!parse_line(p, (size_t)(e - p), fmt, field, p_factor, *pc)) {
            pc->ok = false;
            return;
        }
        p = nl ? nl + 1 : pc->end;
    }
}

}  // namespace

bool fast_read_series(const char* fname, FastFormat fmt, int field, double p_factor, std::vector<double>& out)
{
    const int fd = open(fname, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) {
        close(fd);
        return false;
    }
    const size_t size = (size_t)st.st_size;
    void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return false;
    const char* base = (const char*)map;
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    if (hw > 32) hw = 32;
    size_t nth = size / (1u << 20) + 1;  // at least 1 MB per thread
    if (nth > hw) nth = hw;
    std::vector<Piece> pieces(nth);
    const char* cur = base;
    for (size_t t = 0; t < nth; ++t) {
        const char* stop = (t + 1 == nth) ? base + size : base + size * (t + 1) / nth;
        if (stop < cur) stop = cur;
        if (t + 1 < nth) {  // move to the start of the next line
            const char* nl = (const char*)memchr(stop, '\n', (size_t)(base + size - stop));
            stop = nl ? nl + 1 : base + size;
        }
        pieces[t].begin = cur;
        pieces[t].end = stop;
        cur = stop;
    }
    std::vector<std::thread> threads;
    for (size_t t = 1; t < nth; ++t) threads.emplace_back(parse_piece, &pieces[t], fmt, field, p_factor);
    parse_piece(&pieces[0], fmt, field, p_factor);
    for (auto& th : threads) th.join();
    munmap(map, size);
    size_t total = 0;
    for (const auto& pc : pieces) {
        if (!pc.ok) return false;
        total += pc.values.size();
    }
    out.clear();
    out.reserve(total);
    for (const auto& pc : pieces) out.insert(out.end(), pc.values.begin(), pc.values.end());
    return true;
}

}  // namespace acfhost
