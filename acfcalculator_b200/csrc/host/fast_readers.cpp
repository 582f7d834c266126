#include "fast_readers.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

namespace acfhost {

namespace {

struct Piece {
    const char* begin;
    const char* end;
    std::vector<double> values;
    bool ok = true;
};

// Decimal text -> double without strtod for the common case (Clinger's fast path): when the digit string is an
// integer m <= 2^53 and the decimal exponent satisfies |e| <= 22, both m and 10^|e| are exact doubles and the single
// IEEE multiplication or division rounds correctly -- the same bits strtod returns.  Anything else (more than 19
// digits, large exponents, hex floats, inf/nan, no digits) reports failure and the caller uses strtod.
// Consumes the longest prefix strtod would: [ws][sign]digits[.digits][(e|E)[sign]digits].
inline bool fast_decimal(const char* p, size_t len, double* value, size_t* used)
{
    static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                      1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    size_t i = 0;
    while (i < len && (p[i] == ' ' || (p[i] >= '\t' && p[i] <= '\r'))) ++i;  // isspace in the C locale
    bool neg = false;
    if (i < len && (p[i] == '-' || p[i] == '+')) neg = p[i++] == '-';
    unsigned long long m = 0;
    int digits = 0, frac = 0;
    bool any = false;
    while (i < len && p[i] >= '0' && p[i] <= '9') {
        if (m != 0 || p[i] != '0') ++digits;
        if (digits > 19) return false;
        m = m * 10 + (unsigned)(p[i] - '0');
        ++i;
        any = true;
    }
    if (i < len && p[i] == '.') {
        size_t j = i + 1;
        while (j < len && p[j] >= '0' && p[j] <= '9') {
            if (m != 0 || p[j] != '0') ++digits;
            if (digits > 19) return false;
            m = m * 10 + (unsigned)(p[j] - '0');
            ++frac;
            ++j;
            any = true;
        }
        if (any) i = j;  // "." alone is not a number
    }
    if (!any) return false;  // inf, nan, hex, empty: let strtod decide
    int e10 = 0;
    if (i < len && (p[i] == 'e' || p[i] == 'E')) {
        size_t j = i + 1;
        bool eneg = false;
        if (j < len && (p[j] == '-' || p[j] == '+')) eneg = p[j++] == '-';
        if (j < len && p[j] >= '0' && p[j] <= '9') {
            int e = 0;
            while (j < len && p[j] >= '0' && p[j] <= '9') {
                if (e < 10000) e = e * 10 + (p[j] - '0');
                ++j;
            }
            e10 = eneg ? -e : e;
            i = j;
        }
    }
    if (i < len && (p[i] == 'x' || p[i] == 'X')) return false;  // "0x..." is a hex float for strtod
    e10 -= frac;
    if (m > (1ull << 53) || e10 < -22 || e10 > 22) return false;
    double v = (double)m;
    v = e10 < 0 ? v / kPow10[-e10] : v * kPow10[e10];
    *value = neg ? -v : v;
    *used = i;
    return true;
}

// what atof(std::string(p, len).c_str()) evaluates: the fast path above, else strtod on a NUL-terminated copy
inline double parse_prefix(const char* p, size_t len, bool* converted, size_t* used)
{
    double fv;
    size_t fu;
    if (fast_decimal(p, len, &fv, &fu)) {
        if (converted) *converted = true;
        if (used) *used = fu;
        return fv;
    }
    char small[128];
    std::string big;  // a token can be hundreds of digits long ("%f" of 1e300)
    char* buf = small;
    if (len >= sizeof(small)) {
        big.assign(p, len);
        buf = &big[0];
    } else {
        memcpy(small, p, len);
        small[len] = '\0';
    }
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
            if (field < 1) return false;
            size_t pos = 0;
            double v = 0.0;
            for (int i = 1; i <= field; ++i) {
                bool conv = false;
                size_t used = 0;
                v = parse_prefix(l + pos, len - pos, &conv, &used);
                if (!conv) return false;  // fewer columns than `field`: stale-value quirk, faithful reader
                pos += used;
                // sscanf("%lf") and strtod agree on a number that ends at white space; where they can differ
                // (a dangling exponent such as "1e", "0x", "infi") the faithful reader decides
                if (pos < len && !(l[pos] == ' ' || (l[pos] >= '\t' && l[pos] <= '\r'))) return false;
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

// Maps `fname`, cuts it into per-thread pieces at line boundaries (>= 1 MB each, at most 32) and runs work(t) for every
// piece on its own thread.  false: the file could not be opened / mapped or is empty.
template <class PieceT, class Work>
bool for_each_piece(const char* fname, std::vector<PieceT>& pieces, Work work)
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
    pieces.assign(nth, PieceT());
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
    for (size_t t = 1; t < nth; ++t) threads.emplace_back(work, &pieces[t]);
    work(&pieces[0]);
    for (auto& th : threads) th.join();
    munmap(map, size);
    return true;
}

// ---- NAMD log, "PRESSURE:" lines (viscosity.cpp:117-169) --------------------------------------------------------
struct PressurePiece {
    const char* begin;
    const char* end;
    std::vector<double> values;  // 9 per accepted line
    bool ok = true;
};

inline bool is_space(char ch) { return ch == ' ' || (ch >= '\t' && ch <= '\r'); }

void parse_pressure_piece(PressurePiece* pc)
{
    const char* p = pc->begin;
    pc->values.reserve((size_t)(pc->end - pc->begin) / 40 + 16);
    while (p < pc->end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(pc->end - p));
        const char* e = nl ? nl : pc->end;
        if (e - p >= 9 && memcmp(p, "PRESSURE:", 9) == 0) {  // line.compare(0, 9, "PRESSURE:") == 0   (:138)
            const char* q = p;
            for (int tok = 0; tok < 11; ++tok) {  // operator>> : skip white space, take the run of non-blanks
                while (q < e && is_space(*q)) ++q;
                if (q == e) {  // fewer than 11 tokens: the reference converts a stale string -- faithful reader
                    pc->ok = false;
                    return;
                }
                const char* t0 = q;
                while (q < e && !is_space(*q)) ++q;
                if (tok < 2) continue;  // "PRESSURE:" and the time step   (:143-145)
                bool conv = false;
                errno = 0;
                const double v = parse_prefix(t0, (size_t)(q - t0), &conv, nullptr);  // std::stod = strtod of a prefix
                if (!conv || errno == ERANGE) {  // std::stod throws here and the reference dies: faithful reader
                    pc->ok = false;
                    return;
                }
                pc->values.push_back(v);
            }
        }
        p = nl ? nl + 1 : pc->end;
    }
}

// ---- traj2diffusion trajectories (traj2diffusion.cpp:59-121 fixed columns, :145-260 general) -------------------
struct TrajPiece {
    const char* begin;
    const char* end;
    std::vector<double> values;  // namd: x, y, z per frame; general: x per frame
    bool ok = true;
};

// isNumeric(), traj2diffusion.cpp:123-143
inline bool t2d_numeric(char ch)
{
    return (ch >= '0' && ch <= '9') || ch == '-' || ch == '+' || ch == '.' || ch == 'e' || ch == 'E';
}

void parse_traj_piece(TrajPiece* pc, bool namd)
{
    const char* p = pc->begin;
    pc->values.reserve((size_t)(pc->end - pc->begin) / (namd ? 28 : 40) + 16);
    while (p < pc->end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(pc->end - p));
        const char* e = nl ? nl : pc->end;
        const size_t len = (size_t)(e - p);
        if (namd) {
            if (len == 0 || (p[0] != '#' && len < 61)) {  // blank line: the reference throws; short line: it stops
                pc->ok = false;
                return;
            }
            if (p[0] != '#') {
                const size_t off[3] = {15, 37, 61};
                for (int a = 0; a < 3; ++a) {
                    size_t w = len - off[a];
                    if (w > 23) w = 23;
                    pc->values.push_back(parse_prefix(p + off[a], w, nullptr, nullptr));
                }
            }
        } else {
            const char first = len ? p[0] : '\0';
            if (t2d_numeric(first) || first == ' ') {
                size_t start = 0, end;
                while (start < len && !t2d_numeric(p[start])) ++start;  // time-step column
                end = start;
                while (end < len && t2d_numeric(p[end])) ++end;
                start = end;
                while (start < len && !t2d_numeric(p[start])) ++start;  // x (y and z are copies of it upstream)
                end = start;
                while (end < len && t2d_numeric(p[end])) ++end;
                pc->values.push_back(parse_prefix(p + start, end - start, nullptr, nullptr));
            }
        }
        p = nl ? nl + 1 : pc->end;
    }
}

}  // namespace

bool fast_read_series(const char* fname, FastFormat fmt, int field, double p_factor, std::vector<double>& out)
{
    std::vector<Piece> pieces;
    if (!for_each_piece(fname, pieces, [=](Piece* pc) { parse_piece(pc, fmt, field, p_factor); })) return false;
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

bool fast_read_pressure(const char* fname, std::vector<std::vector<double> >& comps, int& num_samples,
                        bool reference_quirk)
{
    std::vector<PressurePiece> pieces;
    if (!for_each_piece(fname, pieces, [](PressurePiece* pc) { parse_pressure_piece(pc); })) return false;
    size_t lines = 0;
    for (const auto& pc : pieces) {
        if (!pc.ok) return false;
        lines += pc.values.size() / 9;
    }
    // viscosity.cpp:120 starts every component with ten zeros and :162-164 copies the first numSamples entries only
    const size_t lead = reference_quirk ? 10 : 0;
    comps.assign(9, std::vector<double>());
    for (int i = 0; i < 9; ++i) {
        std::vector<double>& c = comps[i];
        c.reserve(lines + lead);
        c.assign(lead, 0.0);
        for (const auto& pc : pieces)
            for (size_t r = 0; r + 8 < pc.values.size(); r += 9) c.push_back(pc.values[r + i]);
        c.resize(lines);
    }
    num_samples = (int)lines;
    return true;
}

bool fast_read_traj(const char* fname, bool namd, std::vector<double>& x, std::vector<double>& y,
                    std::vector<double>& z)
{
    std::vector<TrajPiece> pieces;
    if (!for_each_piece(fname, pieces, [namd](TrajPiece* pc) { parse_traj_piece(pc, namd); })) return false;
    size_t total = 0;
    for (const auto& pc : pieces) {
        if (!pc.ok) return false;
        total += pc.values.size();
    }
    x.clear();
    y.clear();
    z.clear();
    if (!namd) {
        x.reserve(total);
        for (const auto& pc : pieces) x.insert(x.end(), pc.values.begin(), pc.values.end());
        return true;
    }
    x.reserve(total / 3);
    y.reserve(total / 3);
    z.reserve(total / 3);
    for (const auto& pc : pieces)
        for (size_t r = 0; r + 2 < pc.values.size(); r += 3) {
            x.push_back(pc.values[r]);
            y.push_back(pc.values[r + 1]);
            z.push_back(pc.values[r + 2]);
        }
    return true;
}

}  // namespace acfhost
