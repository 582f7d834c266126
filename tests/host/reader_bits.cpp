// Test harness (CPU only): reads one file with the mapped multi-threaded reader and with the faithful line-by-line
// reader and compares the two series bit for bit.  usage: reader_bits <namd|charmm|gromacs|general|pressure|pressure_exact|traj_namd|traj_general> <field> <p> <file>
// prints "same <n>" / "differ at <i>" / "fast declined"; exit 0 / 1 / 2.  Timings go to stderr.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../acfcalculator_b200/csrc/host/fast_readers.hpp"
#include "../../acfcalculator_b200/csrc/host/readers.hpp"
#include "../../acfcalculator_b200/csrc/host/t2d_readers.hpp"

int main(int argc, char** argv)
{
    if (argc != 5) return 3;
    const std::string type = argv[1];
    const int field = atoi(argv[2]);
    const double p = atof(argv[3]);
    const char* fname = argv[4];
    auto now = [] { return std::chrono::steady_clock::now(); };
    if (type == "traj_namd" || type == "traj_general") {  // trajectory readers of the traj2diffusion front-end
        const bool namd = type == "traj_namd";
        std::vector<double> x, y, z;
        auto t0 = now();
        const bool ok = acfhost::fast_read_traj(fname, namd, x, y, z);
        auto t1 = now();
        if (!ok) {
            printf("fast declined\n");
            return 2;
        }
        const acfhost::Trajectory tr = namd ? acfhost::read_traj_namd(fname) : acfhost::read_traj_general(fname);
        auto t2 = now();
        fprintf(stderr, "fast %.3f s, line-by-line %.3f s\n", std::chrono::duration<double>(t1 - t0).count(),
                std::chrono::duration<double>(t2 - t1).count());
        auto same = [](const std::vector<double>& a, const std::vector<double>& b) {
            return a.size() == b.size() && (a.empty() || memcmp(a.data(), b.data(), a.size() * sizeof(double)) == 0);
        };
        if (!same(x, tr.x) || (namd && (!same(y, tr.y) || !same(z, tr.z)))) {
            printf("differ\n");
            return 1;
        }
        printf("same %zu\n", x.size());
        return 0;
    }
    if (type == "pressure" || type == "pressure_exact") {  // NAMD log reader of the viscosity front-end
        const bool quirk = type == "pressure";
        std::vector<std::vector<double> > fastc;
        int nf = 0, ns = 0;
        auto t0 = now();
        const bool ok = acfhost::fast_read_pressure(fname, fastc, nf, quirk);
        auto t1 = now();
        if (!ok) {
            printf("fast declined\n");
            return 2;
        }
        const std::vector<std::vector<double> > slowc = acfhost::read_pressure_namd(fname, ns, quirk);
        auto t2 = now();
        fprintf(stderr, "fast %.3f s, line-by-line %.3f s\n", std::chrono::duration<double>(t1 - t0).count(),
                std::chrono::duration<double>(t2 - t1).count());
        if (nf != ns || fastc.size() != slowc.size()) {
            printf("differ in length %d %d\n", nf, ns);
            return 1;
        }
        for (size_t c = 0; c < fastc.size(); ++c) {
            if (fastc[c].size() != slowc[c].size()) {
                printf("differ in length of component %zu\n", c);
                return 1;
            }
            if (!fastc[c].empty() && memcmp(fastc[c].data(), slowc[c].data(), fastc[c].size() * sizeof(double)) != 0) {
                printf("differ in component %zu\n", c);
                return 1;
            }
        }
        printf("same %d\n", nf);
        return 0;
    }
    acfhost::FastFormat fmt = type == "namd"      ? acfhost::FAST_NAMD
                              : type == "charmm"  ? acfhost::FAST_CHARMM
                              : type == "gromacs" ? acfhost::FAST_GROMACS
                                                  : acfhost::FAST_GENERAL;
    std::vector<double> fast;
    auto t0 = now();
    const bool ok = acfhost::fast_read_series(fname, fmt, field, p, fast);
    auto t1 = now();
    int n = 0;
    std::vector<double> slow = type == "namd"      ? acfhost::read_series_namd(fname, n, field)
                               : type == "charmm"  ? acfhost::read_series_charmm(fname, n)
                               : type == "gromacs" ? acfhost::read_series_gromacs(fname, n, field)
                                                   : acfhost::read_series_general(fname, n, field, p);
    auto t2 = now();
    fprintf(stderr, "fast %.3f s, line-by-line %.3f s\n", std::chrono::duration<double>(t1 - t0).count(),
            std::chrono::duration<double>(t2 - t1).count());
    if (!ok) {
        printf("fast declined\n");
        return 2;
    }
    if (fast.size() != slow.size()) {
        printf("differ in length %zu %zu\n", fast.size(), slow.size());
        return 1;
    }
    for (size_t i = 0; i < fast.size(); ++i)
        if (memcmp(&fast[i], &slow[i], sizeof(double)) != 0) {
            printf("differ at %zu: %.17g %.17g\n", i, fast[i], slow[i]);
            return 1;
        }
    printf("same %zu\n", fast.size());
    return 0;
}
