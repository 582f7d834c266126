// Test driver for the host-side numeric pieces of the GLE stage (csrc/host/exp_cr.cpp, gle.cpp).
//   gle_unit exp          stdin: one double per line as 16 hex digits (bit pattern)
//                         stdout: "<fast> <slow> <array>" bit patterns of exp_cr(x), exp_cr_slow(x) and of
//                         exp_cr_array over the whole input (the AVX2 lanes where the CPU has them)
//   gle_unit toms BUDGET  root of x^3 - 2x - 5 on [2, 3] with the given evaluation budget, twice with ONE budget
//                         variable as ACFcalculator.cpp:780-786 does; prints bracket and the variable after each call
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "exp_cr.hpp"
#include "gle.hpp"

int main(int argc, char** argv)
{
    if (argc >= 2 && !strcmp(argv[1], "exp")) {
        std::vector<double> xs;
        char line[64];
        while (fgets(line, sizeof line, stdin)) {
            const std::uint64_t bits = strtoull(line, nullptr, 16);
            double x;
            memcpy(&x, &bits, 8);
            xs.push_back(x);
        }
        std::vector<double> arr(xs.size());
        acfhost::exp_cr_array(xs.data(), arr.data(), xs.size());
        for (size_t i = 0; i < xs.size(); ++i) {
            const double a = acfhost::exp_cr(xs[i]), b = acfhost::exp_cr_slow(xs[i]);
            std::uint64_t ab, bb, cb;
            memcpy(&ab, &a, 8);
            memcpy(&bb, &b, 8);
            memcpy(&cb, &arr[i], 8);
            printf("%016" PRIx64 " %016" PRIx64 " %016" PRIx64 "\n", ab, bb, cb);
        }
        return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "toms")) {
        std::uintmax_t budget = strtoull(argv[2], nullptr, 10);
        int evals = 0;
        auto f = [&evals](double x) {
            ++evals;
            return x * x * x - 2 * x - 5;
        };
        for (int call = 0; call < 2; ++call) {
            evals = 0;
            const std::pair<double, double> r = acfhost::solve_toms748(f, 2.0, 3.0, 30, budget);
            printf("%.17g %.17g %ju %d\n", r.first, r.second, budget, evals);
        }
        return 0;
    }
    return 2;
}
