// viscosity -- Green-Kubo shear viscosity from the off-diagonal pressure tensor in a NAMD log, drop-in for
// /root/reference/viscosity.cpp (main, :171-224): same reader (including its ten-leading-zeros quirk), the six
// off-diagonal components (i % 4 != 0), maxcorr 1000, 2 fs, box length and temperature as hard-coded there,
// and the same stdout grammar.  The correlations and integrals run on the GPU through libacf_b200.so.
//
//   viscosity <namd.log> [--maxcorr M] [--timestep fs] [--box A] [--temperature K] [--components LIST]
//             [--exact-reader] [--gpus G]
// --components: comma-separated tensor entries, by index 0..8 (row-major, as the log lists them) or name (xy, xz, yx,
// yz, zx, zy, xx, yy, zz); default: the six off-diagonal ones the reference loops over.  "xy,xz,yz" halves the work
// for a symmetric tensor.
// The optional switches are extensions; without them the output is the reference's.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <numeric>
#include <string>
#include <vector>

#include "../../../include/acf_b200.h"
#include "fast_readers.hpp"
#include "gpu_warmup.hpp"
#include "readers.hpp"

int main(int argc, char* argv[])
{
    int ncorr = 1000;                   // viscosity.cpp:173
    double timestep = 2.0;              // :180
    double box_length = 31.0399376148;  // :186
    double temperature = 298.15;        // :188
    bool quirk = true;
    int gpus = 1;  // --gpus G: spread the components (or time blocks of them) over G devices of this box
    std::vector<int> which;
    for (int i = 0; i < 9; ++i)
        if (i % 4 != 0) which.push_back(i);  // off-diagonal components only (:200)
    if (argc < 2) return 1;
    const char* fname = argv[1];
    for (int i = 2; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--exact-reader")
            quirk = false;
        else if (i + 1 < argc && a == "--maxcorr")
            ncorr = atoi(argv[++i]);
        else if (i + 1 < argc && a == "--timestep")
            timestep = atof(argv[++i]);
        else if (i + 1 < argc && a == "--box")
            box_length = atof(argv[++i]);
        else if (i + 1 < argc && a == "--temperature")
            temperature = atof(argv[++i]);
        else if (i + 1 < argc && a == "--gpus")
            gpus = atoi(argv[++i]);
        else if (i + 1 < argc && a == "--components") {
            static const char* const names[9] = {"xx", "xy", "xz", "yx", "yy", "yz", "zx", "zy", "zz"};
            which.clear();
            std::string list = argv[++i];
            for (size_t pos = 0; pos <= list.size();) {
                const size_t end = list.find(',', pos) == std::string::npos ? list.size() : list.find(',', pos);
                const std::string tok = list.substr(pos, end - pos);
                int idx = -1;
                for (int k = 0; k < 9; ++k)
                    if (tok == names[k] || (tok.size() == 1 && tok[0] == '0' + k)) idx = k;
                if (idx < 0) {
                    std::cerr << "viscosity: --components takes entries 0..8 or xx..zz, got '" << tok << "'" << std::endl;
                    return 1;
                }
                which.push_back(idx);
                pos = end + 1;
            }
        }
        else {
            std::cerr << "viscosity: unknown argument " << a << std::endl;
            return 1;
        }
    }
    acfhost::GpuWarmup warm;  // the CUDA context comes up while the log is parsed
    int num_samples = 0;
    // mapped, multi-threaded reader for logs the reference reads without incident; anything unusual falls back to the
    // line-by-line reader (ACFB200_FAST_READER=0 disables it, ACFB200_READER_DEBUG=1 says which ran)
    std::vector<std::vector<double> > comps;
    const char* fr = getenv("ACFB200_FAST_READER");
    bool fast_done = false;
    if (!(fr && fr[0] == '0')) fast_done = acfhost::fast_read_pressure(fname, comps, num_samples, quirk);
    if (getenv("ACFB200_READER_DEBUG")) std::cerr << "#reader: " << (fast_done ? "fast" : "line-by-line") << std::endl;
    if (!fast_done) comps = acfhost::read_pressure_namd(fname, num_samples, quirk);

    std::vector<const double*> ptrs;
    for (int i : which) ptrs.push_back(comps[i].data());
    const int nc = (int)ptrs.size();
    std::vector<double> acf((size_t)nc * ncorr), integrals(nc), eta(nc);
    int rc = ACFB200_OK;
    if (num_samples == 0 && ncorr > 0) {
        // no PRESSURE: line in the log.  Nothing to correlate: the reference's loops give 0/0 at lag 0 and 0/(-k) after it
        // (viscosity.cpp:60-66), and the trapezoid and eta formulas (:95-110, :206) carry the NaN on.
        const double kB = 1.38064852E-23;  // viscosity.cpp:187
        for (int c = 0; c < nc; ++c) {
            double* a = acf.data() + (size_t)c * ncorr;
            volatile double zero = 0.0;  // evaluated at run time: the x86 default NaN prints as "-nan", a folded one as "nan"
            for (int k = 0; k < ncorr; ++k) a[k] = zero / (double)(num_samples - k);
            double I = 0.0;
            for (int i = 0; i < ncorr - 1; ++i) I += 0.5 * (a[i] + a[i + 1]) * timestep;
            integrals[c] = I;
            eta[c] = box_length * box_length * box_length / (kB * temperature) * I * 1E-30 * 1E-15 * 1E10;
        }
    } else {
        warm.wait();
        rc = gpus == 1 ? acfb200_green_kubo(ptrs.data(), nc, num_samples, ncorr, timestep, box_length, temperature,
                                                  acf.data(), integrals.data(), eta.data())
                             : acfb200_multi_gpu_green_kubo(ptrs.data(), nc, num_samples, ncorr, timestep, box_length,
                                                            temperature, acf.data(), integrals.data(), eta.data(), gpus);
    }
    if (rc != ACFB200_OK) {
        std::cerr << "viscosity: " << acfb200_last_error() << std::endl;
        return 2;
    }
    std::vector<double> all;
    for (int c = 0; c < nc; ++c) {
        std::cout << "#" << integrals[c] << std::endl;
        std::cout << "#eta " << which[c] << " " << eta[c] << std::endl;
        all.push_back(eta[c]);
        // same bytes as the reference's per-line std::endl, without a write() per lag
        for (int j = 0; j < ncorr; ++j) std::cout << j << " " << acf[(size_t)c * ncorr + j] << '\n';
        std::cout << std::endl;
    }
    // mean and population standard deviation over the components (:217-222)
    const double sum = std::accumulate(all.begin(), all.end(), 0.0);
    const double mean = sum / all.size();
    const double sq_sum = std::inner_product(all.begin(), all.end(), all.begin(), 0.0);
    const double stdev = sqrt(sq_sum / all.size() - mean * mean);
    std::cout << "#eta = " << mean << " " << stdev << std::endl;
    return 0;
}
