// ACFcalculator -- drop-in front-end: same command line, readers, stdout lines, ACF file and .out file as
// /root/reference/ACFcalculator.cpp (main, :532-912); the numeric hot path (:712-725, :775) runs on the GPU
// through the C ABI of libacf_b200.so (include/acf_b200.h).  There is no CPU implementation of that path
// in this program: if the library reports an error the message goes to stderr and the exit status is 2.
//
// Deliberate differences from the shipped binary:
//   * -r / --factor is accepted as an alias of -p / --p_factor (documented upstream, never implemented);
//   * a CUDA failure ends the run with exit status 2 (the reference has no such failure mode).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../../include/acf_b200.h"
#include "fast_readers.hpp"
#include "gpu_warmup.hpp"
#include "gle.hpp"
#include "options.hpp"
#include "readers.hpp"

using namespace acfhost;

namespace {
const double kAcfWarning = 0.02;   // ACFcalculator.cpp:45
const double kVacfWarning = 0.02;  // :46

std::vector<OptionSpec> option_table()
{
    typedef OptionSpec O;
    return {
        {"help", 'h', O::Flag, false, nullptr, "produce help message", false},
        {"input", 'i', O::String, true, nullptr, "file name of time series", false},
        {"type", 't', O::String, false, nullptr, "type of time series (namd, charmm, gromacs, general)", false},
        {"cutoff", 'c', O::Double, false, "0", "cutoff to integrate ACF", false},
        {"acf", 'a', O::String, true, nullptr, "file name to save autocorrelation functions in", false},
        {"output", 'o', O::String, true, nullptr, "file name to save output to", false},
        {"timestep", 's', O::Double, false, "1", "time between samples in time series file (fs)", false},
        {"maxcorr", 'm', O::Int, false, "1000", "maximum number of time steps to calculate correlation functions over",
         false},
        {"field", 'f', O::Int, false, "1", "index of field to read from time series file", false},
        {"p_factor", 'p', O::Double, false, "1",
         "factor for position to Angstrom when using general, default 1. ex: for nm to Angstrom use 10", false},
        {"spring", 'k', O::Double, false, "0", "spring constant of harmonic restraint (units:kcal A-2 /mol, optional)",
         false},
        {"mass", 'w', O::Double, false, "0", "mass of the solute (units amu, optional)", false},
        {"temperature", 'e', O::Double, false, "0", "temperature (optional)", false},
        {"factor", 'r', O::Double, false, "1", "alias of --p_factor (README.md:14)", true},
        {"fft", '\0', O::String, false, nullptr, "use fft to calculate correlation functions (ACFcalculator.cpp:571)", true},
    };
}

int gpu_failure(const char* what)
{
    std::cerr << "ACFcalculator: " << what << ": " << acfb200_last_error() << std::endl;
    return 2;
}
}  // namespace

int main(int argc, char* argv[])
{
    Options opt(option_table());
    std::string fname, acf_fname, output_fname;
    enum InputType { gromacs, namd, charmm, general } type = general;
    double cutoff = 0.0, timestep = 1.0, p_factor = 1.0, k = 0.0, mass = 0.0, temperature = 0.0;
    int ncorr = 1000, field = 1;
    try {
        opt.parse(argc, argv);
        if (opt.count("help")) {
            std::cout << "Usage: ACFcalculator [options]" << std::endl;
            std::cout << opt.help_text();
            return 1;
        }
        opt.notify();
        fname = opt.get_string("input");
        if (opt.count("type")) {
            const std::string t = opt.get_string("type");
            if (t == "gromacs")
                type = gromacs;
            else if (t == "charmm")
                type = charmm;
            else if (t == "namd")
                type = namd;
            else if (t == "general")
                type = general;
            else {
                std::cout << "Type of file not recognized" << std::endl;
                return 1;
            }
        } else {
            std::cout << "#Type of file not provided, defaulting to general file reader" << std::endl;
            type = general;
        }
        acf_fname = opt.get_string("acf");
        timestep = opt.get_double("timestep", 1.0);
        ncorr = opt.get_int("maxcorr", 1000);
        field = opt.get_int("field", 1);
        p_factor = opt.get_double("p_factor", opt.get_double("factor", 1.0));
        k = opt.get_double("spring", 0.0);
        mass = opt.get_double("mass", 0.0);
        temperature = opt.get_double("temperature", 0.0);
        if (opt.count("fft")) {  // the reference's FFTW-only switch (:644-651); here it selects the GPU FFT path
            std::cout << "#Correlation functions will be calculated using FFT" << std::endl;
            acfb200_set_method(1);
        }
        std::cout << "#Time step of " << timestep << " fs will be used." << std::endl;
        cutoff = opt.get_double("cutoff", 0.0);
        if (cutoff > 0.0)
            std::cout << "#ACF cutoff of " << cutoff << " will be used." << std::endl;
        else {
            std::cout << "#No cutoff will be used to integrate the ACF." << std::endl;
            cutoff = 0;
        }
        output_fname = opt.get_string("output");
        std::cout << "#D(s) will be saved in " << output_fname << std::endl;
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }

    GpuWarmup warm;  // the CUDA context comes up while the file is parsed
    std::vector<double> series;
    int num_samples = 0;
    char* fname_c = const_cast<char*>(fname.c_str());
    // mapped, multi-threaded reader for files the reference reads without incident; anything unusual falls back to
    // the line-by-line readers below (ACFB200_FAST_READER=0 disables it, ACFB200_READER_DEBUG=1 says which ran)
    const char* fr = getenv("ACFB200_FAST_READER");
    const bool dbg = getenv("ACFB200_READER_DEBUG") != nullptr;
    bool fast_done = false;
    if (!(fr && fr[0] == '0')) {
        const FastFormat ff = type == gromacs ? FAST_GROMACS : type == charmm ? FAST_CHARMM : type == namd ? FAST_NAMD
                                                                                                           : FAST_GENERAL;
        fast_done = fast_read_series(fname_c, ff, field, p_factor, series);
        if (fast_done) num_samples = (int)series.size();
    }
    if (dbg) std::cerr << "#reader: " << (fast_done ? "fast" : "line-by-line") << std::endl;
    if (fast_done) {
    } else if (type == gromacs)
        series = read_series_gromacs(fname_c, num_samples, field);
    else if (type == charmm)
        series = read_series_charmm(fname_c, num_samples);
    else if (type == namd)
        series = read_series_namd(fname_c, num_samples, field);
    else
        series = read_series_general(fname_c, num_samples, field, p_factor);
    if (series.size() == 0) {
        std::cerr << "Error: Time series could not be read from file." << std::endl;
        return 1;
    }

    // ---- hot path on the GPU: mean, velocity, acf, vacf, var, varVel (ACFcalculator.cpp:712-725) ----
    std::vector<double> acf(ncorr > 0 ? ncorr : 0), vacf(ncorr > 0 ? ncorr : 0);
    acfb200_result res;
    warm.wait();
    if (acfb200_acf_pipeline(series.data(), num_samples, ncorr, timestep, cutoff, acf.data(), vacf.data(), &res) !=
        ACFB200_OK)
        return gpu_failure("autocorrelation pipeline failed");
    double var = res.var, var_vel = res.var_vel;
    std::cout << "#var " << var << std::endl;
    std::cout << "#varVel " << var_vel << std::endl;

    // optional analytic variances from the spring constant (:730-753)
    if (k > 0.0) {
        acfb200_spring_rescale(acf.data(), vacf.data(), ncorr, k, mass, temperature, &var, &var_vel);
        std::cout << "#varNew " << var << std::endl;
        std::cout << "#varVelNew " << var_vel << std::endl;
    }
    if (acf[ncorr - 1] / var > kAcfWarning)
        std::cout << "#WARNING: ACF has only decayed to " << acf[ncorr - 1] / var * 100
                  << "% of initial value at the end of the region of integration. " << std::endl;
    if (vacf[ncorr - 1] / var > kVacfWarning)  // divided by var, as the reference does (:759)
        std::cout << "#WARNING: VACF has only decayed to " << vacf[ncorr - 1] / var * 100
                  << "% of initial value at the end of the region of integration. " << std::endl;

    {
        std::ofstream acf_file(acf_fname.c_str());
        for (int i = 0; i < ncorr; ++i) acf_file << i << " " << acf[i] << " " << vacf[i] << std::endl;
        acf_file.close();
    }

    // cutoff integral of the (possibly rescaled) ACF (:775)
    double acf_int = res.acf_int;
    if (k > 0.0 &&
        acfb200_integrate_corr_cutoff(acf.data(), ncorr, timestep, cutoff, &acf_int, nullptr) != ACFB200_OK)
        return gpu_failure("cutoff integral failed");
    std::ofstream out_file(output_fname.c_str(), std::ofstream::out);  // created before the stage that may abort

    GleInput gin = {vacf.data(), ncorr, var, var_vel, timestep};
    const GleResult g = run_gle_stage(gin);  // throws like the reference when the roots are not bracketed

    out_file << "#m= " << g.m << std::endl;
    out_file << "#b= " << g.b << std::endl;
    out_file << "#r2= " << g.r2 << std::endl;
    out_file << "#ddDs_min = " << g.dd_ds_min << std::endl;
    out_file << "#avg = " << res.avg << std::endl;
    out_file << "#D (ACF) = " << var * var / acf_int * 0.1 << " cm2/s " << std::endl;
    out_file << "#Ds (VACF) = " << g.b * 0.1 << " cm2/s " << std::endl;
    out_file.close();
    return 0;
}
