// ACFcalculator -- drop-in front-end: same command line, readers, stdout lines, ACF file and .out file as
// /root/reference/ACFcalculator.cpp (main, :532-912); the numeric hot path (:712-725, :775) runs on the GPU
// through the C ABI of libacf_b200.so (include/acf_b200.h).  There is no CPU implementation of that path
// in this program: if the library reports an error the message goes to stderr and the exit status is 2.
//
// Deliberate differences from the shipped binary:
//   * -r / --factor is accepted as an alias of -p / --p_factor (documented upstream, never implemented);
//   * a CUDA failure ends the run with exit status 2 (the reference has no such failure mode);
//   * ACFcalculator --batch FILE [--gpus G] (an extension; see run_batch below) does the work of the reference's
//     setup.sh loop -- one ACFcalculator run per umbrella window -- in ONE process: every line of FILE holds the
//     arguments of one run, all windows go through acfb200_acf_pipeline_batch (or its multi-GPU form) together, and
//     each run's stdout lines, ACF file and .out file are what the single run writes.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <new>
#include <sstream>
#include <string>
#include <thread>
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

enum InputType { gromacs, namd, charmm, general };

// One ACFcalculator run: its options, its series, what the GPU returned for it, and where its text goes.
struct Job {
    std::string fname, acf_fname, output_fname;
    InputType type = general;
    double cutoff = 0.0, timestep = 1.0, p_factor = 1.0, k = 0.0, mass = 0.0, temperature = 0.0;
    int ncorr = 1000, field = 1;
    bool fft = false;
    std::vector<double> series;
    int num_samples = 0;
    std::vector<double> acf, vacf;
    acfb200_result res;
    int status = 0;  // exit status of this run so far (0 = still going)
    int gle_threads = 0;  // host threads the GLE stage may use (0 = all; batch mode shares them between the runs)
};

// Option handling of main() (:554-686) with its stdout lines; returns the exit status the reference would end with
// at this point (1 after --help or an option error) or 0 to go on.
int parse_job(int argc, const char* const* argv, Job& j, std::ostream& out, std::ostream& err)
{
    Options opt(option_table());
    try {
        opt.parse(argc, argv);
        if (opt.count("help")) {
            out << "Usage: ACFcalculator [options]" << std::endl;
            out << opt.help_text();
            return 1;
        }
        opt.notify();
        j.fname = opt.get_string("input");
        if (opt.count("type")) {
            const std::string t = opt.get_string("type");
            if (t == "gromacs")
                j.type = gromacs;
            else if (t == "charmm")
                j.type = charmm;
            else if (t == "namd")
                j.type = namd;
            else if (t == "general")
                j.type = general;
            else {
                out << "Type of file not recognized" << std::endl;
                return 1;
            }
        } else {
            out << "#Type of file not provided, defaulting to general file reader" << std::endl;
            j.type = general;
        }
        j.acf_fname = opt.get_string("acf");
        j.timestep = opt.get_double("timestep", 1.0);
        j.ncorr = opt.get_int("maxcorr", 1000);
        j.field = opt.get_int("field", 1);
        j.p_factor = opt.get_double("p_factor", opt.get_double("factor", 1.0));
        j.k = opt.get_double("spring", 0.0);
        j.mass = opt.get_double("mass", 0.0);
        j.temperature = opt.get_double("temperature", 0.0);
        if (opt.count("fft")) {  // the reference's FFTW-only switch (:644-651); here it selects the GPU FFT path
            out << "#Correlation functions will be calculated using FFT" << std::endl;
            j.fft = true;
        }
        out << "#Time step of " << j.timestep << " fs will be used." << std::endl;
        j.cutoff = opt.get_double("cutoff", 0.0);
        if (j.cutoff > 0.0)
            out << "#ACF cutoff of " << j.cutoff << " will be used." << std::endl;
        else {
            out << "#No cutoff will be used to integrate the ACF." << std::endl;
            j.cutoff = 0;
        }
        j.output_fname = opt.get_string("output");
        out << "#D(s) will be saved in " << j.output_fname << std::endl;
    } catch (const std::exception& e) {
        err << e.what() << std::endl;
        return 1;
    }
    return 0;
}

// The reader call of main() (:688-710).  Mapped, multi-threaded reader for files the reference reads without incident;
// anything unusual falls back to the line-by-line readers (ACFB200_FAST_READER=0 disables the mapped one,
// ACFB200_READER_DEBUG=1 says which ran).
int read_job(Job& j, std::ostream& err)
{
    char* fname_c = const_cast<char*>(j.fname.c_str());
    const char* fr = getenv("ACFB200_FAST_READER");
    const bool dbg = getenv("ACFB200_READER_DEBUG") != nullptr;
    bool fast_done = false;
    if (!(fr && fr[0] == '0')) {
        const FastFormat ff = j.type == gromacs  ? FAST_GROMACS
                              : j.type == charmm ? FAST_CHARMM
                              : j.type == namd   ? FAST_NAMD
                                                 : FAST_GENERAL;
        fast_done = fast_read_series(fname_c, ff, j.field, j.p_factor, j.series);
        if (fast_done) j.num_samples = (int)j.series.size();
    }
    if (dbg) err << "#reader: " << (fast_done ? "fast" : "line-by-line") << std::endl;
    if (fast_done) {
    } else if (j.type == gromacs)
        j.series = read_series_gromacs(fname_c, j.num_samples, j.field);
    else if (j.type == charmm)
        j.series = read_series_charmm(fname_c, j.num_samples);
    else if (j.type == namd)
        j.series = read_series_namd(fname_c, j.num_samples, j.field);
    else
        j.series = read_series_general(fname_c, j.num_samples, j.field, j.p_factor);
    if (j.series.size() == 0) {
        err << "Error: Time series could not be read from file." << std::endl;
        return 1;
    }
    // one sample: velocity_series() asks for new double[n - 2] = new double[-1] (:233) and the reference dies there
    if (j.series.size() == 1) throw std::bad_array_new_length();
    return 0;
}

int gpu_failure(const char* what, std::ostream& err)
{
    err << "ACFcalculator: " << what << ": " << acfb200_last_error() << std::endl;
    return 2;
}

// Everything main() does once acf, vacf and the scalars exist (:726-908): variances, optional rescale, decay warnings,
// the ACF file, the cutoff integral, the GLE stage and the .out file.  The GLE stage throws what the reference throws.
int finish_job(Job& j, std::ostream& out, std::ostream& err)
{
    const int ncorr = j.ncorr;
    double var = j.res.var, var_vel = j.res.var_vel;
    out << "#var " << var << std::endl;
    out << "#varVel " << var_vel << std::endl;

    // optional analytic variances from the spring constant (:730-753)
    if (j.k > 0.0) {
        acfb200_spring_rescale(j.acf.data(), j.vacf.data(), ncorr, j.k, j.mass, j.temperature, &var, &var_vel);
        out << "#varNew " << var << std::endl;
        out << "#varVelNew " << var_vel << std::endl;
    }
    if (j.acf[ncorr - 1] / var > kAcfWarning)
        out << "#WARNING: ACF has only decayed to " << j.acf[ncorr - 1] / var * 100
            << "% of initial value at the end of the region of integration. " << std::endl;
    if (j.vacf[ncorr - 1] / var > kVacfWarning)  // divided by var, as the reference does (:759)
        out << "#WARNING: VACF has only decayed to " << j.vacf[ncorr - 1] / var * 100
            << "% of initial value at the end of the region of integration. " << std::endl;

    {
        std::ofstream acf_file(j.acf_fname.c_str());
        for (int i = 0; i < ncorr; ++i) acf_file << i << " " << j.acf[i] << " " << j.vacf[i] << '\n';  // same bytes as endl
        acf_file.close();
    }

    // cutoff integral of the (possibly rescaled) ACF (:775)
    double acf_int = j.res.acf_int;
    if (j.k > 0.0 &&
        acfb200_integrate_corr_cutoff(j.acf.data(), ncorr, j.timestep, j.cutoff, &acf_int, nullptr) != ACFB200_OK)
        return gpu_failure("cutoff integral failed", err);
    std::ofstream out_file(j.output_fname.c_str(), std::ofstream::out);  // created before the stage that may abort

    GleInput gin = {j.vacf.data(), ncorr, var, var_vel, j.timestep, j.gle_threads};
    const GleResult g = run_gle_stage(gin, out);  // throws like the reference when the roots are not bracketed

    out_file << "#m= " << g.m << std::endl;
    out_file << "#b= " << g.b << std::endl;
    out_file << "#r2= " << g.r2 << std::endl;
    out_file << "#ddDs_min = " << g.dd_ds_min << std::endl;
    out_file << "#avg = " << j.res.avg << std::endl;
    out_file << "#D (ACF) = " << var * var / acf_int * 0.1 << " cm2/s " << std::endl;
    out_file << "#Ds (VACF) = " << g.b * 0.1 << " cm2/s " << std::endl;
    out_file.close();
    return 0;
}

// ---- ACFcalculator --batch FILE [--gpus G] ---------------------------------------------------------------------------
// FILE: one run per line, the arguments exactly as they follow "./ACFcalculator" in setup.sh (a leading program name is
// ignored, so the lines can be pasted from such a script; blank lines and lines starting with '#' are skipped; no shell
// quoting).  Runs that share maxcorr, time step and cutoff go to the GPU in one acfb200_acf_pipeline_batch call
// (acfb200_multi_gpu_pipeline_batch with --gpus G: window w on device w mod G); the host-side GLE stage of the runs is
// spread over the host cores.  Output: for run i a line "#batch run i: <arguments>" followed by the stdout lines of that
// run; its ACF and .out files as named on its line.  A run that fails (option error, unreadable series, input the
// reference crashes on, GLE roots not bracketed) is reported with the exit status the single run would have had and the
// others go on; the program's exit status is that of the first failed run, 0 if none.
std::vector<std::string> split_words(const std::string& line)
{
    std::vector<std::string> w;
    std::istringstream iss(line);
    std::string t;
    while (iss >> t) w.push_back(t);
    return w;
}

int run_batch(const std::string& list_fname, int gpus)
{
    std::ifstream list(list_fname.c_str());
    if (!list) {
        std::cerr << "ACFcalculator: cannot read the batch file " << list_fname << std::endl;
        return 1;
    }
    GpuWarmup warm;  // the CUDA context comes up while the inputs are parsed
    set_reader_faults_throw(true);
    std::vector<Job> jobs;
    std::vector<std::string> lines;
    std::vector<std::ostringstream> outs, errs;
    std::string line;
    while (std::getline(list, line)) {
        std::vector<std::string> words = split_words(line);
        if (words.empty() || words[0][0] == '#') continue;
        if (words[0][0] != '-') words.erase(words.begin());  // "./ACFcalculator -i ..." pasted from a script
        std::string shown;
        for (const auto& w : words) shown += (shown.empty() ? "" : " ") + w;
        lines.push_back(shown);
        jobs.emplace_back();
        outs.emplace_back();
        errs.emplace_back();
        Job& j = jobs.back();
        std::vector<const char*> argv(1, "ACFcalculator");
        for (const auto& w : words) argv.push_back(w.c_str());
        j.status = parse_job((int)argv.size(), argv.data(), j, outs.back(), errs.back());
        if (j.status == 0 && j.ncorr < 1) {
            errs.back() << "ACFcalculator: --maxcorr must be positive in batch mode" << std::endl;
            j.status = 1;
        }
        if (j.status != 0) continue;
        try {
            j.status = read_job(j, errs.back());
        } catch (const ReaderFault& f) {  // the shipped binary segfaults here
            errs.back() << "ACFcalculator: the reference crashes on this input (" << f.what << ")" << std::endl;
            j.status = 139;
        } catch (const std::exception& e) {  // ... and aborts here (uncaught std::out_of_range on a blank line)
            errs.back() << "ACFcalculator: the reference aborts on this input (uncaught exception: " << e.what() << ")"
                        << std::endl;
            j.status = 134;
        }
    }
    const size_t njobs = jobs.size();

    // ---- GPU: one pipeline call per group of runs with the same (maxcorr, time step, cutoff, method)
    std::vector<bool> done(njobs, false);
    warm.wait();
    for (size_t first = 0; first < njobs; ++first) {
        if (done[first] || jobs[first].status != 0) continue;
        const Job& f = jobs[first];
        std::vector<size_t> members;
        for (size_t i = first; i < njobs; ++i)
            if (!done[i] && jobs[i].status == 0 && jobs[i].ncorr == f.ncorr && jobs[i].timestep == f.timestep &&
                jobs[i].cutoff == f.cutoff && jobs[i].fft == f.fft)
                members.push_back(i);
        const int nw = (int)members.size();
        std::vector<const double*> ptrs(nw);
        std::vector<int64_t> lens(nw);
        for (int w = 0; w < nw; ++w) {
            ptrs[w] = jobs[members[w]].series.data();
            lens[w] = jobs[members[w]].num_samples;
        }
        std::vector<double> acf((size_t)nw * f.ncorr), vacf((size_t)nw * f.ncorr);
        std::vector<acfb200_result> res(nw);
        acfb200_set_method(f.fft ? 1 : 0);
        const int rc = gpus == 1 ? acfb200_acf_pipeline_batch(ptrs.data(), lens.data(), nw, f.ncorr, f.timestep, f.cutoff,
                                                              acf.data(), vacf.data(), res.data())
                                 : acfb200_multi_gpu_pipeline_batch(ptrs.data(), lens.data(), nw, f.ncorr, f.timestep,
                                                                    f.cutoff, acf.data(), vacf.data(), res.data(), gpus);
        for (int w = 0; w < nw; ++w) {
            Job& j = jobs[members[w]];
            done[members[w]] = true;
            if (rc != ACFB200_OK) {
                j.status = gpu_failure("autocorrelation pipeline failed", errs[members[w]]);
                continue;
            }
            j.acf.assign(acf.begin() + (size_t)w * f.ncorr, acf.begin() + (size_t)(w + 1) * f.ncorr);
            j.vacf.assign(vacf.begin() + (size_t)w * f.ncorr, vacf.begin() + (size_t)(w + 1) * f.ncorr);
            j.res = res[w];
            std::vector<double>().swap(j.series);
        }
    }

    // ---- host: the rest of every run (files, GLE stage), spread over the host cores
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= njobs) return;
            Job& j = jobs[i];
            if (j.status != 0) continue;
            try {
                j.status = finish_job(j, outs[i], errs[i]);
            } catch (const std::exception& e) {  // the single run ends in std::terminate here (exit status 134)
                errs[i] << "ACFcalculator: the reference aborts here (uncaught exception: " << e.what() << ")" << std::endl;
                j.status = 134;
            }
        }
    };
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    // fewer runs than host threads: each run's GLE stage spreads its D(s) sampling over its share of them
    for (size_t i = 0; i < njobs; ++i) jobs[i].gle_threads = hw > njobs ? (int)(hw / njobs) : 1;
    if (hw > njobs) hw = (unsigned)njobs;
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < hw; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();

    int status = 0;
    for (size_t i = 0; i < njobs; ++i) {
        std::cout << "#batch run " << i << ": " << lines[i] << std::endl;
        std::cout << outs[i].str();
        const std::string e = errs[i].str();
        if (!e.empty()) std::cerr << "#batch run " << i << ": " << e;
        if (jobs[i].status != 0) {
            std::cout << "#batch run " << i << " ended with status " << jobs[i].status << std::endl;
            if (status == 0) status = jobs[i].status;
        }
    }
    return status;
}
}  // namespace

int main(int argc, char* argv[])
{
    // the batch extension is recognised before the reference's option handling, which stays exactly as it was
    for (int i = 1; i < argc; ++i) {
        if (strcmp(argv[i], "--batch") == 0) {
            std::string list_fname;
            int gpus = 1;
            for (int a = 1; a < argc; ++a) {
                if (strcmp(argv[a], "--batch") == 0 && a + 1 < argc)
                    list_fname = argv[++a];
                else if (strcmp(argv[a], "--gpus") == 0 && a + 1 < argc)
                    gpus = atoi(argv[++a]);
                else {
                    std::cerr << "ACFcalculator: --batch FILE [--gpus G] takes no other arguments (" << argv[a] << ")"
                              << std::endl;
                    return 1;
                }
            }
            return run_batch(list_fname, gpus);
        }
    }

    // ACFB200_COLLECT=FILE: record this invocation as one line of FILE instead of running it, so that an existing
    // setup.sh-style script (one process per window) can be replayed as ONE --batch run without editing it:
    //     ACFB200_COLLECT=runs.txt ./setup.sh && ACFcalculator --batch runs.txt
    // The option check is the reference's, so a wrong command line fails here exactly as it would have.
    if (const char* collect = getenv("ACFB200_COLLECT")) {
        Job probe;
        std::ostringstream sink;
        const int prc = parse_job(argc, argv, probe, sink, std::cerr);
        if (prc != 0) {
            std::cout << sink.str();  // --help: the usage text
            return prc;
        }
        std::string line;
        for (int a = 1; a < argc; ++a) {
            if (argv[a][0] == 0 || strpbrk(argv[a], " \t\r\n") != nullptr) {
                std::cerr << "ACFcalculator: ACFB200_COLLECT cannot record an empty argument or one with blanks ("
                          << argv[a] << "): batch files have no quoting" << std::endl;
                return 1;
            }
            line += (a > 1 ? " " : "") + std::string(argv[a]);
        }
        std::ofstream list(collect, std::ofstream::app);
        if (!list || !(list << line << '\n')) {
            std::cerr << "ACFcalculator: cannot append to " << collect << std::endl;
            return 1;
        }
        return 0;
    }

    Job j;
    int rc = parse_job(argc, argv, j, std::cout, std::cerr);
    if (rc != 0) return rc;
    if (j.fft) acfb200_set_method(1);

    GpuWarmup warm;  // the CUDA context comes up while the file is parsed
    rc = read_job(j, std::cerr);
    if (rc != 0) return rc;

    // ---- hot path on the GPU: mean, velocity, acf, vacf, var, varVel (ACFcalculator.cpp:712-725) ----
    j.acf.resize(j.ncorr > 0 ? j.ncorr : 0);
    j.vacf.resize(j.ncorr > 0 ? j.ncorr : 0);
    warm.wait();
    if (acfb200_acf_pipeline(j.series.data(), j.num_samples, j.ncorr, j.timestep, j.cutoff, j.acf.data(), j.vacf.data(),
                             &j.res) != ACFB200_OK)
        return gpu_failure("autocorrelation pipeline failed", std::cerr);
    return finish_job(j, std::cout, std::cerr);
}
