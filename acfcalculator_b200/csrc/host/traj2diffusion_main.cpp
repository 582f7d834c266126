// traj2diffusion -- diffusion coefficient from the mean-square displacement of a trajectory; drop-in for
// /root/reference/traj2diffusion.cpp (main, :357-555): same options, readers (t2d_readers.cpp), stdout
// grammar and exit statuses.  Imaging, the MSD sums, the slope and D run on the GPU through
// libacf_b200.so (acfb200_traj2diffusion); nothing numeric is computed here.
//
// One documented deviation: `-f <anything but namd>` leaves the reference's `type` variable
// uninitialised (:463-470, undefined behaviour); here every value other than "namd" selects the general reader.
//
// Extension: traj2diffusion --batch FILE [--gpus G] -- one run per line of FILE, the arguments as they follow the
// program name in a script (a leading program name is ignored, blank lines and '#' lines skipped, no quoting), optionally
// ending in "> OUTFILE" like the shell redirection such scripts use.  One process and one CUDA context for all
// trajectories; runs that share stride, max, time step and cell go through ONE acfb200_traj2diffusion_batch call
// (acfb200_multi_gpu_traj2diffusion_batch with --gpus G: trajectory t on device t mod G).  Every run's stdout goes to
// its OUTFILE, or to this program's stdout after a "#batch run i: <arguments>" line; a run that fails is reported with
// the status the single run would have had ("#batch run i ended with status s") and the others go on; the exit status
// is the first failure's.
//
// Extension: --method direct|fft|auto (anywhere on the command line, also with --batch) -- acfb200_set_method.  `direct`
// (the default) sums the displacements like the reference; `auto` lets a unit-stride run with a large --max take the
// hybrid Wiener-Khinchin form (include/acf_b200.h, acfb200_msd_last_path: 78x at 1e7 frames, --max 100000) when the
// cost model expects it to be faster, `fft` whenever it applies; both keep MSD within ~1e-12 relative of the direct sums
// or fall back to them.  ACFB200_TRACE=1 reports on stderr which form a single run took.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/acf_b200.h"
#include "fast_readers.hpp"
#include "gpu_warmup.hpp"
#include "options.hpp"
#include "t2d_readers.hpp"

using namespace acfhost;

namespace {
std::vector<OptionSpec> option_table()
{
    typedef OptionSpec O;
    return {
        {"help", 'h', O::Flag, false, nullptr, "produce help message", false},
        {"traj", 't', O::String, true, nullptr, "file name containing the time series (e.g., the colvar.traj file", false},
        {"timestep", 'd', O::Double, false, "1", "time step of the simulation (fs)", false},
        {"type", 'f', O::String, false, nullptr, "type of time series (namd, general)", false},
        {"stride", 's', O::Int, false, "1", "number of steps between the stride", false},
        {"cell", 'c', O::Double, false, nullptr, "length of unit cell", false},
        {"max", 'm', O::Int, false, "1000", "maximum number of steps to calculate MSD over", false},
    };
}

struct Run {
    std::string traj_fname, out_fname;
    bool namd = false, imaging = false;
    double dt = 0.001, L = 0.0;  // DEFAULT_DELTA (:35), overwritten by --timestep's default 1.0 (:442)
    int stride = 100, max_s = 1000;
    Trajectory traj;
    int nframes = 0;
    std::vector<double> msd;
    acfb200_msd_result res;
    int status = 0;  // exit status of this run so far (0 = still going)
    bool finished = false;
};

// Option handling of main() (:396-487).  Returns 0 to go on, otherwise the exit status (1 after --help or an error).
int parse_run(int argc, const char* const* argv, Run& r, std::ostream& out, std::ostream& err)
{
    Options opt(option_table());
    try {
        opt.parse(argc, argv);
        if (opt.count("help")) {
            out << "Usage: ACFcalculator [options]" << std::endl;  // sic (:414)
            out << opt.help_text();
            return 1;
        }
        opt.notify();
        r.traj_fname = opt.get_string("traj");
        r.stride = opt.get_int("stride", 1);
        r.dt = opt.get_double("timestep", 1.0);
        r.max_s = opt.get_int("max", 1000);
        if (opt.count("cell")) {
            r.L = opt.get_double("cell", 0.0);
            r.imaging = true;
        } else {
            out << "#No imaging is performed. This assumes that the simulation code does NOT wrap positions from the "
                   "periodic boundary conditions"
                << std::endl;
        }
        if (opt.count("type")) {
            r.namd = opt.get_string("type") == "namd";
        } else {
            out << "#Type of file not provided, defaulting to general file reader" << std::endl;
        }
    } catch (const std::exception& e) {
        err << e.what() << std::endl;
        return 1;
    }
    return 0;
}

// The reader (:489-496) and the checks before the MSD loop.  May throw what the reference throws (std::out_of_range on
// an empty line of a colvars file, std::bad_alloc for a negative --max).  Returns 0 to go on.
int read_run(Run& r, std::ostream& out, std::ostream& err)
{
    // mapped, multi-threaded reader for files the reference reads without incident; anything unusual falls back to the
    // line-by-line readers (ACFB200_FAST_READER=0 disables it, ACFB200_READER_DEBUG=1 says which ran)
    const char* fr = getenv("ACFB200_FAST_READER");
    bool fast_done = false;
    if (!(fr && fr[0] == '0')) {
        fast_done = fast_read_traj(r.traj_fname.c_str(), r.namd, r.traj.x, r.traj.y, r.traj.z);
        if (fast_done) r.traj.xyz_identical = !r.namd;
    }
    if (getenv("ACFB200_READER_DEBUG")) err << "#reader: " << (fast_done ? "fast" : "line-by-line") << std::endl;
    if (!fast_done) r.traj = r.namd ? read_traj_namd(r.traj_fname.c_str()) : read_traj_general(r.traj_fname.c_str());
    r.nframes = r.traj.frames();
    out << "#" << r.nframes << std::endl;
    if (r.max_s < 0) {
        // new double[negative] throws std::bad_array_new_length, which the reference's catch(std::bad_alloc&) takes (:515-524)
        err << "Error: Could not allocate memory." << std::endl;
        return 1;
    }
    if (r.stride < 1) {
        // the reference never leaves its origin loop for stride <= 0 (:530); refuse instead of hanging
        err << "Error: stride must be a positive number of frames." << std::endl;
        return 1;
    }
    r.msd.assign((size_t)r.max_s, 0.0);
    return 0;
}

// no frames: every msd[j] is 0.0/0 and so are the sum and D (:545, :342-355); nothing to compute
void no_frames(Run& r)
{
    const double nan = -std::numeric_limits<double>::quiet_NaN();
    for (double& v : r.msd) v = nan;
    r.res.sum = r.max_s > 1 ? nan : 0.0;
    r.res.slope = r.max_s > 0 ? r.res.sum / r.max_s : nan;
    r.res.diffusivity = r.res.slope / (r.dt * r.stride) / 6.0;
}

// the MSD table and the four summary lines (:543-552; the reference's printf lines through the same stream)
void print_run(const Run& r, std::ostream& out)
{
    for (int i = 0; i < r.max_s; ++i) out << i * r.dt * r.stride << " " << r.msd[i] << '\n';  // same bytes as endl
    const double D = r.res.diffusivity;
    char buf[256];
    snprintf(buf, sizeof buf, "#sum = %lf\n#D %le (A2/fs)\n#D %le (cm2/s)\n#D %le (m2/s)\n", r.res.sum, D, D * 1E-1, D * 1E-5);
    out << buf;
    out.flush();
}

int gpu_failure(std::ostream& err)
{
    err << "traj2diffusion: " << acfb200_last_error() << std::endl;
    return 2;
}

int run_batch(const std::string& list_fname, int gpus)
{
    std::ifstream list(list_fname.c_str());
    if (!list) {
        std::cerr << "traj2diffusion: cannot read the batch file " << list_fname << std::endl;
        return 1;
    }
    GpuWarmup warm;  // the CUDA context comes up while the trajectories are parsed
    std::vector<Run> runs;
    std::vector<std::string> lines;
    std::vector<std::ostringstream> outs, errs;
    std::string line;
    while (std::getline(list, line)) {
        std::vector<std::string> words;
        {
            std::istringstream iss(line);
            std::string t;
            while (iss >> t) words.push_back(t);
        }
        if (words.empty() || words[0][0] == '#') continue;
        if (words[0][0] != '-') words.erase(words.begin());  // "./traj2diffusion -t ..." pasted from a script
        std::string shown;
        for (const auto& w : words) shown += (shown.empty() ? "" : " ") + w;
        lines.push_back(shown);
        runs.emplace_back();
        outs.emplace_back();
        errs.emplace_back();
        Run& r = runs.back();
        // "> OUTFILE" / ">OUTFILE" at the end of the line
        if (words.size() >= 2 && words[words.size() - 2] == ">") {
            r.out_fname = words.back();
            words.resize(words.size() - 2);
        } else if (!words.empty() && words.back().size() > 1 && words.back()[0] == '>') {
            r.out_fname = words.back().substr(1);
            words.pop_back();
        }
        std::vector<const char*> argv(1, "traj2diffusion");
        for (const auto& w : words) argv.push_back(w.c_str());
        r.status = parse_run((int)argv.size(), argv.data(), r, outs.back(), errs.back());
        if (r.status != 0) continue;
        try {
            r.status = read_run(r, outs.back(), errs.back());
        } catch (const std::exception& e) {  // the single run ends in std::terminate here (exit status 134)
            errs.back() << "traj2diffusion: the reference aborts here (uncaught exception: " << e.what() << ")" << std::endl;
            r.status = 134;
        }
        if (r.status == 0 && (r.nframes == 0 || r.max_s == 0)) {
            no_frames(r);
            r.finished = true;
        }
    }
    const size_t nruns = runs.size();

    // ---- GPU: one call per group of runs with the same (stride, max, time step, cell, imaging)
    warm.wait();
    for (size_t first = 0; first < nruns; ++first) {
        const Run& f = runs[first];
        if (f.finished || f.status != 0) continue;
        std::vector<size_t> members;
        for (size_t i = first; i < nruns; ++i) {
            const Run& r = runs[i];
            if (!r.finished && r.status == 0 && r.stride == f.stride && r.max_s == f.max_s && r.dt == f.dt && r.L == f.L &&
                r.imaging == f.imaging)
                members.push_back(i);
        }
        const int nt = (int)members.size();
        std::vector<const double*> px(nt), py(nt), pz(nt);
        std::vector<int64_t> pn(nt);
        for (int t = 0; t < nt; ++t) {
            const Run& r = runs[members[t]];
            px[t] = r.traj.x.data();
            py[t] = r.traj.xyz_identical ? px[t] : r.traj.y.data();
            pz[t] = r.traj.xyz_identical ? px[t] : r.traj.z.data();
            pn[t] = r.nframes;
        }
        std::vector<double> msd((size_t)nt * f.max_s);
        std::vector<acfb200_msd_result> res(nt);
        const int rc =
            gpus == 1 ? acfb200_traj2diffusion_batch(px.data(), py.data(), pz.data(), pn.data(), nt, f.stride, f.max_s, f.dt,
                                                     f.L, f.imaging ? 1 : 0, msd.data(), nullptr, res.data())
                      : acfb200_multi_gpu_traj2diffusion_batch(px.data(), py.data(), pz.data(), pn.data(), nt, f.stride,
                                                               f.max_s, f.dt, f.L, f.imaging ? 1 : 0, msd.data(), nullptr,
                                                               res.data(), gpus);
        for (int t = 0; t < nt; ++t) {
            Run& r = runs[members[t]];
            r.finished = true;
            if (rc != ACFB200_OK) {
                r.status = gpu_failure(errs[members[t]]);
                continue;
            }
            r.msd.assign(msd.begin() + (size_t)t * f.max_s, msd.begin() + (size_t)(t + 1) * f.max_s);
            r.res = res[t];
            r.traj = Trajectory();  // the coordinates are no longer needed
        }
    }

    int status = 0;
    for (size_t i = 0; i < nruns; ++i) {
        Run& r = runs[i];
        if (r.status == 0) print_run(r, outs[i]);
        std::cout << "#batch run " << i << ": " << lines[i] << std::endl;
        if (!r.out_fname.empty()) {
            std::ofstream f(r.out_fname.c_str());
            if (!f || !(f << outs[i].str())) {
                errs[i] << "traj2diffusion: cannot write " << r.out_fname << std::endl;
                if (r.status == 0) r.status = 1;
            }
        } else {
            std::cout << outs[i].str();
        }
        const std::string e = errs[i].str();
        if (!e.empty()) std::cerr << "#batch run " << i << ": " << e;
        if (r.status != 0) {
            std::cout << "#batch run " << i << " ended with status " << r.status << std::endl;
            if (status == 0) status = r.status;
        }
    }
    return status;
}
}  // namespace

int main(int argc0, char* argv0[])
{
    // --method M is taken out of the command line first; what is left is handled exactly as before
    std::vector<char*> kept;
    for (int i = 0; i < argc0; ++i) {
        if (i > 0 && strcmp(argv0[i], "--method") == 0) {
            const char* m = i + 1 < argc0 ? argv0[i + 1] : "";
            const int method = strcmp(m, "direct") == 0 ? 0 : strcmp(m, "fft") == 0 ? 1 : strcmp(m, "auto") == 0 ? 2 : -1;
            if (method < 0) {
                std::cerr << "traj2diffusion: --method takes direct, fft or auto" << std::endl;
                return 1;
            }
            acfb200_set_method(method);
            ++i;
            continue;
        }
        kept.push_back(argv0[i]);
    }
    const int argc = (int)kept.size();
    char** argv = kept.data();
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
                    std::cerr << "traj2diffusion: --batch FILE [--gpus G] takes no other arguments (" << argv[a] << ")"
                              << std::endl;
                    return 1;
                }
            }
            if (gpus < 1) {
                std::cerr << "traj2diffusion: --gpus must be at least 1" << std::endl;
                return 1;
            }
            return run_batch(list_fname, gpus);
        }
    }

    Run r;
    int rc = parse_run(argc, argv, r, std::cout, std::cerr);
    if (rc != 0) return rc;
    GpuWarmup warm;  // the CUDA context comes up while the trajectory is parsed
    rc = read_run(r, std::cout, std::cerr);
    if (rc != 0) return rc;
    if (r.nframes > 0 && r.max_s > 0) {
        const double* x = r.traj.x.data();
        const double* y = r.traj.xyz_identical ? x : r.traj.y.data();
        const double* z = r.traj.xyz_identical ? x : r.traj.z.data();
        warm.wait();
        if (acfb200_traj2diffusion(x, y, z, r.nframes, r.stride, r.max_s, r.dt, r.L, r.imaging ? 1 : 0, r.msd.data(), nullptr,
                                   &r.res) != ACFB200_OK)
            return gpu_failure(std::cerr);
        if (getenv("ACFB200_TRACE")) {
            int path = 0;
            double cond = 0.0;
            int64_t j0 = 0;
            acfb200_msd_last_path(&path, &cond, &j0);
            std::cerr << "#msd path: " << (path == 1 ? "hybrid" : path == 2 ? "hybrid-refused" : "direct") << " j0 " << j0
                      << " cond " << cond << std::endl;
        }
    } else {
        no_frames(r);
    }
    print_run(r, std::cout);
    return 0;
}
