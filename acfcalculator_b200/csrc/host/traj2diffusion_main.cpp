// traj2diffusion -- diffusion coefficient from the mean-square displacement of a trajectory; drop-in for
// /root/reference/traj2diffusion.cpp (main, :357-555): same options, readers (t2d_readers.cpp), stdout
// grammar and exit statuses.  Imaging, the MSD sums, the slope and D run on the GPU through
// libacf_b200.so (acfb200_traj2diffusion); nothing numeric is computed here.
//
// One documented deviation: `-f <anything but namd>` leaves the reference's `type` variable
// uninitialised (:463-470, undefined behaviour); here every value other than "namd" selects the general reader.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
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
}  // namespace

int main(int argc, char* argv[])
{
    Options opt(option_table());
    std::string traj_fname;
    bool namd = false, imaging = false;
    double dt = 0.001, L = 0.0;  // DEFAULT_DELTA (:35), overwritten by --timestep's default 1.0 (:442)
    int stride = 100, max_s = 1000;
    try {
        opt.parse(argc, argv);
        if (opt.count("help")) {
            std::cout << "Usage: ACFcalculator [options]" << std::endl;  // sic (:414)
            std::cout << opt.help_text();
            return 1;
        }
        opt.notify();
        traj_fname = opt.get_string("traj");
        stride = opt.get_int("stride", 1);
        dt = opt.get_double("timestep", 1.0);
        max_s = opt.get_int("max", 1000);
        if (opt.count("cell")) {
            L = opt.get_double("cell", 0.0);
            imaging = true;
        } else {
            std::cout << "#No imaging is performed. This assumes that the simulation code does NOT wrap positions from the "
                         "periodic boundary conditions"
                      << std::endl;
        }
        if (opt.count("type")) {
            namd = opt.get_string("type") == "namd";
        } else {
            std::cout << "#Type of file not provided, defaulting to general file reader" << std::endl;
        }
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }

    GpuWarmup warm;  // the CUDA context comes up while the trajectory is parsed
    // mapped, multi-threaded reader for files the reference reads without incident; anything unusual falls back to the
    // line-by-line readers (ACFB200_FAST_READER=0 disables it, ACFB200_READER_DEBUG=1 says which ran)
    Trajectory traj;
    const char* fr = getenv("ACFB200_FAST_READER");
    bool fast_done = false;
    if (!(fr && fr[0] == '0')) {
        fast_done = fast_read_traj(traj_fname.c_str(), namd, traj.x, traj.y, traj.z);
        if (fast_done) traj.xyz_identical = !namd;
    }
    if (getenv("ACFB200_READER_DEBUG")) std::cerr << "#reader: " << (fast_done ? "fast" : "line-by-line") << std::endl;
    if (!fast_done) traj = namd ? read_traj_namd(traj_fname.c_str()) : read_traj_general(traj_fname.c_str());
    const int nframes = traj.frames();
    std::cout << "#" << nframes << std::endl;
    if (max_s < 0) throw std::bad_alloc();  // new double[negative] (:514)
    if (stride < 1) {
        // the reference never leaves its origin loop for stride <= 0 (:530); refuse instead of hanging
        std::cerr << "Error: stride must be a positive number of frames." << std::endl;
        return 1;
    }

    std::vector<double> msd((size_t)max_s);
    acfb200_msd_result res;
    if (nframes > 0 && max_s > 0) {
        const double* x = traj.x.data();
        const double* y = traj.xyz_identical ? x : traj.y.data();
        const double* z = traj.xyz_identical ? x : traj.z.data();
        warm.wait();
        const int rc = acfb200_traj2diffusion(x, y, z, nframes, stride, max_s, dt, L, imaging ? 1 : 0, msd.data(), nullptr,
                                              &res);
        if (rc != ACFB200_OK) {
            std::cerr << "traj2diffusion: " << acfb200_last_error() << std::endl;
            return 2;
        }
    } else {
        // no frames: every msd[j] is 0.0/0 and so are the sum and D (:545, :342-355); nothing to compute
        const double nan = -std::numeric_limits<double>::quiet_NaN();
        for (double& v : msd) v = nan;
        res.sum = max_s > 1 ? nan : 0.0;
        res.slope = max_s > 0 ? res.sum / max_s : nan;
        res.diffusivity = res.slope / (dt * stride) / 6.0;
    }
    for (int i = 0; i < max_s; ++i) std::cout << i * dt * stride << " " << msd[i] << std::endl;

    const double D = res.diffusivity;
    printf("#sum = %lf\n", res.sum);
    printf("#D %le (A2/fs)\n", D);
    printf("#D %le (cm2/s)\n", D * 1E-1);
    printf("#D %le (m2/s)\n", D * 1E-5);
    return 0;
}
